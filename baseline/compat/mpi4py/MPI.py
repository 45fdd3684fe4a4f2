import sys

import torch
import torch.distributed as dist

SUM, MIN, MAX = "sum", "min", "max"


class _World:
    """``MPI.COMM_WORLD``: the default torch.distributed group (a single rank without one)."""

    def _on(self):
        return dist.is_available() and dist.is_initialized()

    def Get_rank(self):
        return dist.get_rank() if self._on() else 0

    def Get_size(self):
        return dist.get_world_size() if self._on() else 1

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self):
        if self._on():
            dist.barrier()

    def allreduce(self, value, op=SUM):
        if not self._on():
            return value
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={SUM: dist.ReduceOp.SUM, MIN: dist.ReduceOp.MIN, MAX: dist.ReduceOp.MAX}[op])
        return type(value)(t.item()) if isinstance(value, (int, float)) else t.item()

    def bcast(self, obj, root=0):
        if not self._on():
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=root)
        return box[0]

    def Abort(self, code=1):
        sys.stdout.flush()
        sys.stderr.flush()
        import os
        os._exit(code)


COMM_WORLD = _World()
