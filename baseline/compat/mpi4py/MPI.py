import os
import sys

import torch
import torch.distributed as dist

SUM, MIN, MAX = "sum", "min", "max"
_OPS = {SUM: dist.ReduceOp.SUM, MIN: dist.ReduceOp.MIN, MAX: dist.ReduceOp.MAX}


def _boot():
    """Join the job described by the torchrun environment on first use (mpirun would have done this)."""
    if dist.is_initialized() or "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available() and os.environ.get("USE_CUDA", "1") != "0":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        dist.init_process_group("gloo")


class _World:
    def Get_rank(self):
        _boot()
        return dist.get_rank() if dist.is_initialized() else 0

    def Get_size(self):
        _boot()
        return dist.get_world_size() if dist.is_initialized() else 1

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self):
        _boot()
        if dist.is_initialized():
            dist.barrier()

    def allreduce(self, value, op=SUM):
        _boot()
        if not dist.is_initialized():
            return value
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        box = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(box, op=_OPS[op])
        return type(value)(box.item()) if isinstance(value, (int, float)) else box.item()

    def bcast(self, obj, root=0):
        _boot()
        if not dist.is_initialized():
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=root)
        return box[0]

    def Abort(self, code=1):
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(code)


COMM_WORLD = _World()
