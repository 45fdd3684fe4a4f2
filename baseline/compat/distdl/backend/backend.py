"""``distdl.backend.backend.Partition`` on torch.distributed.

A partition is an ordered list of world ranks laid out row-major on a Cartesian grid.  All
collectives issued by ``distdl.nn`` run on the WORLD process group (ranks that own nothing
contribute empty messages), so no sub-communicator ever has to be created."""
import numpy as np
import torch
import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized()


def my_world_rank():
    return dist.get_rank() if _on() else 0


def n_world():
    return dist.get_world_size() if _on() else 1


def comm_device():
    """Device collectives must use: the current CUDA device under NCCL, the host under gloo."""
    if _on() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class _Comm:
    """What the scripts reach through ``P._comm``: Barrier / allreduce / rank / size."""

    def __init__(self, members):
        self._members = members

    def Get_rank(self):
        me = my_world_rank()
        return self._members.index(me) if me in self._members else -1

    def Get_size(self):
        return len(self._members)

    rank = property(Get_rank)
    size = property(Get_size)

    def Barrier(self):
        if _on():
            dist.barrier()

    def allreduce(self, value, op="sum"):
        from mpi4py import MPI
        return MPI.COMM_WORLD.allreduce(value, op=op)


class Partition:
    def __init__(self, comm=None, members=None, shape=None):
        if members is None:
            members = list(range(comm.Get_size() if comm is not None else n_world()))
        self._members = [int(r) for r in members]
        self.shape = np.asarray([len(self._members)] if shape is None else [int(s) for s in shape], dtype=int)
        assert int(np.prod(self.shape)) == len(self._members)
        self.dim = len(self.shape)
        self.size = len(self._members)
        me = my_world_rank()
        self.active = me in self._members
        self.rank = self._members.index(me) if self.active else -1   # DistDL: MPI.PROC_NULL-like for inactive
        self.index = tuple(int(i) for i in np.unravel_index(self.rank, self.shape)) if self.active else None
        self._comm = _Comm(self._members)

    # construction -------------------------------------------------------------------------------
    def create_partition_inclusive(self, ranks):
        picks = [int(r) for r in np.asarray(ranks).reshape(-1)]
        return Partition(members=[self._members[r] for r in picks])

    def create_cartesian_topology_partition(self, shape, **_unused):
        shape = [int(s) for s in np.asarray(shape).reshape(-1)]
        count = int(np.prod(shape))
        if count > self.size:
            raise ValueError(f"a {shape} grid needs {count} workers, this partition has {self.size}")
        return Partition(members=self._members[:count], shape=shape)

    # helpers used inside this package -----------------------------------------------------------
    def member(self, grid_index):
        return self._members[int(np.ravel_multi_index(tuple(int(i) for i in grid_index), self.shape))]

    def grid_index_of(self, world_rank):
        if world_rank not in self._members:
            return None
        return tuple(int(i) for i in np.unravel_index(self._members.index(world_rank), self.shape))

    def __eq__(self, other):
        return (isinstance(other, Partition) and self._members == other._members
                and tuple(self.shape) == tuple(other.shape))

    def __hash__(self):
        return hash((tuple(self._members), tuple(int(s) for s in self.shape)))

    def __repr__(self):
        return f"Partition(grid={tuple(int(s) for s in self.shape)}, members={self._members}, rank={self.rank})"
