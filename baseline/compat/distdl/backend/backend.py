"""``distdl.backend.backend.Partition``: built from a communicator (``Partition(MPI.COMM_WORLD)``,
``/root/reference/dfno/utils.py:79``) instead of a rank list."""
from dfno_b200.parallel.partition import Partition as _Partition


class Partition(_Partition):
    def __init__(self, comm=None, shape=None, ranks=None):
        if ranks is None and comm is not None and not hasattr(comm, "Get_size"):
            ranks, comm = comm, None                       # positional rank list (native signature)
        super().__init__(ranks, shape)

    # sub-partitions must stay instances of this class (the reference calls the same methods on them)
    def create_partition_inclusive(self, ranks):
        p = super().create_partition_inclusive(ranks)
        return Partition(ranks=p.world_ranks, shape=[int(s) for s in p.shape])

    def create_cartesian_topology_partition(self, shape):
        p = super().create_cartesian_topology_partition(shape)
        return Partition(ranks=p.world_ranks, shape=[int(s) for s in p.shape])
