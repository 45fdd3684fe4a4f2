"""``from mpi4py import MPI`` for a box without MPI: COMM_WORLD is the default torch.distributed group."""
from . import MPI                                           # noqa: F401
