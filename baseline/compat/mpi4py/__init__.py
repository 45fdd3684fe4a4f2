"""``from mpi4py import MPI`` for the unmodified reference, on torch.distributed (see ../README.md)."""
from . import MPI                                           # noqa: F401
