"""Stand-alone stand-in for the DistDL import surface that slimgroup/dfno uses, written on plain
``torch.distributed`` (NCCL on GPUs, gloo on CPUs).  Nothing in this package imports ``dfno_b200``:
it exists so the UNMODIFIED reference under ``baseline/_ref`` can run on a box without MPI
(see ../README.md)."""
from . import backend, functional, nn, utilities          # noqa: F401

__version__ = "0.0-torchdist-shim"
