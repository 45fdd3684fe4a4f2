"""Import surface of DistDL used by slimgroup/dfno, forwarded to dfno_b200 (see ../README.md)."""
from . import backend, functional, nn, utilities          # noqa: F401
