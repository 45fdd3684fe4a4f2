"""Compatibility namespace: ``import dfno`` gives scripts written against slimgroup/dfno the
same flat public API (the reference ``dfno/__init__.py`` star-exports its three modules),
served by :mod:`dfno_b200`: models, losses, partition helpers, the small utilities and the
DistDL-style primitives the reference's scripts reach for (``Partition``, ``Broadcast``,
``SumReduce``, ``Repartition``/``DistributedTranspose``, ``zero_volume_tensor``)."""
from dfno_b200 import *                      # noqa: F401,F403
from dfno_b200 import __version__            # noqa: F401
from . import dfno, loss, utils              # noqa: F401  (``from dfno.utils import ...`` etc.)
