"""Compatibility namespace: ``import dfno`` gives scripts written against slimgroup/dfno the
same flat public API, served by :mod:`dfno_b200`.

Names re-exported (reference ``dfno/__init__.py`` star-exports its three modules):
models (``DistributedFNO``, ``DistributedFNONd``, ``DistributedFNOBlock``,
``BroadcastedLinear``), losses (``DistributedRelativeLpLoss``, ``DistributedMSELoss``),
partition helpers (``create_standard_partitions``, ``create_root_partition``,
``compute_distribution_info``, ``get_env``) and the small utilities (``alphabet``,
``unit_guassian_normalize``, ``unit_gaussian_denormalize``, ``get_gpu_memory``,
``profile_gpu_memory``) plus the DistDL-style primitives the reference's scripts use
(``Partition``, ``Broadcast``, ``SumReduce``, ``Repartition``, ``DistributedTranspose``,
``zero_volume_tensor``).
"""
from dfno_b200 import *                      # noqa: F401,F403
from dfno_b200 import __version__            # noqa: F401
from dfno_b200.models import fno as dfno     # noqa: F401  (``from dfno.dfno import ...``)
from dfno_b200.models import loss            # noqa: F401
from dfno_b200.utils import misc as utils    # noqa: F401
import sys as _sys

_sys.modules.setdefault("dfno.dfno", dfno)
_sys.modules.setdefault("dfno.loss", loss)
_sys.modules.setdefault("dfno.utils", _sys.modules["dfno_b200.utils"])
