"""``distdl.nn`` layers used by slimgroup/dfno, on WORLD-group torch.distributed collectives."""
from .batchnorm import DistributedBatchNorm                # noqa: F401
from .broadcast import Broadcast                           # noqa: F401
from .loss import DistributedMSELoss                       # noqa: F401
from .repartition import DistributedTranspose, Repartition  # noqa: F401
from .sum_reduce import SumReduce                          # noqa: F401
