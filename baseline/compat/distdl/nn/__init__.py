from dfno_b200.models.loss import DistributedMSELoss                       # noqa: F401
from dfno_b200.models.norm import DistributedBatchNorm                     # noqa: F401
from dfno_b200.parallel.primitives import (Broadcast, DistributedTranspose,  # noqa: F401
                                            Repartition, SumReduce)
