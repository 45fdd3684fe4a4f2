"""``dfno.dfno``: model classes under their reference module path."""
from dfno_b200.models.fno import DistributedFNO, DistributedFNOBlock, DistributedFNONd   # noqa: F401
from dfno_b200.models.linear import BroadcastedAffineOperator, BroadcastedLinear          # noqa: F401
from dfno_b200.parallel.partition import Partition                                       # noqa: F401
