"""``dfno.utils``: runtime helpers under their reference module path."""
from dfno_b200.parallel.partition import (Partition, create_root_partition,      # noqa: F401
                                          create_standard_partitions)
from dfno_b200.utils.env import get_env                                           # noqa: F401
from dfno_b200.utils.misc import (alphabet, compute_distribution_info, get_gpu_memory,   # noqa: F401
                                  profile_gpu_memory, unit_gaussian_denormalize,
                                  unit_gaussian_normalize, unit_guassian_normalize)
