"""``distdl.utilities.torch``: zero-volume placeholder + TensorStructure (no ``__all__``, see
tensor_decomposition.py)."""
import numpy as np                                          # noqa: F401
import torch                                                # noqa: F401

from dfno_b200.parallel.primitives import zero_volume_tensor   # noqa: F401
from dfno_b200.utils.misc import TensorStructure            # noqa: F401
