"""Kept for path parity with the reference tree: the dataset now lives in the package."""
from dfno_b200.data.datasets import DistributedSleipnerDataset3D   # noqa: F401
