from . import slicing, tensor_decomposition, torch        # noqa: F401
