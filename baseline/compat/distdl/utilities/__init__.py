from . import tensor_decomposition, torch                  # noqa: F401
