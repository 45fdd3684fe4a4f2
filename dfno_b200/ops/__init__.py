"""Python wrappers (autograd, planning) around the sm_100a kernels in ``dfno_b200/csrc``."""
