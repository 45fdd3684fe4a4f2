"""Balanced block decomposition helpers with DistDL's names and array conventions
(used by ``/root/reference/dfno/utils.py:58-70``).  No ``__all__``: the reference relies on
``np`` arriving through ``from ... import *`` (``utils.py:80``)."""
import numpy as np

from dfno_b200.parallel.decomposition import assemble_slices, axis_table   # noqa: F401


def compute_subtensor_shapes_balanced(tensor_structure, P_shape):
    """Array of shape ``[*P_shape, ndim]``: the shard shape of every grid position."""
    shape = [int(s) for s in tensor_structure.shape]
    grid = [int(p) for p in P_shape]
    out = np.zeros([*grid, len(shape)], dtype=np.int64)
    for ax, (n, p) in enumerate(zip(shape, grid)):
        ext = axis_table(n, p)                              # [p, 2] (start, stop)
        view = [1] * len(grid)
        view[ax] = p
        out[..., ax] = (ext[:, 1] - ext[:, 0]).reshape(view)
    return out


def compute_subtensor_start_indices(shapes):
    out = np.zeros_like(shapes)
    for ax in range(shapes.shape[-1]):
        ext = np.moveaxis(shapes[..., ax], ax, 0)           # extents along the grid axis `ax`
        starts = np.cumsum(ext, axis=0) - ext
        out[..., ax] = np.moveaxis(starts, 0, ax)
    return out


def compute_subtensor_stop_indices(shapes):
    return compute_subtensor_start_indices(shapes) + shapes
