"""``distdl.utilities.tensor_decomposition``: balanced block decomposition tables.
Rule: a length-``n`` axis over ``p`` workers gives the first ``n mod p`` workers ``ceil(n/p)``
entries and the rest ``floor(n/p)``.  No ``__all__`` on purpose -- the reference's ``utils.py``
gets ``np`` through ``from ... import *``."""
import numpy as np

from .slicing import _cuts, assemble_slices                # noqa: F401


def compute_subtensor_shapes_balanced(tensor_structure, P_shape):
    grid = tuple(int(p) for p in P_shape)
    dims = [int(s) for s in tensor_structure.shape]
    table = np.zeros(grid + (len(dims),), dtype=int)
    for pos in np.ndindex(*grid):
        table[pos] = [np.diff(_cuts(n, p))[i] for n, p, i in zip(dims, grid, pos)]
    return table


def compute_subtensor_start_indices(shapes):
    grid = shapes.shape[:-1]
    out = np.zeros_like(shapes)
    for pos in np.ndindex(*grid):
        for ax in range(shapes.shape[-1]):
            before = list(pos)
            total = 0
            for i in range(pos[ax]):
                before[ax] = i
                total += shapes[tuple(before)][ax]
            out[pos][ax] = total
    return out


def compute_subtensor_stop_indices(shapes):
    return compute_subtensor_start_indices(shapes) + shapes
