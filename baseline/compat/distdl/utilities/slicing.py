"""``distdl.utilities.slicing`` helpers used by the reference's dataset / utils."""
import numpy as np


def assemble_slices(starts, stops):
    return tuple(slice(int(a), int(b), None) for a, b in zip(starts, stops))


def _cuts(length, workers):
    base, extra = divmod(int(length), int(workers))
    sizes = np.full(int(workers), base, dtype=int)
    sizes[:extra] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def compute_start_index(P_shape, index, shape):
    return np.asarray([_cuts(n, p)[i] for n, p, i in zip(shape, P_shape, index)], dtype=int)


def compute_stop_index(P_shape, index, shape):
    return np.asarray([_cuts(n, p)[i + 1] for n, p, i in zip(shape, P_shape, index)], dtype=int)
