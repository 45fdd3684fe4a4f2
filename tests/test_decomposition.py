import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from dfno_b200.parallel import decomposition as D
from dfno_b200.parallel.partition import Partition
from dfno_b200.parallel.primitives import build_repartition_plan
from dfno_b200 import compute_distribution_info, create_standard_partitions


def test_balanced_rule_first_workers_get_the_remainder():
    # 10 entries over 8 workers -> 2,2,1,1,1,1,1,1 (SURVEY.md §2.6 cfg2 imbalance)
    assert [D.balanced_extent(10, 8, i) for i in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]
    assert D.balanced_bounds(10, 8, 1) == (2, 4)
    assert D.balanced_bounds(10, 8, 7) == (9, 10)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 200), st.integers(1, 17))
def test_axis_shards_tile_the_axis(n, p):
    tab = D.axis_table(n, p)
    assert tab[0, 0] == 0 and tab[-1, 1] == n
    assert (tab[1:, 0] == tab[:-1, 1]).all()
    ext = tab[:, 1] - tab[:, 0]
    assert ext.max() - ext.min() <= 1 and (np.diff(ext) <= 0).all()


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(1, 9), min_size=1, max_size=4).flatmap(
    lambda shape: st.tuples(st.just(shape), st.tuples(*[st.integers(1, 3) for _ in shape]))))
def test_subtensor_tables_cover_every_cell_once(arg):
    shape, grid = arg
    shapes = D.compute_subtensor_shapes_balanced(shape, grid)
    starts = D.compute_subtensor_start_indices(shapes)
    stops = D.compute_subtensor_stop_indices(shapes)
    hits = np.zeros(shape, dtype=int)
    for idx in D.grid_indices(grid):
        hits[D.assemble_slices(starts[idx], stops[idx])] += 1
        lo, hi = D.shard_bounds(shape, grid, idx)
        assert list(starts[idx]) == lo and list(stops[idx]) == hi
    assert (hits == 1).all()


def test_distribution_info_single_rank():
    _, P_x, P_0 = create_standard_partitions((1, 1, 1))
    info = compute_distribution_info(P_x, [3, 4, 5])
    assert tuple(info["shape"]) == (3, 4, 5) and info["index"] == (0, 0, 0)
    assert info["slice"] == (slice(0, 3, 1), slice(0, 4, 1), slice(0, 5, 1))
    assert P_0.active and P_0.dim == 3


@settings(max_examples=25, deadline=None)
@given(st.sampled_from([((1, 4), (4, 1)), ((2, 2), (1, 4)), ((1, 1), (2, 2)), ((2, 2), (1, 1)),
                        ((4, 1), (2, 1)), ((1, 2), (4, 1))]),
       st.tuples(st.integers(4, 9), st.integers(4, 9)))
def test_repartition_plans_move_every_element_exactly_once(grids, shape):
    """Single-process simulation of the all-to-all-v: pack with the senders' plans, unpack
    with the receivers' plans, compare with the global tensor."""
    ga, gb = grids
    world = max(int(np.prod(ga)), int(np.prod(gb)))
    Pa = Partition(range(int(np.prod(ga))), ga)
    Pb = Partition(range(int(np.prod(gb))), gb)
    G = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
    plans = [build_repartition_plan(Pa, Pb, shape, me=r) for r in range(world)]

    def shard(P, r):
        if r >= P.size:
            return None
        lo, hi = D.shard_bounds(shape, P.shape, P.index_of(r))
        return G[D.assemble_slices(lo, hi)]

    for dst in range(world):
        want = shard(Pb, dst)
        if want is None:
            assert plans[dst].out_shape == (0,)
            continue
        got = np.full(plans[dst].out_shape, -1, dtype=np.int64)
        for src in range(world):
            sp = plans[src]
            sbox = sp.send_boxes[sp.ranks.index(dst)] if dst in sp.ranks else None
            rbox = plans[dst].recv_boxes[plans[dst].ranks.index(src)] if src in plans[dst].ranks else None
            assert (sbox is None) == (rbox is None)
            if sbox is not None:
                got[rbox] = shard(Pa, src)[sbox]
        assert (got == want).all()
