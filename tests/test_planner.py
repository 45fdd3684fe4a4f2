"""Pencil planner, mode corners and engine eligibility (pure CPU)."""
import numpy as np
import pytest

from dfno_b200.parallel.planner import corner_boxes, make_pencil_plan, spectrum_shape, validate_modes
from dfno_b200.parallel.decomposition import shard_bounds, index_of_rank


@pytest.mark.parametrize("gx,gm,gy", [
    ((1, 1, 1, 8, 1, 1), (1, 1, 1, 8, 1, 1), (1, 1, 1, 1, 1, 8)),      # cfg2: R1/R4 identity
    ((1, 1, 2, 2, 2, 1), (1, 1, 4, 2, 1, 1), (1, 1, 1, 1, 4, 2)),      # cfg3
    ((1, 1, 1, 1, 1, 8), (1, 1, 1, 8, 1, 1), (1, 1, 1, 1, 1, 8)),      # cfg4
    ((1, 1, 8, 8, 4, 1), (1, 1, 32, 8, 1, 1), (1, 1, 1, 1, 32, 8)),    # Perlmutter top row
    ((1, 1, 2, 2, 1), (1, 1, 2, 2, 1), (1, 1, 1, 1, 2)),               # odd n: idle ranks in stage y
])
def test_reference_plan_matches_survey_table(gx, gm, gy):
    p = make_pencil_plan(gx)
    assert p.grid_m == gm and p.grid_y == gy
    n = len(gx) - 2
    assert p.dim_m == tuple(range(2 + (n + 1) // 2, len(gx))) and p.dim_y == tuple(range(2, 2 + (n + 1) // 2))


def test_balanced_plan_uses_every_worker_and_divides_modes():
    spec = spectrum_shape([1, 20, 128, 128, 128, 20], (12, 12, 12, 10))
    assert spec == [1, 20, 24, 24, 24, 10]
    p = make_pencil_plan((1, 1, 1, 8, 1, 1), kind="balanced", spectrum=spec)
    assert int(np.prod(p.grid_y)) == 8 and p.grid_y[2] == p.grid_y[3] == 1
    # reference plan: 10 time modes over 8 ranks -> 2,2,1,1,1,1,1,1 (imbalance noted in SURVEY §2.6)
    ref = make_pencil_plan((1, 1, 1, 8, 1, 1))
    ext = [shard_bounds(spec, ref.grid_y, index_of_rank(ref.grid_y, r)) for r in range(8)]
    assert [hi[5] - lo[5] for lo, hi in ext] == [2, 2, 1, 1, 1, 1, 1, 1]
    bal = [shard_bounds(spec, p.grid_y, index_of_rank(p.grid_y, r)) for r in range(8)]
    sizes = [int(np.prod([h - l for l, h in zip(lo, hi)])) for lo, hi in bal]
    assert max(sizes) == min(sizes)


def test_corners_tile_the_local_slab_in_reference_order():
    modes, spec = (3, 2, 4), [1, 5, 6, 4, 4]
    # whole spectrum on one rank: 2^(n-1) corners, x toggles fastest
    boxes = corner_boxes(spec, modes, [0] * 5, spec)
    assert boxes == [[(0, 3), (0, 2), (0, 4)], [(3, 6), (0, 2), (0, 4)], [(0, 3), (2, 4), (0, 4)], [(3, 6), (2, 4), (0, 4)]]
    # a slab cutting through the x axis keeps only the corners it intersects, in local coordinates
    boxes = corner_boxes(spec, modes, [0, 0, 2, 0, 0], [1, 5, 5, 4, 4])
    assert boxes == [[(0, 1), (0, 2), (0, 4)], [(1, 3), (0, 2), (0, 4)], [(0, 1), (2, 4), (0, 4)], [(1, 3), (2, 4), (0, 4)]]
    hits = np.zeros([3, 4, 4], dtype=int)
    for b in boxes:
        hits[tuple(slice(a, c) for a, c in b)] += 1
    assert (hits == 1).all()


def test_invalid_mode_counts_are_rejected():
    validate_modes([1, 4, 16, 16, 8], (8, 8, 5))
    for bad in [(9, 8, 5), (8, 8, 6)]:
        with pytest.raises(ValueError):
            validate_modes([1, 4, 16, 16, 8], bad)
    with pytest.raises(ValueError):
        validate_modes([1, 4, 16, 16, 7], (4, 4, 3))          # odd time axis


def test_fused_engine_eligibility_rules():
    import dfno_b200 as d
    from dfno_b200.models.fused import supports, EnginePlan
    P6 = d.Partition([0], [1] * 6)
    ok, _ = supports(P6, [1, 1, 128, 128, 128, 1], 20, 20, (12, 12, 12, 10))
    assert ok
    assert not supports(P6, [1, 1, 128, 128, 128, 1], 20, 21, (12, 12, 12, 10))[0]      # width
    assert not supports(P6, [1, 1, 128, 128, 100, 1], 20, 20, (12, 12, 12, 10))[0]      # Z % 8
    assert supports(d.Partition([0], [1] * 5), [1, 1, 64, 64, 1], 20, 20, (12, 12, 10))[0]      # 2-D + time (round 2)
    assert not supports(d.Partition([0], [1] * 4), [1, 1, 64, 1], 20, 20, (12, 10))[0]             # 1-D + time: portable
    pl = EnginePlan(1, 1, 1, 20, 20, 128, 128, 128, (12, 12, 12, 10), world=8, rank=5)
    pl.finish(4)
    assert (pl.Yl, pl.kzl, pl.kz_off, pl.Q) == (16, 3, 15, 3 * 10 * 24 * 24)
    # per-rank sizes of SURVEY §2.6 cfg2, in bf16 instead of fp32/complex64
    assert pl.n_act * 2 == 209715200                      # 210 MB activation per rank
    assert pl.n_S1 * 2 == 39321600                        # R2 payload: 78.6 MB c64 -> 39.3 MB bf16 pairs
    names = list(pl.segments)
    assert names[:4] == ["linear1.W", "linear1.b", "linear2.W", "linear2.b"] and names[-1] == "blocks.3.spectral"
    assert pl.segments["linear4.b"][0] == pl.segments["linear4.W"][0] + 128
