"""Multi-process (gloo) tests of the data-movement primitives: values, zero-volume
conventions and adjointness (<Ax, y> == <x, A*y>), which the reference never tests
(SURVEY.md §4)."""
import numpy as np
import pytest
import torch

from dfno_b200.utils.testing import run_distributed


def _bcast_sumreduce(rank, ws):
    import dfno_b200 as d
    _, P_x, P_0 = d.create_standard_partitions((ws,))
    B, S = d.Broadcast(P_0, P_x), d.SumReduce(P_x, P_0)
    torch.manual_seed(0)
    w = torch.rand(3, 4, dtype=torch.float64)
    x = (w.clone() if P_0.active else d.zero_volume_tensor(dtype=torch.float64)).requires_grad_()
    y = B(x)
    assert torch.equal(y, w)
    g = torch.full_like(y, float(rank + 1))
    y.backward(g)
    if P_0.active:
        assert torch.allclose(x.grad, torch.full_like(w, sum(range(1, ws + 1))))
    else:
        assert x.grad.numel() == 0
    # SumReduce forward / backward
    z = torch.full((2, 2), float(rank + 1), dtype=torch.float64, requires_grad=True)
    s = S(z)
    if P_0.active:
        assert torch.allclose(s, torch.full((2, 2), float(sum(range(1, ws + 1))), dtype=torch.float64))
        s.backward(torch.full_like(s, 7.0))
    else:
        assert s.numel() == 0
        s.backward(torch.zeros_like(s))
    assert torch.allclose(z.grad, torch.full_like(z, 7.0))
    return True


@pytest.mark.parametrize("ws", [2, 3])
def test_broadcast_sumreduce_pair(ws):
    assert all(run_distributed(_bcast_sumreduce, ws))


def _repartition(rank, ws, grid_a, grid_b, shape, cplx):
    import dfno_b200 as d
    from dfno_b200.parallel.decomposition import shard_bounds, assemble_slices
    P_w = d.Partition()
    Pa = P_w.create_partition_inclusive(range(int(np.prod(grid_a)))).create_cartesian_topology_partition(grid_a)
    Pb = P_w.create_partition_inclusive(range(int(np.prod(grid_b)))).create_cartesian_topology_partition(grid_b)
    dt = torch.complex128 if cplx else torch.float64
    torch.manual_seed(1)
    G = torch.rand(*shape, dtype=dt)
    H = torch.rand(*shape, dtype=dt)          # cotangent

    def shard(P, T):
        if not P.active:
            return d.zero_volume_tensor(dtype=dt)
        lo, hi = shard_bounds(shape, P.shape, P.index)
        return T[assemble_slices(lo, hi)].clone()

    lazy = d.Repartition(Pa, Pb)                       # discovers shape/dtype on first call
    eager = d.Repartition(Pa, Pb, shape, dtype=dt)
    x = shard(Pa, G).requires_grad_()
    y = lazy(x)
    y2 = eager(shard(Pa, G))
    want = shard(Pb, G)
    assert y.shape == want.shape and torch.equal(y.detach(), want), (rank, y.shape, want.shape)
    assert torch.equal(y2, want)
    # adjoint: <R x, h_b> == <x, R* h_b>; checked globally by summing the local dots
    hb = shard(Pb, H)
    y.backward(hb)
    lhs = (y.detach().conj() * hb).sum() if Pb.active else torch.zeros((), dtype=dt)
    rhs = (x.detach().conj() * x.grad).sum() if Pa.active else torch.zeros((), dtype=dt)
    both = torch.view_as_real(torch.stack([lhs, rhs]).to(torch.complex128)).clone()
    torch.distributed.all_reduce(both)
    assert torch.allclose(both[0], both[1], atol=1e-12)
    # and the backward *is* the reverse repartition of the cotangent
    assert torch.equal(x.grad, shard(Pa, H))
    return True


@pytest.mark.parametrize("grid_a,grid_b,shape,cplx", [
    ((1, 1, 2, 2), (1, 1, 4, 1), (2, 3, 9, 8), False),      # pencil transpose
    ((1, 1, 1, 4), (1, 1, 4, 1), (1, 2, 7, 10), True),      # complex, uneven shards
    ((1, 1, 1, 1), (1, 1, 2, 2), (2, 1, 6, 5), False),      # scatter root -> grid
    ((1, 1, 2, 2), (1, 1, 1, 1), (2, 1, 6, 5), False),      # gather grid -> root
    ((1, 1, 2, 2), (1, 1, 1, 2), (1, 1, 8, 8), False),      # shrinking partition (idle ranks)
])
def test_repartition_values_and_adjoint(grid_a, grid_b, shape, cplx):
    assert all(run_distributed(_repartition, 4, grid_a, grid_b, shape, cplx))


def test_zero_volume_corrector():
    """Empty result -> scalar 0 with an empty gradient; anything else passes through (E7)."""
    import dfno_b200 as d
    e = d.zero_volume_tensor(dtype=torch.float64).requires_grad_()
    out = d.ZeroVolumeCorrectorFunction.apply(e * 2)
    assert out.shape == () and float(out.detach()) == 0.0
    out.backward()
    assert e.grad is not None and e.grad.numel() == 0
    x = torch.tensor(3.0, dtype=torch.float64, requires_grad=True)
    y = d.ZeroVolumeCorrectorFunction.apply(x * x)
    y.backward()
    assert float(y.detach()) == 9.0 and float(x.grad) == 6.0
