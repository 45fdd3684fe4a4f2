"""Distributed FNO == serial FNO, Taylor gradient test, checkpoint round trips (gloo)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from dfno_b200.utils.testing import run_distributed


def _make(d, P_x, cfg, plan="reference"):
    return d.DistributedFNO(P_x, cfg["in_shape"], cfg["nt"], cfg["width"], cfg["modes"],
                            num_blocks=cfg["blocks"], dtype=torch.float64, plan=plan, backend="torch")


def _global_io(cfg, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(*cfg["in_shape"], dtype=torch.float64, generator=g)
    return x


def _serial_equiv(rank, ws, grid, cfg, plan, tmp):
    import dfno_b200 as d
    from dfno_b200.parallel.decomposition import shard_bounds, assemble_slices
    _, P_x, P_0 = d.create_standard_partitions(grid)
    P_1 = d.Partition([rank], [1] * len(grid))            # a private single-rank world
    torch.manual_seed(7)
    serial = _make(d, P_1, cfg)
    state = d.gather_global_state(serial, to_all=True)
    net = _make(d, P_x, cfg, plan)
    d.load_global_state(net, state)
    xg = _global_io(cfg).requires_grad_()
    yg = serial(xg)
    yg.square().sum().backward()
    out = {}
    if P_x.active:
        lo, hi = shard_bounds(cfg["in_shape"], P_x.shape, P_x.index)
        xl = xg.detach()[assemble_slices(lo, hi)].clone().requires_grad_()
        yl = net(xl)
        oshape = list(cfg["in_shape"]); oshape[1] = 1; oshape[-1] = cfg["nt"]
        lo_o, hi_o = shard_bounds(oshape, P_x.shape, P_x.index)
        want = yg.detach()[assemble_slices(lo_o, hi_o)]
        out["fwd"] = float((yl.detach() - want).abs().max() / want.abs().max())
        yl.square().sum().backward()
        out["dx"] = float((xl.grad - xg.grad[assemble_slices(lo, hi)]).abs().max() / xg.grad.abs().max())
        # parameter gradients: compare in canonical (global) form
    for model in (net, serial):
        for p in model.parameters():
            p.data = p.grad.clone() if p.grad is not None else torch.zeros_like(p.data)
    gd = d.gather_global_state(net, to_all=True)
    gs = d.gather_global_state(serial, to_all=True)
    out["dparam"] = max(float((gd[k] - gs[k]).abs().max() / gs[k].abs().max().clamp_min(1e-30))
                        for k in gs if gs[k].is_floating_point() or gs[k].is_complex())
    # per-rank checkpoint round trip + reshard into the serial layout
    d.load_global_state(net, state)
    d.save_checkpoint(net, tmp, epoch=3, extra={"plan": plan})
    torch.distributed.barrier()
    net2 = _make(d, P_x, cfg, plan)
    info = d.load_checkpoint(net2, tmp, epoch=3)
    assert info["epoch"] == 3
    for (n1, a), (n2, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert n1 == n2 and torch.equal(a, b)
    serial2 = _make(d, P_1, cfg)
    d.reshard_checkpoint(tmp, serial2, epoch=3)
    s2 = d.gather_global_state(serial2, to_all=True)
    out["reshard"] = max(float((s2[k] - state[k]).abs().max()) for k in state
                         if k.startswith(("linear", "blocks")))
    out["nkeys"] = len(net.state_dict())
    return out


CFG_2D = dict(in_shape=[2, 1, 12, 10, 3], nt=8, width=5, modes=(3, 2, 3), blocks=2)
CFG_3D = dict(in_shape=[1, 2, 8, 9, 8, 2], nt=6, width=4, modes=(2, 3, 2, 3), blocks=2)
CFG_4D_T = dict(in_shape=[1, 1, 8, 8, 8, 8], nt=8, width=4, modes=(2, 2, 2, 3), blocks=1)
CFG_3D_8 = dict(in_shape=[1, 1, 8, 8, 8, 2], nt=4, width=4, modes=(2, 2, 2, 2), blocks=1)


@pytest.mark.parametrize("ws,grid,cfg,plan", [
    (4, (1, 1, 2, 2, 1), CFG_2D, "reference"),        # odd n: P_y leaves ranks idle
    (4, (1, 1, 2, 2, 1, 1), CFG_3D, "reference"),
    (4, (1, 1, 1, 4, 1, 1), CFG_3D, "reference"),     # 1 x k pencil: R1/R4 identity
    (4, (1, 1, 1, 4, 1, 1), CFG_3D, "balanced"),
    (2, (2, 1, 1, 1, 1), CFG_2D, "reference"),        # batch (data) parallel axis
    (4, (1, 1, 1, 1, 1, 4), CFG_4D_T, "reference"),   # time-axis partition (BASELINE cfg4 semantics)
    (8, (1, 1, 2, 2, 2, 1), CFG_3D_8, "reference"),   # 2x2x2 spatial partition (BASELINE cfg3 topology)
])
def test_distributed_equals_serial(ws, grid, cfg, plan):
    with tempfile.TemporaryDirectory() as tmp:
        res = run_distributed(_serial_equiv, ws, grid, cfg, plan, tmp)
    for r in res:
        assert r.get("fwd", 0) < 1e-11 and r.get("dx", 0) < 1e-10, r
        assert r["dparam"] < 1e-9 and r["reshard"] == 0.0, r


def _taylor(rank, ws, grid, cfg, names):
    import dfno_b200 as d
    _, P_x, P_0 = d.create_standard_partitions(grid)
    torch.manual_seed(11 + rank)
    net = _make(d, P_x, cfg)
    info = d.compute_distribution_info(P_x, cfg["in_shape"])
    bad = []
    for r in d.gradient_test(net, tuple(int(s) for s in info["shape"]), names=names):
        if not r.ok:
            bad.append(str(r))
    return bad


@pytest.mark.slow
def test_taylor_gradient_2d_64x64_world2():
    """BASELINE.json config 1: 2-D FNO 64x64, 4 layers, 12 modes, world_size=2, CPU/gloo."""
    cfg = dict(in_shape=[1, 1, 64, 64, 4], nt=8, width=8, modes=(12, 12, 4), blocks=4)
    names = ["linear1.W", "linear2.b", "blocks.0.weights.0", "blocks.1.weights.1", "blocks.2.linear.W",
             "blocks.3.weights.0", "linear3.W", "linear4.W", "linear4.b"]
    res = run_distributed(_taylor, 2, (1, 1, 2, 1, 1), cfg, names, timeout=1500)
    assert all(not bad for bad in res), "\n".join(sum(res, []))


def test_taylor_gradient_small_all_params_world4():
    cfg = dict(in_shape=[1, 1, 8, 8, 2], nt=4, width=3, modes=(2, 2, 2), blocks=1)
    res = run_distributed(_taylor, 4, (1, 1, 2, 2, 1), cfg, None)
    assert all(not bad for bad in res), "\n".join(sum(res, []))


def _fnond(rank, ws):
    import dfno_b200 as d
    _, P_x, P_0 = d.create_standard_partitions((1, 1, 2, 1, 1))
    f = d.DistributedFNONd(P_x=P_x, width=4, modes=(2, 2, 2), out_timesteps=4, decomposition_order=1,
                           num_blocks=1, device=torch.device("cpu"), dtype=torch.float64, P_y=P_x)
    y = f(torch.rand(1, 1, 4, 8, 1, dtype=torch.float64))
    crit = d.DistributedRelativeLpLoss(P_x)
    mse = d.DistributedMSELoss(P_x)
    t = torch.rand_like(y)
    l1, l2 = crit(y, t), mse(y, t)
    (l1 + l2).backward()
    # reference values from the gathered tensors
    G = d.Repartition(P_x, P_0)
    yg, tg = G(y.detach()), G(t)
    if P_0.active:
        assert f.net.in_shape == [1, 1, 8, 8, 1]
        want1 = (torch.linalg.vector_norm((yg - tg).reshape(1, -1), dim=1) /
                 torch.linalg.vector_norm(tg.reshape(1, -1), dim=1)).mean()
        assert torch.allclose(l1, want1) and torch.allclose(l2, (yg - tg).square().mean())
    else:
        assert float(l1) == 0.0 and float(l2) == 0.0
    return True


def test_lazy_fnond_and_losses():
    assert all(run_distributed(_fnond, 2))


def _fold(rank, ws, grid):
    """The fused engine's entry/exit fold, with the portable network standing in for the engine:
    net(P_x) must equal R_out(net(P_work)(R_in(x))) in values and input/parameter gradients."""
    import dfno_b200 as d
    from dfno_b200.models.fused import fold_onto_pencil
    from dfno_b200.parallel.decomposition import shard_bounds, assemble_slices
    in_shape, nt, width, modes = [1, 2, 8, 8, 8, 2], 4, 3, (2, 2, 2, 2)
    _, P_x, _ = d.create_standard_partitions(grid)
    P_work, R_in, R_out = fold_onto_pencil(P_x, in_shape, nt)
    assert tuple(int(v) for v in P_work.shape) == (1, 1, 1, ws, 1, 1) and R_in is not None
    torch.manual_seed(3)
    a = d.DistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=1, dtype=torch.float64, backend="torch")
    b = d.DistributedFNO(P_work, in_shape, nt, width, modes, num_blocks=1, dtype=torch.float64, backend="torch")
    d.load_global_state(b, d.gather_global_state(a, to_all=True))
    g = torch.Generator().manual_seed(1)
    xg = torch.randn(*in_shape, dtype=torch.float64, generator=g)
    lo, hi = shard_bounds(in_shape, P_x.shape, P_x.index)
    xa = xg[assemble_slices(lo, hi)].clone().requires_grad_()
    xb = xg[assemble_slices(lo, hi)].clone().requires_grad_()
    ya = a(xa)
    yb = R_out(b(R_in(xb)))
    assert ya.shape == yb.shape
    err = float((ya - yb).abs().max())
    w = torch.randn(ya.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(7 + rank))
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    gerr = float((xa.grad - xb.grad).abs().max()) if xa.numel() else 0.0
    ga, gb = d.gather_global_state(_grads(a), to_all=True), d.gather_global_state(_grads(b), to_all=True)
    perr = max(float((torch.view_as_real(ga[k]) if ga[k].is_complex() else ga[k]).sub(
        torch.view_as_real(gb[k]) if gb[k].is_complex() else gb[k]).abs().max()) for k in ga if ga[k].numel())
    return err, gerr, perr


def _grads(net):
    for p in net.parameters():
        p.data = p.grad if p.grad is not None else torch.zeros_like(p.data)
    return net


@pytest.mark.parametrize("ws,grid", [(2, (1, 1, 2, 1, 1, 1)), (2, (1, 1, 1, 1, 1, 2)), (4, (1, 1, 2, 1, 2, 1)),
                                     (4, (1, 1, 1, 2, 1, 2))])
def test_fold_onto_pencil_is_transparent(ws, grid):
    for err, gerr, perr in run_distributed(_fold, ws, grid):
        assert err < 1e-10 and gerr < 1e-10 and perr < 1e-10, (err, gerr, perr)


def _replica_init(rank, ws):
    import dfno_b200 as d
    _, P_x, _ = d.create_standard_partitions((2, 1, 1, 1, 1))
    torch.manual_seed(0)                                     # identical seeds: the root still consumes extra RNG
    net = d.DistributedFNO(P_x, [2, 1, 8, 8, 1], 4, 3, (2, 2, 2), num_blocks=1, dtype=torch.float64, backend="torch")
    return [float(w.detach().abs().sum()) for w in net.blocks[0].weights]


def test_data_parallel_replicas_start_from_the_same_spectral_weights():
    """on a batch-partitioned P_x every replica drew its own spectral shard and only the gradients
    were synchronised; the replicas must hold equal weights right after construction (no load_global_state)."""
    a, b = run_distributed(_replica_init, 2, timeout=300)
    assert a == b and len(a) > 0, (a, b)
