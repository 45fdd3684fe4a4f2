"""Fused engine on a y-pencil over several B200s: NVLink peer-scatter epilogues (R2/R3),
device flag barrier, p2p gradient all-reduce -- against the single-GPU fused engine and the
portable backend."""
import pytest
import torch

from dfno_b200.utils.testing import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

CFG = dict(in_shape=[1, 2, 16, 16, 16, 2], nt=8, width=8, modes=(4, 4, 4, 3), blocks=2)


def _world() -> int:
    """2 ranks by default; ``DFNO_TEST_WORLD=4`` widens the pencil when the box has the GPUs."""
    import os
    n = int(os.environ.get("DFNO_TEST_WORLD", "2"))
    return n if n in (2, 4) and torch.cuda.device_count() >= n else 2


def _worker(rank, ws, cfg, use_p2p, staged=False, grid=None):
    import os
    os.environ["DFNO_STAGED_SCATTER"] = "1" if staged else "0"
    import dfno_b200 as d
    from dfno_b200.models.fused import FusedAdam, FusedDistributedFNO
    from dfno_b200.parallel.decomposition import shard_bounds, assemble_slices
    dev = torch.device("cuda", torch.cuda.current_device())
    _, P_x, P_0 = d.create_standard_partitions(tuple(grid) if grid else (1, 1, 1, ws, 1, 1))
    P_1 = d.Partition([rank], [1] * 6)
    torch.manual_seed(5)
    ref = d.DistributedFNO(P_1, cfg["in_shape"], cfg["nt"], cfg["width"], cfg["modes"], num_blocks=cfg["blocks"],
                           device=dev, dtype=torch.float32, backend="torch")
    state = d.gather_global_state(ref, to_all=True)
    net = FusedDistributedFNO(P_x, cfg["in_shape"], cfg["nt"], cfg["width"], cfg["modes"],
                              num_blocks=cfg["blocks"], device=dev, use_p2p=use_p2p)
    assert net.staged_scatter == staged
    d.load_global_state(net, state, strict=False)
    g = torch.Generator().manual_seed(9)
    xg = torch.randn(*cfg["in_shape"], generator=g).to(dev)
    oshape = list(cfg["in_shape"]); oshape[1] = 1; oshape[-1] = cfg["nt"]
    tg = torch.randn(*oshape, generator=g).to(dev)
    lo, hi = shard_bounds(cfg["in_shape"], P_x.shape, P_x.index)
    lo_o, hi_o = shard_bounds(oshape, P_x.shape, P_x.index)
    xl, tl = xg[assemble_slices(lo, hi)].contiguous(), tg[assemble_slices(lo_o, hi_o)].contiguous()

    crit = d.DistributedMSELoss(P_x, engine=net)          # peer-memory reduction when use_p2p, NCCL otherwise
    rel_p2p, rel_nccl = d.DistributedRelativeLpLoss(P_x, engine=net), d.DistributedRelativeLpLoss(P_x)
    y_ref = ref(xg)
    ((y_ref - tg) ** 2).mean().backward()
    res = {}
    for it in range(2):                      # twice: buffers/epochs are reused across steps
        net.theta.grad = None
        y = net(xl)
        loss = crit(y, tl)
        loss.backward()
    want = y_ref.detach()[assemble_slices(lo_o, hi_o)]
    la, lb = rel_p2p(y.detach(), tl), rel_nccl(y.detach(), tl)
    if P_0.active:
        assert abs(float(la) - float(lb)) < 1e-5 * abs(float(lb)), (float(la), float(lb))
    res["fwd"] = float((y.detach() - want).norm() / want.norm())
    if P_0.active:
        res["loss"] = abs(float(loss) - float(((y_ref - tg) ** 2).mean())) / float(((y_ref - tg) ** 2).mean())
    # gradients in canonical form
    for p in ref.parameters():
        p.data = p.grad if p.grad is not None else torch.zeros_like(p.data)
    G = d.gather_global_state(ref, to_all=True)
    net.theta.data.copy_(net.theta.grad)
    Gf = d.gather_global_state(net, to_all=True)
    worst = 0.0
    for k in G:
        if k.startswith(("linear", "blocks")) and not k.endswith("linear.b"):
            a, b = Gf[k], G[k]
            a = torch.view_as_real(a) if a.is_complex() else a
            b = torch.view_as_real(b) if b.is_complex() else b
            worst = max(worst, float((a.float().reshape(-1) - b.float().reshape(-1)).norm() / b.float().norm().clamp_min(1e-30)))
    res["grad"] = worst
    # one optimizer step keeps the replicated pointwise weights identical on all ranks
    d.load_global_state(net, state, strict=False)
    opt = FusedAdam(net, lr=1e-2)
    net.theta.grad = None
    crit(net(xl), tl).backward()
    opt.step()
    small = net.theta.data[:net.plan.n_small].clone()
    ref_small = small.clone()
    torch.distributed.broadcast(ref_small, src=0)
    res["replica_drift"] = float((small - ref_small).abs().max())
    torch.cuda.synchronize()
    return res


@pytest.mark.parametrize("use_p2p,staged", [(True, False), (False, False), (True, True)])
def test_two_gpu_pencil_matches_reference(use_p2p, staged):
    """``staged``: per-source staging blocks + local permutation instead of direct interleaved peer stores."""
    n = _world()
    for r in run_distributed(_worker, n, CFG, use_p2p, staged, cuda=True, timeout=300):
        assert r["fwd"] < 5e-2 and r["grad"] < 1e-1, r
        assert r.get("loss", 0) < 5e-2, r
        assert r["replica_drift"] == 0.0, r
