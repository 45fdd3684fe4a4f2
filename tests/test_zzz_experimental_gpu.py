"""Paths that were written after the last GPU session and are still *opt-in* (off by default, see
DESIGN.md section 7).  They are exercised here as non-strict xfail: a pass is reported as XPASS and
means the flag can be dropped; a failure does not turn the suite red because nothing ships on them.
Every case runs in a spawned child process (``run_distributed`` with one rank) so that a device fault or
a hang in an unvalidated kernel configuration cannot take the pytest process -- and the tests before it --
down; the file sorts last for the same reason."""
import os

import pytest
import torch

from dfno_b200.utils.testing import run_distributed

pytestmark = pytest.mark.gpu
experimental = pytest.mark.xfail(strict=False, reason="opt-in path, not yet validated on hardware")


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))


def _padded_t(rank, ws):
    import dfno_b200 as d
    from dfno_b200.models.fused import FusedDistributedFNO
    os.environ["DFNO_FUSED_PADDED_T"] = "1"
    try:
        dev = torch.device("cuda", 0)
        _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
        in_shape, nt, width, modes = [1, 1, 16, 16, 16, 1], 6, 8, (4, 4, 4, 3)
        torch.manual_seed(0)
        ref = d.DistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=2, device=dev, dtype=torch.float32,
                               backend="torch")
        net = FusedDistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=2, device=dev)
        d.load_global_state(net, d.gather_global_state(ref, to_all=True), strict=False)
        x = torch.randn(*in_shape, device=dev)
        y, y_ref = net(x), ref(x)
        assert _rel(y, y_ref) < 8e-2, _rel(y, y_ref)
        t = torch.randn_like(y_ref)
        ((y - t) ** 2).mean().backward()
        assert bool(torch.isfinite(net.theta.grad).all())
        torch.cuda.synchronize()
    finally:
        os.environ.pop("DFNO_FUSED_PADDED_T", None)
    return True


@experimental
def test_padded_t_pitch_t_not_multiple_of_4():
    """T = 6 (like the reference's T = 30): Z1 carries a padded t pitch, G1b reads K = 2T of a 2*Tp row."""
    assert all(run_distributed(_padded_t, 1, cuda=True, timeout=180))


def _lean(rank, ws):
    import dfno_b200 as d
    dev = torch.device("cuda", 0)
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    in_shape = [1, 1, 16, 16, 16, 1]
    torch.manual_seed(0)
    net = d.DistributedFNO(P_x, in_shape, 8, 8, (4, 4, 4, 3), num_blocks=3, device=dev, dtype=torch.bfloat16)
    x = torch.randn(*in_shape, device=dev)
    with torch.no_grad():
        want = net._forward(x, save=True).clone()
        got = net._forward(x, save=False)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-4), float((got - want).abs().max())
    torch.cuda.synchronize()
    return True


@experimental
def test_lean_inference_dataflow_matches_training_dataflow():
    """``_forward(save=False)``: ping-pong activations, no pre-activation write."""
    assert all(run_distributed(_lean, 1, cuda=True, timeout=180))


def _four_rank_worker(rank, ws, grid):
    from test_fused_multigpu import CFG, _worker
    return _worker(rank, ws, CFG, True, False, grid)


@experimental
@pytest.mark.multigpu
@pytest.mark.parametrize("grid", [None, (1, 1, 2, 1, 2, 1), (1, 1, 1, 2, 1, 2)])
def test_four_rank_pencil_and_folds(grid):
    if torch.cuda.device_count() < 4:
        pytest.skip("needs >= 4 GPUs")
    for r in run_distributed(_four_rank_worker, 4, grid, cuda=True, timeout=300):
        assert r["fwd"] < 5e-2 and r["grad"] < 1e-1 and r["replica_drift"] == 0.0, r


def _tanh3(rank, ws):
    """Runs in a child whose environment selects the pre-built ``_build_tanh3`` variant."""
    import dfno_b200 as d
    from dfno_b200.ops import build
    assert build.BUILD_DIR.endswith("_build_tanh3") and build.is_built(), "variant not pre-built"
    x = torch.linspace(-9, 9, 200001, device="cuda")
    y, dy = build.load().gelu_probe(x)
    xd = x.double().requires_grad_()
    ref = torch.nn.functional.gelu(xd)
    ref.sum().backward()
    # error budget relative to bf16 storage (2^-8 of the value): 1e-3 * max(1, |x|) for the value (the MUFU
    # error of tanh is multiplied by x / 2), 4e-3 absolute for the derivative
    scale = x.double().abs().clamp_min(1.0)
    assert float(((y.double() - ref.detach()).abs() / scale).max()) < 1e-3
    assert float((dy.double() - xd.grad).abs().max()) < 4e-3
    dev = torch.device("cuda", 0)
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    in_shape, nt, width, modes = [1, 1, 16, 16, 16, 1], 8, 8, (4, 4, 4, 3)
    torch.manual_seed(0)
    ref_net = d.DistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=2, device=dev, dtype=torch.float32,
                               backend="torch")
    net = d.FusedDistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=2, device=dev)
    d.load_global_state(net, d.gather_global_state(ref_net, to_all=True), strict=False)
    xin = torch.randn(*in_shape, device=dev)
    assert _rel(net(xin), ref_net(xin)) < 8e-2
    torch.cuda.synchronize()
    return True


@experimental
def test_tanh_form_gelu_variant():
    """``DFNO_GELU_TANH3=1`` build: one MUFU.TANH instead of the 17-instruction A&S erf."""
    variant = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dfno_b200", "_build_tanh3")
    if not os.path.isdir(variant):
        pytest.skip("variant not pre-built (DFNO_GELU_TANH3=1 python -c 'from dfno_b200.ops import build; build.build()')")
    os.environ["DFNO_GELU_TANH3"] = "1"             # inherited by the spawned child; this process never loads it
    try:
        assert all(run_distributed(_tanh3, 1, cuda=True, timeout=240))
    finally:
        os.environ.pop("DFNO_GELU_TANH3", None)


def _tma_box(rank, ws):
    """Does a 4-D TMA box land in shared memory as the dense K-major tile with the 128-byte swizzle
    applied to the dense offsets?  (What the permutation-free staged layout would rely on.)"""
    from dfno_b200.ops import build
    C = build.load()

    def expect(dense):                              # dense: [128 rows, 64 bf16] K-major tile
        by = dense.contiguous().view(torch.uint8).reshape(-1).cpu()
        o = torch.arange(by.numel())
        p = o ^ (((o >> 7) & 7) << 4)               # SWIZZLE_128B: 16-byte chunk index ^= row index mod 8
        out = torch.empty_like(by)
        out[p] = by[o]
        return out

    res = {}
    for Yl, box1 in ((16, 2), (32, 1)):             # (32 elems x 2 sources) and (64 elems x 1 source) per row
        A_, P_, X_ = 2, 4, 128
        n = A_ * P_ * X_ * Yl * 2
        src = (torch.arange(n, device="cuda") % 251).to(torch.bfloat16).view(A_, P_, X_, Yl * 2)
        dims = [Yl * 2, P_, X_, A_]
        strides = [X_ * Yl * 2, Yl * 2, P_ * X_ * Yl * 2]
        a, r0 = 1, 2
        raw = C.tma_probe_4d(src.view(-1), dims, strides, [Yl * 2, box1, 128, 1], [0, r0, 0, a]).cpu()
        dense = src[a, r0:r0 + box1].permute(1, 0, 2).reshape(128, box1 * Yl * 2)
        res[f"Yl{Yl}"] = bool(torch.equal(raw, expect(dense)))
    torch.cuda.synchronize()
    assert all(res.values()), res
    return True


@experimental
def test_tma_4d_box_is_a_dense_swizzled_k_major_tile():
    assert all(run_distributed(_tma_box, 1, cuda=True, timeout=180))
