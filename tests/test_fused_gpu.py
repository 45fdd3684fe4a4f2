"""Fused sm_100a engine vs the portable fp32 backend on one B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(in_shape, nt, width, modes, blocks=2, seed=0):
    import dfno_b200 as d
    from dfno_b200.models.fused import FusedDistributedFNO
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    ref = d.DistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=blocks, device=dev,
                           dtype=torch.float32, backend="torch")
    fused = FusedDistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=blocks, device=dev)
    d.load_global_state(fused, d.gather_global_state(ref, to_all=True), strict=False)
    return d, ref, fused


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))


@pytest.mark.parametrize("in_shape,nt,width,modes", [
    ([1, 1, 16, 16, 16, 1], 8, 8, (4, 4, 4, 3)),
    ([2, 2, 12, 8, 24, 3], 12, 20, (2, 4, 6, 7)),
    ([1, 1, 256, 256, 8, 1], 4, 8, (4, 6, 2, 2)),          # axes > 128: K = 512 stages, column-part inverse stages
])
def test_forward_backward_match_portable_backend(in_shape, nt, width, modes):
    d, ref, fused = _pair(in_shape, nt, width, modes)
    x = torch.randn(*in_shape, device="cuda")
    y_ref = ref(x)
    y = fused(x)
    assert y.shape == y_ref.shape
    # bf16 storage + random-init cancellation in the 128-term output sum: errors of a few %
    assert _rel(y, y_ref) < 8e-2, _rel(y, y_ref)
    print("forward rel err", _rel(y, y_ref))
    t = torch.randn_like(y_ref)
    ((y_ref - t) ** 2).mean().backward()
    ((y - t) ** 2).mean().backward()
    g_ref = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    for p in ref.parameters():          # canonical form of the reference gradients
        p.data = p.grad if p.grad is not None else torch.zeros_like(p.data)
    G = d.gather_global_state(ref, to_all=True)
    views = {n: v for n, v in fused.named_views().items()}
    gflat = fused.theta.grad
    for name, (off, shape) in fused.plan.segments.items():
        got = gflat[off:off + int(torch.tensor(shape).prod())].view(shape).cpu()
        if name.endswith(".spectral"):
            pl = fused.plan
            want = torch.view_as_real(G[name].permute(0, 1, 4, 5, 3, 2).contiguous()).reshape(shape)
        else:
            want = G[name].reshape(shape)
        assert _rel(got, want) < 1e-1, (name, _rel(got, want))
        print(name, "grad rel err", _rel(got, want))


def test_device_gelu_matches_erf_gelu():
    from dfno_b200.ops import build
    x = torch.linspace(-9, 9, 400001, device="cuda")
    y, dy = build.load().gelu_probe(x)
    xd = x.double().requires_grad_()
    ref = torch.nn.functional.gelu(xd)
    ref.sum().backward()
    assert float((y.double() - ref).abs().max()) < 2e-6
    assert float((dy.double() - xd.grad).abs().max()) < 2e-6


def test_eval_mode_and_state_round_trip():
    d, ref, fused = _pair([1, 1, 16, 16, 16, 1], 8, 8, (4, 4, 4, 3))
    x = torch.randn(1, 1, 16, 16, 16, 1, device="cuda")
    with torch.no_grad():
        y0 = fused(x)
    y1 = fused(x)
    assert torch.allclose(y0, y1.detach())
    state = d.gather_global_state(fused, to_all=True)
    want = d.gather_global_state(ref, to_all=True)
    for k in want:
        if k.startswith(("linear", "blocks")):
            assert torch.allclose(state[k].float() if not state[k].is_complex() else torch.view_as_real(state[k]),
                                  want[k].float() if not want[k].is_complex() else torch.view_as_real(want[k]),
                                  atol=1e-6), k


def test_fused_adam_matches_torch_adam():
    from dfno_b200.models.fused import FusedAdam
    d, ref, fused = _pair([1, 1, 16, 16, 16, 1], 8, 8, (4, 4, 4, 3), blocks=1)
    opt = FusedAdam(fused, lr=1e-2, weight_decay=1e-3)
    p0 = fused.theta.detach().clone()
    pt = torch.nn.Parameter(p0.clone())
    topt = torch.optim.Adam([pt], lr=1e-2, weight_decay=1e-3)
    for i in range(3):
        g = torch.randn_like(p0)
        fused.theta.grad = g.clone()
        pt.grad = g.clone()
        opt.step()
        topt.step()
    assert torch.allclose(fused.theta, pt, atol=1e-6, rtol=1e-5)
