"""Shared-memory Stockham FFT (csrc/fft_radix.cu) against torch.fft, with fused truncation / zero padding."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("N", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("m_frac", [0, 4])
def test_forward_inverse_match_torch_fft(N, m_frac):
    from dfno_b200.ops import fft as F
    g = torch.Generator(device="cuda").manual_seed(N)
    lines = 37 if N < 2048 else 5
    m = 0 if (m_frac == 0 or N < 8) else N // m_frac
    xr = torch.randn(3, lines, N, device="cuda", generator=g)
    xc = torch.randn(3, lines, N, 2, device="cuda", generator=g)
    # complex forward, two-sided truncation
    want = torch.fft.fft(torch.view_as_complex(xc), dim=-1)
    want = want if not m else torch.cat([want[..., :m], want[..., N - m:]], -1)
    got = F.fft_trunc(xc, m)
    assert got.shape == (*xc.shape[:-2], 2 * m if m else N, 2) and rel(got, torch.view_as_real(want)) < 2e-6
    # real forward, one-sided truncation
    keep = m or N // 2 + 1
    want_r = torch.fft.rfft(xr, dim=-1)[..., :keep]
    assert rel(F.rfft_trunc(xr, m), torch.view_as_real(want_r)) < 2e-6
    # zero-padded inverses
    back = F.ifft_pad(got, N)
    full = torch.zeros(3, lines, N, dtype=torch.complex64, device="cuda")
    if m:
        full[..., :m], full[..., N - m:] = want[..., :m], want[..., m:]
    else:
        full = want
    assert rel(back, torch.view_as_real(torch.fft.ifft(full, dim=-1))) < 2e-6
    spec = torch.view_as_real(want_r.clone()).contiguous()
    want_real = torch.fft.irfft(want_r, n=N, dim=-1)
    assert rel(F.irfft_pad(spec, N), want_real) < 2e-6
    if not m:                                                   # un-truncated round trips
        assert rel(F.irfft_pad(F.rfft_trunc(xr), N), xr) < 2e-6
        assert rel(F.ifft_pad(F.fft_trunc(xc), N), xc) < 2e-6


def test_bf16_storage_and_cpu_fallback():
    from dfno_b200.ops import fft as F
    x = torch.randn(11, 256, device="cuda")
    got = F.rfft_trunc(x.to(torch.bfloat16), 32)
    want = torch.view_as_real(torch.fft.rfft(x.to(torch.bfloat16).float(), dim=-1)[..., :32])
    assert got.dtype == torch.bfloat16 and rel(got.float(), want) < 6e-3
    xc = torch.randn(4, 96, 2)                                   # CPU, not a power of two: torch.fft fallback
    assert rel(F.ifft_pad(F.fft_trunc(xc, 8), 96), torch.view_as_real(torch.fft.ifft(
        torch.cat([torch.fft.fft(torch.view_as_complex(xc))[..., :8], torch.zeros(4, 80, dtype=torch.complex64),
                   torch.fft.fft(torch.view_as_complex(xc))[..., 88:]], -1)))) < 1e-5


@pytest.mark.parametrize("dim", [-1, 2])
def test_differentiable_wrappers_match_torch_autograd(dim):
    from dfno_b200.ops.fft import fwd_transform, inv_transform
    g = torch.Generator(device="cuda").manual_seed(3)
    N, m = 64, 12
    shape = [3, 5, 64, 64]
    xr = torch.randn(*shape, device="cuda", generator=g, requires_grad=True)
    xc = torch.randn(*shape, 2, device="cuda", generator=g)
    xc = torch.view_as_complex(xc).requires_grad_()
    d = dim % 4
    cases = [
        (lambda t: fwd_transform(t, d, m, True), lambda t: torch.fft.rfft(t, dim=d).narrow(d, 0, m), xr),
        (lambda t: fwd_transform(t, d, m, False),
         lambda t: torch.cat([torch.fft.fft(t, dim=d).narrow(d, 0, m), torch.fft.fft(t, dim=d).narrow(d, N - m, m)], d), xc),
    ]
    for ours, theirs, x in cases:
        a, b = ours(x), theirs(x)
        assert rel(torch.view_as_real(a), torch.view_as_real(b)) < 3e-6
        w = torch.randn_like(torch.view_as_real(b))
        ga, = torch.autograd.grad((torch.view_as_real(a) * w).sum(), x)
        gb, = torch.autograd.grad((torch.view_as_real(b) * w).sum(), x)
        assert rel(torch.view_as_real(ga) if ga.is_complex() else ga, torch.view_as_real(gb) if gb.is_complex() else gb) < 3e-6
    # inverses (zero padding fused): complex two-sided, real one-sided
    X2 = torch.view_as_complex(torch.randn(*[s if i != d else 2 * m for i, s in enumerate(shape)], 2, device="cuda", generator=g)).requires_grad_()
    X1 = torch.view_as_complex(torch.randn(*[s if i != d else m for i, s in enumerate(shape)], 2, device="cuda", generator=g)).requires_grad_()

    def pad2(t):
        z = t.new_zeros(*[s if i != d else N - 2 * m for i, s in enumerate(shape)])
        return torch.cat([t.narrow(d, 0, m), z, t.narrow(d, m, m)], d)

    def pad1(t):
        z = t.new_zeros(*[s if i != d else N // 2 + 1 - m for i, s in enumerate(shape)])
        return torch.cat([t, z], d)

    for ours, theirs, x in [(lambda t: inv_transform(t, d, N, False), lambda t: torch.fft.ifft(pad2(t), dim=d), X2),
                            (lambda t: inv_transform(t, d, N, True), lambda t: torch.fft.irfft(pad1(t), n=N, dim=d), X1)]:
        a, b = ours(x), theirs(x)
        ar, br = (torch.view_as_real(a), torch.view_as_real(b)) if a.is_complex() else (a, b)
        assert rel(ar, br) < 3e-6
        w = torch.randn_like(br)
        ga, = torch.autograd.grad((ar * w).sum(), x)
        gb, = torch.autograd.grad((br * w).sum(), x)
        assert rel(torch.view_as_real(ga), torch.view_as_real(gb)) < 3e-6


def test_portable_backend_runs_on_the_native_fft():
    """fft_impl='native': the portable backend on a GPU without cuFFT on its path (power-of-two axes)."""
    import dfno_b200 as d
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    kw = dict(num_blocks=2, device=torch.device("cuda"), dtype=torch.float32, backend="torch")
    torch.manual_seed(0)
    a = d.DistributedFNO(P_x, [1, 2, 16, 16, 16, 2], 8, 6, (4, 6, 8, 5), **kw)
    torch.manual_seed(0)
    b = d.DistributedFNO(P_x, [1, 2, 16, 16, 16, 2], 8, 6, (4, 6, 8, 5), fft_impl="native", **kw)
    x = torch.randn(1, 2, 16, 16, 16, 2, device="cuda")
    ya, yb = a(x), b(x)
    assert rel(yb, ya) < 1e-5
    ya.square().mean().backward()
    yb.square().mean().backward()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if p.grad is not None and p.numel():
            gp = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
            gq = torch.view_as_real(q.grad) if q.grad.is_complex() else q.grad
            assert rel(gq, gp) < 1e-4, n
