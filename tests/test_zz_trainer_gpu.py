"""The public training-step API on the fused engine: pinned host batch in, loss out, whole step replayed
from one CUDA graph (what ``bench.py`` times as ``e2e``)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trainer_replays_the_step_from_a_cuda_graph():
    import dfno_b200 as d
    dev = torch.device("cuda", 0)
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    in_shape = [1, 1, 16, 16, 16, 1]
    torch.manual_seed(0)
    net = d.DistributedFNO(P_x, in_shape, 8, 8, (4, 4, 4, 3), num_blocks=2, device=dev, dtype=torch.bfloat16)
    assert isinstance(net, d.FusedDistributedFNO)
    opt = d.FusedAdam(net, lr=1e-2)
    crit = d.DistributedRelativeLpLoss(P_x, engine=net)
    tr = d.Trainer(net, crit, opt, device=dev, cuda_graph=True)
    x = torch.randn(*in_shape).pin_memory()
    y = torch.randn(1, 1, 16, 16, 16, 8).pin_memory()
    before = net.theta.detach().clone()
    losses = [tr.step(x, y, next_batch=(x, y)) for _ in range(6)]
    assert tr._graph is not None, "the step must be capturable (no host sync, no NCCL inside)"
    assert all(math.isfinite(l) and l > 0 for l in losses), losses
    assert losses[-1] < 1.05 * losses[0], losses             # Adam on a fixed batch: no blow-up across replays
    assert not torch.equal(before, net.theta.detach())
    assert tr.h2d_bytes == (x.numel() + y.numel()) * 4 and tr.d2h_bytes == 4
    # evaluation does not touch the weights
    w = net.theta.detach().clone()
    ev = tr.evaluate(x, y)
    assert math.isfinite(ev) and torch.equal(w, net.theta.detach())


def test_inference_session_replays_the_forward():
    import dfno_b200 as d
    dev = torch.device("cuda", 0)
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
    in_shape = [1, 1, 16, 16, 16, 1]
    torch.manual_seed(0)
    net = d.DistributedFNO(P_x, in_shape, 8, 8, (4, 4, 4, 3), num_blocks=2, device=dev, dtype=torch.bfloat16)
    sess = d.InferenceSession(net, device=dev, cuda_graph=True)
    xs = [torch.randn(*in_shape).pin_memory() for _ in range(3)]
    outs = [sess.run(x).clone() for x in xs]
    assert sess._graph is not None and sess.requests == 3
    with torch.no_grad():
        for x, y in zip(xs, outs):
            want = net(x.to(dev)).cpu()
            assert torch.allclose(y, want, atol=1e-5, rtol=1e-4), float((y - want).abs().max())
    assert not torch.equal(outs[0], outs[1])                   # the replay really consumes the new input
