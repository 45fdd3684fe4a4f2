"""Datasets, batch-index helper, Trainer (CPU) and the compat namespace."""
import os
import tempfile

import numpy as np
import pytest
import torch

from dfno_b200.utils.testing import run_distributed


def test_compat_namespace_exports_reference_names():
    import dfno
    for name in ["DistributedFNO", "DistributedFNONd", "DistributedFNOBlock", "BroadcastedLinear",
                 "DistributedRelativeLpLoss", "DistributedMSELoss", "create_standard_partitions",
                 "create_root_partition", "compute_distribution_info", "get_env", "alphabet",
                 "unit_guassian_normalize", "unit_gaussian_denormalize", "get_gpu_memory", "profile_gpu_memory",
                 "Partition", "Broadcast", "SumReduce", "Repartition", "DistributedTranspose",
                 "zero_volume_tensor", "BroadcastedAffineOperator", "DistributedBatchNorm"]:
        assert hasattr(dfno, name), name
    from dfno.utils import alphabet, create_standard_partitions   # noqa: F401
    from dfno.dfno import DistributedFNO                          # noqa: F401
    assert dfno.alphabet(3) == "abc" and dfno.alphabet(2, as_array=True) == ["a", "b"]


def test_normalisers_round_trip():
    import dfno_b200 as d
    x = torch.randn(16, 3, 5)
    xh, mu, std = d.unit_guassian_normalize(x)
    assert torch.allclose(d.unit_gaussian_denormalize(xh, mu, std), x, atol=1e-5)
    assert abs(float(xh.mean())) < 1e-5


def _dataset_worker(rank, ws, tmp):
    import dfno_b200 as d
    _, P_x, P_0 = d.create_standard_partitions((1, 1, 2, ws // 2, 1, 1))
    shape = (8, 6, 4, 5)
    store = d.SyntheticTwoPhaseStore(shape, seed=3)
    ds = d.DistributedFieldDataset(P_x, [1, 2, 3], store, shape, savepath=tmp, filename="s")
    x, y = ds[1]
    # gather to the root and compare with an unpartitioned read
    xg, yg = d.Repartition(P_x, P_0)(x.unsqueeze(0)), d.Repartition(P_x, P_0)(y.unsqueeze(0))
    idx = d.generate_batch_indices(P_x, 10, 3, shuffle=True, seed=5)
    x2, _ = ds[1]                                       # second read comes from the per-rank cache
    assert torch.equal(x, x2) and os.path.exists(os.path.join(tmp, f"s_0002_{P_x.rank:04d}.npz"))
    if P_0.active:
        P1 = d.Partition([0], [1] * 6)
        full = d.DistributedFieldDataset(P1, [1, 2, 3], store, shape)
        xf, yf = full[1]
        assert torch.allclose(xg[0], xf, atol=1e-6) and torch.allclose(yg[0], yf, atol=1e-6)
        assert float(xf.min()) == 0.0 and float(xf[0].max()) == 1.0
    return idx


def test_distributed_dataset_slabs_and_shared_batch_order():
    with tempfile.TemporaryDirectory() as tmp:
        res = run_distributed(_dataset_worker, 4, tmp)
    assert all(r == res[0] for r in res) and sorted(res[0]) == [(0, 3), (3, 6), (6, 9), (9, 10)]


def test_npy_store_reads_only_the_slab():
    import dfno_b200 as d
    with tempfile.TemporaryDirectory() as tmp:
        a = np.arange(4 * 5 * 6, dtype=np.float32).reshape(4, 5, 6)
        np.save(os.path.join(tmp, "permz_7.npy"), a)
        got = d.NpyDirStore(tmp).read(7, "permz", (slice(1, 3), slice(0, 5), slice(2, 4)))
        assert np.array_equal(got, a[1:3, :, 2:4])


def test_trainer_reduces_loss_on_cpu():
    import dfno_b200 as d
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1))
    torch.manual_seed(0)
    net = d.DistributedFNO(P_x, [2, 1, 8, 8, 2], 4, 6, (2, 2, 2), num_blocks=2, dtype=torch.float32)
    opt = torch.optim.Adam(net.parameters(), lr=5e-3)
    tr = d.Trainer(net, d.DistributedRelativeLpLoss(P_x), opt, device=torch.device("cpu"))
    x, y = torch.randn(2, 1, 8, 8, 2), torch.randn(2, 1, 8, 8, 4)
    losses = [tr.step(x, y) for _ in range(25)]
    assert losses[-1] < 0.99 * losses[0] and all(b <= a + 1e-3 for a, b in zip(losses, losses[1:]))
    assert abs(tr.evaluate(x, y) - losses[-1]) < 0.1


def test_rank_logger_and_metrics_writer(tmp_path, capsys):
    import json
    import dfno_b200 as d
    log = d.get_logger("dfno_b200.test")
    log.info("hello")
    d.print0("root line")
    out = capsys.readouterr().out
    assert "r0] hello" in out and "root line" in out
    with d.MetricsWriter(str(tmp_path)) as m:
        m.log(step=3, loss=torch.tensor(0.5), note="x")
        m.log(epoch=1, valid_loss=0.25)
    recs = [json.loads(l) for l in open(tmp_path / "metrics_0000.jsonl")]
    assert recs[0]["step"] == 3 and recs[0]["loss"] == 0.5 and recs[0]["rank"] == 0 and recs[1]["valid_loss"] == 0.25


def test_inference_session_cpu():
    import dfno_b200 as d
    _, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1))
    net = d.DistributedFNO(P_x, [1, 1, 8, 8, 2], 4, 4, (2, 2, 2), num_blocks=1)
    sess = d.InferenceSession(net, device=torch.device("cpu"))
    x = torch.randn(1, 1, 8, 8, 2)
    y = sess.run(x)
    with torch.no_grad():
        assert torch.equal(y, net(x))
    sess.submit(x)
    with pytest.raises(RuntimeError):
        sess.submit(x)
    assert torch.equal(sess.result(), y) and sess.requests == 2
