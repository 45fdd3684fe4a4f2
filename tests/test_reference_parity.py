"""The UNMODIFIED reference package (``baseline/_ref/dfno``, installed with ``pip --no-deps`` from
``/root/reference``) against this framework, from the same weights.

The reference cannot import on its own here (DistDL / mpi4py are not installable offline); it runs on
the import-surface layer in ``baseline/compat`` that forwards DistDL's primitives to
``dfno_b200.parallel``.  Everything else on the reference side -- model code, einsums, restrict /
zeropad / ``torch.fft`` calls, per-forward weight broadcasts, the loss -- is the reference's own.
Checked: identical state-dict keys and per-rank shard shapes, outputs, loss and gradients on 1 and
4 ranks (``(1,1,2,2,1,1)``, the reference's in-module demo grid, ``/root/reference/dfno/dfno.py:359``)."""
import os
import sys

import pytest
import torch

from dfno_b200.utils.testing import run_distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
COMPAT = os.path.join(ROOT, "baseline", "compat")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dfno")),
                                reason="reference package not installed under baseline/_ref")


def _import_reference():
    for p in (REF, COMPAT):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [COMPAT, REF]
    sys.modules.pop("dfno", None)
    import dfno as ref
    assert os.path.abspath(ref.__file__).startswith(REF), ref.__file__
    return ref


def _parity(rank, ws, grid, in_shape, nt, width, modes, blocks):
    import warnings
    warnings.filterwarnings("ignore")
    ref = _import_reference()
    import dfno_b200 as d
    _, P_ref, _ = ref.create_standard_partitions(grid)
    _, P_x, _ = d.create_standard_partitions(grid)
    torch.manual_seed(10 + rank)
    theirs = ref.DistributedFNO(P_ref, in_shape, nt, width, modes, num_blocks=blocks, dtype=torch.float64)
    ours = d.DistributedFNO(P_x, in_shape, nt, width, modes, num_blocks=blocks, dtype=torch.float64,
                            backend="torch", plan="reference")
    sd = theirs.state_dict()
    assert sorted(sd) == sorted(ours.state_dict()), "state-dict keys differ"
    for k, v in ours.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, tuple(v.shape), tuple(sd[k].shape))
    ours.load_state_dict(sd)
    info = d.compute_distribution_info(P_x, in_shape)
    theirs_info = ref.compute_distribution_info(P_ref, in_shape)
    assert tuple(info["shape"]) == tuple(theirs_info["shape"]) and tuple(info["start"]) == tuple(theirs_info["start"])
    g = torch.Generator().manual_seed(99)
    xg = torch.randn(*in_shape, dtype=torch.float64, generator=g)
    x = xg[tuple(info["slice"])].contiguous()
    y0, y1 = theirs(x.clone()), ours(x.clone())
    t = torch.randn(y0.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(5 + rank))
    l0 = ref.DistributedRelativeLpLoss(P_ref)(y0, t)
    l1 = d.DistributedRelativeLpLoss(P_x)(y1, t)
    l0.backward()
    l1.backward()
    g0 = {k: p.grad for k, p in theirs.named_parameters() if p.grad is not None}
    g1 = {k: p.grad for k, p in ours.named_parameters() if p.grad is not None}
    assert sorted(g0) == sorted(g1)
    gerr = max([float((g0[k] - g1[k]).abs().max()) for k in g0 if g0[k].numel()] or [0.0])
    # off the root the reference's loss is mean(empty / empty) = NaN (it only ever prints the root's value);
    # ours is a well-defined 0 there
    lerr = abs(float(l0) - float(l1)) if rank == 0 else float(l1)
    return float((y0 - y1).abs().max()), lerr, gerr


CFG = dict(in_shape=[1, 2, 8, 8, 8, 2], nt=4, width=3, modes=(2, 2, 2, 2), blocks=2)


@pytest.mark.parametrize("ws,grid", [(1, (1, 1, 1, 1, 1, 1)), (4, (1, 1, 2, 2, 1, 1)), (2, (1, 1, 1, 2, 1, 1))])
def test_unmodified_reference_matches_portable_backend(ws, grid):
    res = run_distributed(_parity, ws, grid, CFG["in_shape"], CFG["nt"], CFG["width"], CFG["modes"], CFG["blocks"],
                          timeout=600)
    for yerr, lerr, gerr in res:
        assert yerr < 1e-12 and lerr < 1e-12 and gerr < 1e-12, (yerr, lerr, gerr)
