"""General (non-pencil) partitions on the fused engine: x/z-split and time-partitioned ``P_x`` are folded
onto the engine's y-pencil once at entry / exit (BASELINE configs 3 and 4 in miniature).  Sorted after the
kernel tests on purpose: these paths are the newest."""
import pytest
import torch

from dfno_b200.utils.testing import run_distributed
from test_fused_multigpu import CFG, _worker, _world

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


@pytest.mark.parametrize("grid2,grid4", [((1, 1, 2, 1, 1, 1), (1, 1, 2, 1, 2, 1)), ((1, 1, 1, 1, 1, 2), (1, 1, 1, 2, 1, 2))])
def test_general_partition_is_folded_onto_the_pencil(grid2, grid4):
    n = _world()
    for r in run_distributed(_worker, n, CFG, True, False, grid4 if n == 4 else grid2, cuda=True, timeout=300):
        assert r["fwd"] < 5e-2 and r["grad"] < 1e-1, r
        assert r.get("loss", 0) < 5e-2, r
        assert r["replica_drift"] == 0.0, r
