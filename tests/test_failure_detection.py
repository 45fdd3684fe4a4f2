"""Fault injection on CPU (gloo): a peer that stops participating must not hang the job -- the
StepWatchdog aborts the stuck rank with exit code 75 (the reference hangs forever).
Plus the NaN-poisoning helpers used to catch unwritten tiles."""
import multiprocessing as mp
import os
import time

import pytest
import torch

from dfno_b200.utils.testing import free_port


def _rank(rank, port, stall):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    import torch.distributed as dist
    import dfno_b200 as d
    dist.init_process_group("gloo", rank=rank, world_size=2)
    dist.barrier()                                  # everybody is up
    if rank == 1 and stall:
        time.sleep(120)                             # injected fault: alive, but never reaches the collective
        os._exit(0)
    with d.StepWatchdog(3.0, what="barrier"):
        dist.barrier()
    os._exit(0)


@pytest.mark.parametrize("stall", [True, False])
def test_watchdog_aborts_a_rank_stuck_on_a_dead_peer(stall):
    ctx = mp.get_context("spawn")
    port = free_port()
    procs = [ctx.Process(target=_rank, args=(r, port, stall)) for r in range(2)]
    t0 = time.time()
    for p in procs:
        p.start()
    procs[0].join(60)
    took = time.time() - t0
    try:
        assert not procs[0].is_alive(), "rank 0 hung"
        if stall:
            assert procs[0].exitcode == 75, procs[0].exitcode     # aborted by the watchdog, loudly
            assert took < 45
        else:
            procs[1].join(30)
            assert procs[0].exitcode == 0 and procs[1].exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(10)


def test_poison_and_assert_finite():
    import dfno_b200 as d

    class Fake:
        world = 1
        ws = {"a": torch.zeros(8), "b": [torch.ones(4), torch.ones(2)], "idx": torch.zeros(3, dtype=torch.int64)}

    m = Fake()
    d.poison(m)
    assert torch.isnan(m.ws["a"]).all() and all(torch.isnan(t).all() for t in m.ws["b"])
    assert (m.ws["idx"] == 0).all()                 # integer buffers are left alone
    with pytest.raises(FloatingPointError):
        d.assert_finite(torch.ones(3), m.ws["a"], what="out")
    d.assert_finite(torch.ones(3), None)
    assert "racecheck" in d.sanitizer_command("racecheck")
