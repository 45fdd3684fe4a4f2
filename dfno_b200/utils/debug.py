"""Debugging aids for the peer-memory kernels (the reference has no race
detection and hangs forever when a rank dies).

* :func:`poison` -- fill the engine's symmetric / scratch buffers with NaNs before a step, so a
  tile that was never written (wrong scatter table, missed peer) shows up as a NaN in the output
  instead of as stale-but-plausible data; :func:`assert_finite` checks outputs and gradients.
* :func:`sanitizer_command` -- the ``compute-sanitizer`` invocations (memcheck / racecheck /
  synccheck) for the single-GPU kernel tests; see ``tools/sanitize.sh``.
* :class:`StepWatchdog` -- host-side failure detection: if a step does not finish within the
  deadline (a peer died inside a collective), the process aborts loudly instead of hanging the
  job.  The device-side counterpart is the bounded spin of ``p2p_barrier`` (it traps after
  ``DFNO_BARRIER_TIMEOUT_S`` seconds and records which peer was late).
"""
from __future__ import annotations

import os
import sys
import threading
from typing import Iterable

import torch

__all__ = ["poison", "assert_finite", "sanitizer_command", "StepWatchdog"]


def poison(model) -> None:
    """NaN-fill every scratch buffer of a fused model (not the saved activations)."""
    ws = getattr(model, "ws", None)
    if ws is None:
        return
    for name, buf in ws.items():
        bufs = buf if isinstance(buf, (list, tuple)) else [buf]
        for b in bufs:
            if torch.is_tensor(b) and b.is_floating_point():
                b.fill_(float("nan"))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if getattr(model, "world", 1) > 1:
        model.barrier()                       # nobody starts writing into a buffer still being poisoned


def assert_finite(*tensors: Iterable[torch.Tensor], what: str = "tensor") -> None:
    for i, t in enumerate(tensors):
        if t is not None and not bool(torch.isfinite(t).all()):
            bad = int((~torch.isfinite(t)).sum())
            raise FloatingPointError(f"{what}[{i}]: {bad} non-finite values (unwritten tile or numeric blow-up)")


def sanitizer_command(tool: str = "memcheck", target: str = "tests/test_dft_gemm_gpu.py") -> str:
    assert tool in ("memcheck", "racecheck", "synccheck", "initcheck")
    return (f"compute-sanitizer --tool {tool} --error-exitcode 1 --launch-timeout 120 "
            f"python -m pytest {target} -x -q -k 'rowmajor or scatter'")


class StepWatchdog:
    """``with StepWatchdog(120): trainer.step(...)`` -- abort the process if the block does not
    finish in time.  Exits with code 75 so launchers (torchrun) tear the whole job down."""

    def __init__(self, seconds: float, what: str = "training step"):
        self.seconds, self.what = seconds, what
        self._done = threading.Event()

    def _watch(self):
        if not self._done.wait(self.seconds):
            sys.stderr.write(f"[dfno_b200] {self.what} exceeded {self.seconds:.0f}s on rank "
                             f"{os.environ.get('RANK', '0')}: aborting the job (a peer is likely dead)\n")
            sys.stderr.flush()
            os._exit(75)

    def __enter__(self):
        self._done.clear()
        self._t = threading.Thread(target=self._watch, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._done.set()
        return False
