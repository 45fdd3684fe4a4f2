"""Losses over a ``P_x``-decomposed prediction.

Both reduce per-rank partial sums onto rank 0 with :class:`SumReduce` (adjoint:
Broadcast), so the scalar is *valid on the root rank* and a differentiable ``0`` elsewhere
-- every rank can call ``loss.backward()``.

* ``DistributedRelativeLpLoss``: batch mean of ``||y^-y||_p / ||y||_p`` with the norms taken
  over the whole (global) sample -- ``/root/reference/dfno/loss.py:8-35``.
* ``DistributedMSELoss``: global mean squared error (DistDL module the reference's
  scripts use: ``experiment_navier_stokes.py:118``, ``dfno.py:374``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..parallel.partition import Partition, create_root_partition
from ..parallel.primitives import SumReduce, ZeroVolumeCorrectorFunction

__all__ = ["DistributedRelativeLpLoss", "DistributedMSELoss"]


def _acc_dtype(t: torch.Tensor) -> torch.dtype:
    return torch.float32 if t.dtype in (torch.bfloat16, torch.float16) else t.dtype


class DistributedRelativeLpLoss(nn.Module):
    def __init__(self, P_x: Partition, p: float = 2):
        super().__init__()
        self.P_x, self.p = P_x, p
        self.P_0 = create_root_partition(P_x)
        self.sr0 = SumReduce(P_x, self.P_0)
        self.sr1 = SumReduce(P_x, self.P_0)

    def forward(self, y_hat: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        B = y_hat.shape[0]
        acc = _acc_dtype(y_hat)
        d = (y_hat.to(acc) - y.to(acc)).reshape(B, -1)
        r = y.to(acc).reshape(B, -1)
        if self.p == 2:
            num, den = (d * d).sum(dim=1), (r * r).sum(dim=1)
        else:
            num, den = d.abs().pow(self.p).sum(dim=1), r.abs().pow(self.p).sum(dim=1)
        num, den = self.sr0(num), self.sr1(den)
        if self.P_0.active:
            out = (num.pow(1.0 / self.p) / den.pow(1.0 / self.p)).mean()
        else:
            out = num            # zero-volume; corrected below
        return ZeroVolumeCorrectorFunction.apply(out)


class DistributedMSELoss(nn.Module):
    def __init__(self, P_x: Partition):
        super().__init__()
        self.P_x = P_x
        self.P_0 = create_root_partition(P_x)
        self.sr = SumReduce(P_x, self.P_0)

    def forward(self, y_hat: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        acc = _acc_dtype(y_hat)
        d = y_hat.to(acc) - y.to(acc)
        part = torch.stack([(d * d).sum(), d.new_tensor(float(d.numel()))])
        tot = self.sr(part)
        out = tot[0] / tot[1] if self.P_0.active else tot
        return ZeroVolumeCorrectorFunction.apply(out)
