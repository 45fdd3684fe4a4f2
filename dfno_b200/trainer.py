"""Training-step driver: host batch -> device (pinned, asynchronous, double buffered) ->
forward -> distributed loss -> backward -> optimizer, loss read back to the host.

This is the loop of ``/root/reference/training/two_phase/train_two_phase.py:99-121`` as a
reusable object.  The H2D copies of step ``i+1`` run on a side stream while step ``i``
computes; the loss value is read back from a pinned scalar.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

__all__ = ["Trainer"]


class Trainer:
    def __init__(self, model, criterion, optimizer, device: Optional[torch.device] = None,
                 target_dtype: Optional[torch.dtype] = None):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.cuda = self.device.type == "cuda"
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._slots = [None, None]
        self._ready = [None, None]
        self._next = 0
        self._pending = None
        self.target_dtype = target_dtype
        self._loss_host = torch.zeros((), dtype=torch.float32).pin_memory() if self.cuda else torch.zeros(())
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -------------------------------------------------------------------- data movement
    def _upload(self, x_host: torch.Tensor, y_host: torch.Tensor):
        if not self.cuda:
            return x_host, y_host, None
        slot = self._next
        self._next ^= 1
        if self._slots[slot] is None or self._slots[slot][0].shape != x_host.shape \
                or self._slots[slot][1].shape != y_host.shape:
            self._slots[slot] = (torch.empty_like(x_host, device=self.device),
                                 torch.empty_like(y_host, device=self.device))
        xd, yd = self._slots[slot]
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)          # the slot's previous consumer has been enqueued
        with torch.cuda.stream(self.copy_stream):
            xd.copy_(x_host, non_blocking=True)
            yd.copy_(y_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.h2d_bytes = x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
        return xd, yd, ev

    def prefetch(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        """Start copying the *next* batch while the current step runs."""
        self._pending = self._upload(x_host, y_host)

    # -------------------------------------------------------------------- one step
    def step(self, x_host: torch.Tensor, y_host: torch.Tensor,
             next_batch: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> float:
        """One optimisation step on a host batch (this rank's shard).  Returns the loss
        (valid on the root rank, 0 elsewhere -- the distributed-loss convention)."""
        if self._pending is not None:
            xd, yd, ev = self._pending
            self._pending = None
        else:
            xd, yd, ev = self._upload(x_host, y_host)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        if next_batch is not None:
            self.prefetch(*next_batch)
        self.optimizer.zero_grad(set_to_none=True)
        y_hat = self.model(xd)
        loss = self.criterion(y_hat, yd)
        loss.backward()
        self.optimizer.step()
        if self.cuda:
            self._loss_host.copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self.d2h_bytes = 4
            return float(self._loss_host)
        self.d2h_bytes = 0
        return float(loss.detach())

    @torch.no_grad()
    def evaluate(self, x_host: torch.Tensor, y_host: torch.Tensor) -> float:
        xd, yd, ev = self._upload(x_host, y_host)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return float(self.criterion(self.model(xd), yd))
