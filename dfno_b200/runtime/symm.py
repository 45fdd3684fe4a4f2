"""Symmetric (peer-mapped) device buffers: the NVLink data plane of the fused engine.

``SymmetricBuffer`` allocates ``nbytes`` on this rank's GPU, exchanges CUDA IPC handles over
the ``torch.distributed`` control plane and maps every peer's allocation, yielding
``ptrs[r]`` = a device pointer *valid on this GPU* to rank ``r``'s buffer.  Kernels store to /
load from those pointers directly (NVSwitch routes the traffic); see ``csrc/symm_mem.cpp``,
``csrc/p2p.cu`` and the peer-scatter epilogue of ``csrc/dft_gemm_sm100.cu``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ..ops import build

__all__ = ["SymmetricBuffer", "PeerBarrier"]


class SymmetricBuffer:
    def __init__(self, nbytes: int, group=None, rank: int = 0, world: int = 1, device: Optional[int] = None):
        self.nbytes = int((nbytes + 255) // 256 * 256)
        self.group, self.rank, self.world = group, rank, world
        self.device = torch.cuda.current_device() if device is None else device
        C = build.load()
        self._C = C
        self.local_ptr, handle = C.symm_alloc(self.nbytes, self.device)
        self.ptrs: List[int] = [0] * world
        self.ptrs[rank] = self.local_ptr
        self._opened: List[int] = []
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, (rank, bytes(handle)), group=group)
            for r, h in handles:
                if r != rank:
                    p = C.symm_open(h, self.device)
                    self.ptrs[r] = p
                    self._opened.append(p)
            dist.barrier(group=group)

    def view(self, shape: Sequence[int], dtype: torch.dtype, byte_offset: int = 0) -> torch.Tensor:
        """Tensor view of the *local* buffer."""
        return self._C.tensor_from_ptr(self.local_ptr + byte_offset, list(shape), dtype, self.device)

    def peer_ptrs(self, byte_offset: int = 0) -> List[int]:
        return [p + byte_offset for p in self.ptrs]

    def close(self) -> None:
        if getattr(self, "_C", None) is None:
            return
        for p in self._opened:
            self._C.symm_close(p)
        self._opened = []
        if self.local_ptr:
            self._C.symm_free(self.local_ptr)
            self.local_ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBarrier:
    """Device-side flag barrier across the ranks sharing a signal pad (``p2p.cu``)."""

    def __init__(self, group=None, rank: int = 0, world: int = 1):
        self.rank, self.world = rank, world
        self.epoch = 0
        self.pad = SymmetricBuffer(256, group, rank, world) if world > 1 else None

    def __call__(self) -> None:
        if self.world <= 1:
            return
        self.epoch += 1
        self.pad._C.p2p_barrier(self.pad.peer_ptrs(), self.rank, self.epoch)
