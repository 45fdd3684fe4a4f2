"""Small helpers shared by the layers: complex-safe wire views and a one-shot metadata exchange."""
import torch
import torch.distributed as dist

from ..backend.backend import _on, comm_device


def wire(t):
    """Contiguous real view of ``t`` (NCCL and gloo have no complex collectives)."""
    t = t.contiguous()
    return torch.view_as_real(t) if t.is_complex() else t


def unwire(buf, like_complex):
    return torch.view_as_complex(buf) if like_complex else buf


def tell_everyone(obj, src):
    """Python object from world rank ``src`` to all ranks (used once per layer, on its first call)."""
    if not _on():
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src, device=comm_device())
    return box[0]


def collect_from_everyone(obj):
    if not _on():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
