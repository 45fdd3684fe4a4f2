"""In-process "cluster" for tests: spawn ``world_size`` ranks on this host (gloo on CPU,
NCCL when ``cuda=True``), run ``fn(rank, world_size, *args)`` in each, and return the
per-rank results.  The reference can only be tested under ``mpirun``."""
from __future__ import annotations

import os
import socket
import tempfile
import traceback
from typing import Any, Callable, List

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

__all__ = ["run_distributed", "free_port"]


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world_size: int, port: int, cuda: bool, fn: Callable, args, outdir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world_size))
    result: Any
    try:
        if cuda:
            torch.cuda.set_device(rank % torch.cuda.device_count())
            dist.init_process_group("nccl", rank=rank, world_size=world_size,
                                    device_id=torch.device("cuda", rank % torch.cuda.device_count()))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world_size)
        result = ("ok", fn(rank, world_size, *args))
    except BaseException:                                   # noqa: BLE001 - report, never hang
        result = ("error", traceback.format_exc())
    torch.save(result, os.path.join(outdir, f"r{rank}.pt"))
    try:
        if dist.is_initialized():
            if result[0] == "ok":
                dist.barrier()
            dist.destroy_process_group()
    except BaseException:                                   # noqa: BLE001
        pass


def run_distributed(fn: Callable, world_size: int, *args, cuda: bool = False, timeout: float = 240.0) -> List[Any]:
    """Run ``fn`` on ``world_size`` spawned ranks; raises if any rank failed."""
    port = free_port()
    with tempfile.TemporaryDirectory() as outdir:
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_worker, args=(r, world_size, port, cuda, fn, args, outdir))
                 for r in range(world_size)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout)
        hung = [p for p in procs if p.is_alive()]
        for p in hung:
            p.terminate()
        results = []
        for r in range(world_size):
            path = os.path.join(outdir, f"r{r}.pt")
            results.append(torch.load(path, weights_only=False) if os.path.exists(path)
                           else ("error", "rank produced no result (crashed or hung)"))
    errs = [f"[rank {r}] {res[1]}" for r, res in enumerate(results) if res[0] != "ok"]
    if errs or hung:
        raise RuntimeError("distributed run failed:\n" + "\n".join(errs))
    return [res[1] for res in results]
