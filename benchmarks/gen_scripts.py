#!/usr/bin/env python
"""Generate weak-scaling run scripts (``{eval,grad}_weak_scaling_{spatial,temporal}_gpu.sh``)
and a driver ``submit_<system>.sh`` -- the role of ``/root/reference/benchmarks/gen_scripts.py``
(Summit/Perlmutter tables at ``:119-161``), re-targeted at one-process-per-GPU ``torchrun``
launches on NVSwitch boxes:

* ``b200``  : one 8 x B200 box, y-pencil partitions ``(1,1,1,N,1,1)``, N = 1, 2, 4, 8
* ``local`` : CPU/gloo development runs, N <= 4

"spatial" grows the y extent (and its modes) with N at fixed per-GPU size; "temporal" keeps
space fixed and grows ``nt`` and ``modes_t`` with N -- the reference's two scaling modes
(``gen_scripts.py:44-52``).  Zero-size shards are rejected at generation time (``:55-63``).
"""
import os
from argparse import ArgumentParser
from pathlib import Path

ap = ArgumentParser()
ap.add_argument("--system", default="b200", choices=["b200", "local"])
ap.add_argument("--max-workers", "-mw", type=int, default=-1)
ap.add_argument("--clean-old", "-co", action="store_true")
ap.add_argument("--out", type=Path, default=Path(os.path.dirname(os.path.abspath(__file__))))
args = ap.parse_args()

SYSTEMS = {
    # per-GPU local shape (X, Y, Z, T), per-GPU modes, device, dtype
    "b200": dict(shape=(128, 16, 128, 20), modes=(12, 2, 12, 10), device="cuda", dtype="bf16", counts=(1, 2, 4, 8)),
    "local": dict(shape=(16, 8, 16, 8), modes=(4, 2, 4, 4), device="cpu", dtype="fp32", counts=(1, 2, 4)),
}
cfg = SYSTEMS[args.system]
counts = [n for n in cfg["counts"] if args.max_workers < 0 or n <= args.max_workers]


def launcher(n):
    return (f"python -m torch.distributed.run --nnodes=1 --nproc-per-node {n} --master-addr 127.0.0.1 "
            f"--master-port $((29500 + RANDOM % 1000)) bench.py")


def point(n, mode):
    X, Y, Z, T = cfg["shape"]
    mx, my, mz, mt = cfg["modes"]
    part = (1, 1, 1, n, 1, 1)
    if mode == "spatial":
        shape, modes, nt = (1, 1, X, Y * n, Z, 1), (mx, my * n, mz, mt), T
    else:
        shape, modes, nt = (1, 1, X, Y * max(cfg["counts"]), Z, 1), (mx, my * max(cfg["counts"]), mz, mt * n), T * n
    if n > shape[3] or 2 * modes[1] > shape[3] or modes[3] > nt // 2 + 1 or (2 * modes[2]) % n:
        raise ValueError(f"invalid configuration {shape} / {modes} / {part}: a shard would be empty")
    return shape, modes, nt, part


def make(name, run_type, mode):
    lines = ["#!/bin/bash", "set -x", f"data_dir={name}",
             'if test "x$1" = x; then echo "Usage: $0 <numranks>"; exit 0; fi', "ranks=$1"]
    for n in counts:
        shape, modes, nt, part = point(n, mode)
        lines.append(f"[[ $ranks -eq '{n}' ]] && {launcher(n)} --input-shape {' '.join(map(str, shape))} "
                     f"--modes {' '.join(map(str, modes))} --partition_shape {' '.join(map(str, part))} --width 20 "
                     f"--num-timesteps {nt} --device {cfg['device']} --num-gpus {n} --dtype {cfg['dtype']} "
                     f"--benchmark-type {run_type} --output-dir $data_dir")
    path = args.out / f"{name}.sh"
    path.write_text("\n".join(lines) + "\n")
    os.chmod(path, 0o755)
    print(f"created script for {args.system}: {path.name}")
    return [f"./{path.name} {n}" for n in counts]


args.out.mkdir(parents=True, exist_ok=True)
if args.clean_old:
    for f in args.out.glob("*_weak_scaling_*_gpu.sh"):
        f.unlink()
jobs = []
for run_type in ("eval", "grad"):
    for mode in ("spatial", "temporal"):
        jobs += make(f"{run_type}_weak_scaling_{mode}_gpu", run_type, mode)
sub = args.out / f"submit_{args.system}.sh"
sub.write_text("#!/bin/bash\nset -x\ncd \"$(dirname \"$0\")\"\n" + "\n".join(jobs) + "\n")
os.chmod(sub, 0o755)
print(f"created batch submission script: {sub.name}")
