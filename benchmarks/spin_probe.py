#!/usr/bin/env python
"""(Needs an extension built with -DDFNO_SPIN_PROBE, e.g. DFNO_EXTRA_NVCC_FLAGS=-DDFNO_SPIN_PROBE; the product build
ignores the switches and this script then only times the shipped kernel.)

spectral_in in isolation at the headline shape, with the probe switches of csrc/spectral_in_sm100.cu
(DFNO_SPIN_DBG bits: 1 = no epi-1 body, 2 = no epi-2 body, 4 = no TMA stores, 8 = no proxy fences; DFNO_SPIN_E,
DFNO_SPIN_ST = epilogue groups / ring stages), against the two dft_gemm launches it replaces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfno_b200.ops import build, operators as OPS
from dfno_b200.ops.gemm import pad_operator

C_ = build.load()
dev = torch.device("cuda", 0)
BC, X, Yl, T, Z, mz, mt = 20, 128, 128, 20, 128, 12, 10
KZ = 2 * mz
p1, p2 = pad_operator(OPS.fwd_real_to_complex(Z, mz), device=dev), pad_operator(OPS.fwd_complex(T, mt, False), device=dev)
h = torch.randn(BC, X, Yl, T, Z, device=dev).to(torch.bfloat16)
dst = torch.empty(BC * KZ * mt * X * Yl * 2, device=dev, dtype=torch.bfloat16)
dstr = [Yl * 2, X * Yl * 2, mt * X * Yl * 2, KZ * mt * X * Yl * 2]
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def run(tag, **env):
    for k in ("DFNO_SPIN_DBG", "DFNO_SPIN_E", "DFNO_SPIN_ST"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        ts = []
        for i in range(6):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            C_.spectral_in(h, p1, p2, [dst.data_ptr()], 0, dstr, BC, X, Yl, T, Z, KZ, mt)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(f"{tag:44s} {min(ts[1:]) * 1e3:8.1f} us", flush=True)
    except Exception as ex:      # noqa: BLE001
        print(f"{tag:44s} failed: {str(ex)[:80]}", flush=True)


run("default")
run("E=2", DFNO_SPIN_E=2)
run("no epi bodies", DFNO_SPIN_DBG=3)
run("no TMA loads", DFNO_SPIN_DBG=16)
