#!/usr/bin/env python
"""Time head_bwd2 alone on the headline shape, with DFNO_HEAD_DBG experiment masks (kernel-design probe)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for m in (0, 1, 2, 3, 4, 7, 8, 16, 31):
        out = subprocess.run([sys.executable, __file__, str(m)], capture_output=True, text=True,
                             env=dict(os.environ, DFNO_HEAD_DBG=str(m)))
        print(out.stdout.strip() or out.stderr[-300:])
    sys.exit(0)
import torch
from dfno_b200.ops import build
C_ = build.load()
dev = "cuda"
B, X, Y, Z, T, C, H = 1, 128, 128, 128, 20, 20, 128
S = X * Y * T * Z
h = torch.randn(B * C, S, device=dev).to(torch.bfloat16)
w3a = (torch.randn(H, 64, device=dev) * 0.2).to(torch.bfloat16)
w3t = (torch.randn(32, H, device=dev) * 0.2).to(torch.float16)
W4 = torch.randn(H, device=dev) * 0.1
dy = torch.randn(B, 1, X, Y, Z, T, device=dev) * 1e-7
g = torch.empty(B * C, S, device=dev, dtype=torch.bfloat16)
gW3, gb3, gW4, gb4 = torch.zeros(H, C, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(1, device=dev)
ws = torch.zeros(1, device=dev, dtype=torch.int32)
R, SR = [Z, T, B * X * Y], [T, 1, Z * T]
def run():
    C_.head_bwd2(h, w3a, w3t, W4, dy, ws, g, gW3, gb3, gW4, gb4, B, C, S, R, SR)
for _ in range(3): run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): run()
e.record(); torch.cuda.synchronize()
print(f"dbg={sys.argv[1]:>3s}  head_bwd2 {s.elapsed_time(e) / 5:7.3f} ms")
