#!/usr/bin/env python
"""One warm-up step + one step of the headline config (used under ncu; never a benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dfno_b200 as d
G = int(os.environ.get("G", 128)); T = 20
dev = torch.device("cuda", 0)
_, P_x, _ = d.create_standard_partitions((1, 1, 1, 1, 1, 1))
net = d.DistributedFNO(P_x, [1, 1, G, G, G, 1], T, 20, (12, 12, 12, 10), device=dev, dtype=torch.bfloat16, backend="fused")
opt = d.FusedAdam(net)
crit = d.DistributedRelativeLpLoss(P_x)
x = torch.randn(1, 1, G, G, G, 1, device=dev); y = torch.randn(1, 1, G, G, G, T, device=dev)
for _ in range(int(os.environ.get("STEPS", 2))):
    opt.zero_grad(); loss = crit(net(x), y); loss.backward(); opt.step()
torch.cuda.synchronize()
print("done", float(loss))
