#!/usr/bin/env python
"""Accuracy of the device GELU (value and derivative) against torch's erf form in float64, for the
extension variant selected by the environment (default A&S erf; ``DFNO_GELU_TANH3=1``: fitted tanh
form).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfno_b200.ops import build

x = torch.linspace(-9, 9, 400001, device="cuda")
y, dy = build.load().gelu_probe(x)
xd = x.double().requires_grad_()
ref = torch.nn.functional.gelu(xd)
ref.sum().backward()
print(json.dumps({"variant": "tanh3" if os.environ.get("DFNO_GELU_TANH3", "0") != "0" else "erf_as7126",
                  "build_dir": build.BUILD_DIR,
                  "max_abs_err_value": float((y.double() - ref.detach()).abs().max()),
                  "max_abs_err_grad": float((dy.double() - xd.grad).abs().max())}))
