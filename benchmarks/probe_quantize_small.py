#!/usr/bin/env python
"""quantize f32 -> int8 at 4096^2 (84 MB of traffic): launch-bound regime; L2 flushed before every call, median of 30."""
import os, sys, subprocess, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import kernels as K
DEV = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ts = []
for _ in range(35):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); K.quantize(x, 8, 0.03, -5); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
med = float(np.median(ts[5:]))
print(json.dumps({"variant": os.environ.get("NQ_QUANT_VARIANT", "0"), "us": med, "gb_s": 5 * x.numel() / med / 1e3}))
