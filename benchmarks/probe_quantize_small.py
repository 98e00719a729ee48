#!/usr/bin/env python
"""Streaming kernels at 4096^2 (the launch-bound regime of BASELINE config 5): quantize f32 -> int8, accumulator
dequantize / requantize; L2 flushed before every call, median of 30 event-timed calls."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import kernels as K
DEV = torch.device("cuda:0")
n = 4096
x = torch.randn(n, n, device=DEV)
acc = torch.randint(-2_000_000, 2_000_000, (1, n, n), device=DEV, dtype=torch.int32)
rs = torch.randint(-9000, 9000, (1, n), device=DEV, dtype=torch.int32)
cs = torch.randint(-9000, 9000, (1, n), device=DEV, dtype=torch.int32)
azp = K.AccZeroPoint(3, -4, n, rs, cs, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def med(fn):
    ts = []
    for _ in range(35):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[5:]))
for name, fn, nbytes in (("quantize_f32_to_s8", lambda: K.quantize(x, 8, 0.03, -5), 5 * n * n),
                         ("dequantize_acc", lambda: K.dequantize_acc(acc, 1e-4, azp), 8 * n * n),
                         ("requantize_acc", lambda: K.requantize_acc(acc, 1e-4, azp, None, 8, 0.05, -3), 5 * n * n)):
    us = med(fn)
    print(json.dumps({"kernel": name, "us": us, "gb_s": nbytes / us / 1e3}))
