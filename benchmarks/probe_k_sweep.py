#!/usr/bin/env python
"""Is a quantizing GEMM bound by its main loop (operand feed, MMA) or by its epilogue?  Sweep K at fixed M, N: the
epilogue work is constant, operand bytes and MMA work scale with K.  (Per-call sync: compare the rows, not absolutes.)"""
import sys, torch
sys.path.insert(0, '/root/repo')
from numpy_quant_b200 import _lib, kernels as K
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, iters=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]
for N in (3072, 768):
    for Kd in (768, 512, 384, 256, 128):
        a = torch.randint(-128, 128, (1, 50432, Kd), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
        bias = torch.randn(N, device=DEV)
        if N == 3072:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, 50432, False, gelu=(1.4142135381698608, 1.0, 0.5)))
        else:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "split_rows", 12, 197, True))
        print(f"N={N} K={Kd}: {t*1e3:.1f} us", flush=True)
