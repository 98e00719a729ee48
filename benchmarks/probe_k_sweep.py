#!/usr/bin/env python
"""Is a quantizing GEMM bound by its main loop (operand feed, MMA) or by its epilogue?  Sweep K at fixed M, N: the
epilogue work is constant, operand bytes and MMA work scale with K.  Kernel time = 20 back-to-back launches between two
CUDA events / 20 (the round-1 version timed single launches with a sync each and measured the host's launch overhead:
its flat ~85 us floor was the Python -> ctypes call, not the kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import _lib, kernels as K
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for N in (3072, 768):
    for Kd in (3072, 1536, 768, 512, 384, 256, 128):
        a = torch.randint(-128, 128, (1, 50432, Kd), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
        bias = torch.randn(N, device=DEV)
        if N == 3072:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, 50432, False, gelu=(1.4142135381698608, 1.0, 0.5)))
        else:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "split_rows", 12, 197, False))
        raw = timed(lambda: K.qgemm(oa, ob))
        print(f"N={N} K={Kd}: quantizing epilogue {t:.1f} us, raw int32 epilogue {raw:.1f} us", flush=True)
