#!/usr/bin/env python
"""What bounds the quantizing K = 768 GEMMs?  The same launch with parts of the kernel switched off (NQ_GEMM_DBG bit mask:
1 no epilogue math, 2 no TMEM loads, 4 no stores, 8 no MMAs, 16 no epilogue chunks, 32 no TMA loads) -- results are garbage, times are not.
Kernel time = one CUDA graph of 20 launches between two events / 20."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import kernels as K
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, iters=20):
    """20 launches captured in one CUDA graph (the host cannot enqueue a 40 us kernel fast enough to time it otherwise)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for N, Kd, gelu in ((768, 768, False), (3072, 768, True)):
    a = torch.randint(-128, 128, (1, 50432, Kd), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
    azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
    bias = torch.randn(N, device=DEV)
    for dbg in (0, 1, 2, 4, 8, 16, 24, 32, 40, 41, 47, 56):
        os.environ["NQ_GEMM_DBG"] = str(dbg)
        if gelu:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, 50432, False, gelu=(1.4142135381698608, 1.0, 0.5)))
        else:
            t = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "split_rows", 12, 197, False))
        print(f"N={N} K={Kd} dbg={dbg:3d}: {t:.1f} us", flush=True)
os.environ.pop("NQ_GEMM_DBG", None)
