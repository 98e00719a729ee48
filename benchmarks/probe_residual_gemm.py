#!/usr/bin/env python
"""The float32-output GEMMs of the ViT-B layer (output projection K = 768, MLP-2 K = 3072) with and without bias /
residual, and the int8-output Q projection for comparison; one CUDA graph of 20 launches each, us per launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import _lib, kernels as K
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
M, N = 50432, 768
for Kd in (768, 3072):
    a = torch.randint(-128, 128, (1, M, Kd), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
    azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    print(f"K={Kd} int32 out          : {timed(lambda: K.qgemm(oa, ob)):.1f} us")
    print(f"K={Kd} float32            : {timed(lambda: K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp)):.1f} us")
    print(f"K={Kd} float32 + bias     : {timed(lambda: K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias)):.1f} us")
    print(f"K={Kd} float32 + bias+res : {timed(lambda: K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias, residual=res)):.1f} us")
    print(f"K={Kd} int8 (QUANT rows)  : {timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, 'rows', 1, M, False)):.1f} us", flush=True)
