import os, sys, torch
sys.path.insert(0, '/root/repo')
from numpy_quant_b200 import kernels as K
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
for M in (50432, 50432 * 4):
  for Kd in (128, 256, 768, 1536):
    a = torch.randint(-128, 128, (1, M, Kd), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, Kd, 768), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
    azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
    bias = torch.randn(768, device=DEV)
    S = 197
    res = []
    for dbg in (0, 16, 24):
        os.environ["NQ_GEMM_DBG"] = str(dbg)
        res.append(timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "split_rows", 12, S, False)))
    print(f"M={M} K={Kd}: full {res[0]:.1f}  no-epilogue-chunks {res[1]:.1f}  no-chunks-no-MMA {res[2]:.1f} us", flush=True)
