#!/usr/bin/env python
"""Per-tile timeline of CTA 0 of the quantizing GEMM (NQ_GEMM_TRACE: clock64 stamps written by the kernel's roles).
Events per tile: 0 producer starts the tile, 1 producer issued its last K block, 2 MMA thread got the accumulator
buffer, 3 MMA thread committed the accumulator, 4 epilogue warp 4 starts waiting, 5 epilogue warp 4 got the
accumulator, 6 / 7 epilogue warps 4 / 19 release it.  Times in SM cycles relative to tile 8's producer start."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_quant_b200 import kernels as K
DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
trace = torch.zeros(64 * 8 + 64, dtype=torch.int64, device=DEV)
for N, Kd, gelu in ((768, 768, False), (3072, 768, True)):
    a = torch.randint(-128, 128, (1, 50432, Kd), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
    azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
    bias = torch.randn(N, device=DEV)
    for dbg in ([0, 56] if len(sys.argv) < 2 else [int(v) for v in sys.argv[1:]]):
        os.environ["NQ_GEMM_DBG"] = str(dbg)
        os.environ["NQ_GEMM_TRACE"] = str(trace.data_ptr())
        for _ in range(3):
            trace.zero_()
            if gelu:
                K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, 50432, False, gelu=(1.4142135381698608, 1.0, 0.5))
            else:
                K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "split_rows", 12, 197, False)
            torch.cuda.synchronize()
        kbt = trace.cpu()[512:].view(16, 4)
        t = trace.cpu()[:512].view(64, 8)
        n_t = int((t[:, 4] != 0).sum())
        base = int(t[0, 0]) if not (dbg & 256) else int(t[0, :].min())
        if dbg & 256:
            print("(alternative trace: release stamps of warps 16 17 18 19 (h = 3) | 4 5 6 7 (h = 0), i.e. SMSP 0 1 2 3 each)")
        print(f"--- N={N} K={Kd} dbg={dbg}: {n_t} tiles on CTA 0; cycles since the first producer stamp")
        print("tile  prod0  prod1   mma2   mma3   epi4   epi5   epi6  epi7(w19)")
        for i in range(min(n_t, 12)):
            print(f"{i:3d} " + " ".join(f"{int(v) - base:6d}" if v else "     -" for v in t[i]))
        print("tile 5, per K block: mma got stage | mma committed | producer got slot | producer issued")
        for kb in range(6):
            print(f"  kb {kb}: " + " ".join(f"{int(v) - base:7d}" if v else "      -" for v in kbt[kb]))
        if n_t > 4:
            print("period (cycles / tile, tiles 2..n-1):", (int(t[n_t - 1, 6]) - int(t[2, 6])) / max(n_t - 3, 1))
os.environ.pop("NQ_GEMM_DBG", None)
os.environ.pop("NQ_GEMM_TRACE", None)
