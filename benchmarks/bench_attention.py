#!/usr/bin/env python
"""Fused attention kernel alone at the ViT-B/16 batch-256 shape (3072 heads of 197 x 64), CUDA events after an L2
flush; one JSON line per parameter variant (the constant-operand passes depend on the zero-points).

    python benchmarks/bench_attention.py [--heads 3072]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpy_quant_b200 import kernels as K  # noqa: E402

DEV = torch.device("cuda:0")
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=3072)
    args = ap.parse_args()
    bt, S, D = args.heads, 197, 64
    g = torch.Generator(device="cuda").manual_seed(1)
    q8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
    kt8 = torch.randint(-128, 128, (bt, D, S), generator=g, device=DEV, dtype=torch.int8)
    v8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
    fq, fk, fv = K.operand_from_codes(q8, "A", False), K.operand_from_codes(kt8, "B", False), K.operand_from_codes(v8, "B", False)
    for o in (fq, fk, fv):
        o.batch_shape = (bt // 12, 12)
    for label, zq, zk, zv, zp_p, sp in (("typical (zp_p = lo: no P constant pass)", 3, -4, 9, -128, 1 / 255),
                                         ("P zero-point below the range (bench.py's synthetic weights)", 3, -4, 9, -137, 1 / 2550),
                                         ("symmetric everything", None, None, None, None, 1 / 127),
                                         ("all constant passes doubled", -128, 5, -128, 100, 1 / 255)):
        med, best = timed(lambda: K.attention(fq, fk, fv, 1e-4, zq, zk, 8.0, 8, sp, zp_p, sp * 0.02, zv, 8, 0.05, -3, False))
        print(json.dumps(dict(case="fused attention " + label, heads=bt, S=S, D=D, ms_median=med, ms_best=best,
                              tops=4.0 * bt * S * S * D / med / 1e9, us_per_head=1e3 * med / bt)), flush=True)


if __name__ == "__main__":
    main()
