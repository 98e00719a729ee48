#!/usr/bin/env python
"""Per-entry-point time breakdown of one quantized ViT-B/16 forward (eager interpreter, b256).

Every C-ABI call is bracketed with CUDA events on the launching stream and grouped by
(entry point, epilogue mode, M, N, K, batch).  Diagnostic only: event pairs add launch gaps, so
compare shares; bench.py holds the numbers of record.

    python benchmarks/step_breakdown.py [--batch 256] [--bits 8] [--steps 3]
"""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpy_quant_b200 import kernels as K, zoo  # noqa: E402
from numpy_quant_b200.model import Model  # noqa: E402

VIT = dict(image_size=224, patch_size=16, hidden=768, heads=12, intermediate=3072, layers=12, classes=1000)
MODES = {0: "raw", 1: "dequant", 2: "requant", 3: "quant", 4: "softmax_quant", 5: "gelu_quant"}  # nq_attention_s8 is listed by name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--layers", type=int, default=12)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = dict(VIT, layers=args.layers)
    proto = zoo.vit_graph(batch=args.batch, seed=0, **cfg)
    model = Model.from_onnx(proto)
    x = torch.from_numpy(np.random.default_rng(1).normal(size=(args.batch, 3, 224, 224)).astype(np.float32)).to(dev)
    qmodel = model.quantize([x], bit_width=args.bits)
    model.release()
    qmodel.release()
    for _ in range(2):
        qmodel([x], retain=False, device_outputs=True)
    torch.cuda.synchronize()

    records = []
    orig = K.call

    def timed_call(name, *a):
        key = name
        if name == "nq_qgemm_s8":
            ep = a[-2]._obj
            key = f"nq_qgemm_s8[{MODES.get(ep.mode, ep.mode)}{'+res' if ep.residual else ''}] M={a[3]} N={a[4]} K={a[5]} b={a[6]}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((key, e0, e1))

    K.call = timed_call
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        qmodel([x], retain=False, device_outputs=True)
    s1.record()
    torch.cuda.synchronize()
    K.call = orig
    total = s0.elapsed_time(s1) / args.steps
    agg = collections.OrderedDict()
    for key, e0, e1 in records:
        d = agg.setdefault(key, [0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    print(json.dumps({"case": "step", "batch": args.batch, "bits": args.bits, "ms_per_step_with_events": total,
                      "sum_bracketed_ms": sum(v[1] for v in agg.values()) / args.steps}))
    for key, (n, ms) in rows:
        print(json.dumps({"kernel": key, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                          "us_per_launch": 1e3 * ms / n, "share": ms / args.steps / total}))


if __name__ == "__main__":
    main()
