#!/usr/bin/env python
"""Where does the free-running GPU forward leave the oracle?  Stem + 1 layer + head of ViT-B/16 at batch 2 with the
oracle's parameters: node-by-node retained run vs the oracle's environment, in graph order (first 12 deviating values
and the worst ones), then the fused run's logits."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
from numpy_quant_b200.tensor import FTensor, QTensor  # noqa: E402
from oracle import ref_graph as rg  # noqa: E402
import test_gpu_vitb_parity as T  # noqa: E402

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c = T.Case(bits)
out_r = c.q([c.x])[0]
print("retained logits dev (steps): max %.3f mean %.3f" % ((np.abs(out_r - c.want) / c.step).max(), (np.abs(out_r - c.want) / c.step).mean()))
rows = []
for op in c.plan.graph.ops:
    for name in op.outs:
        ov, v = c.env.get(name), c.by.get(name)
        if v is None or v.data is None or ov is None:
            continue
        if isinstance(ov, rg.F) and isinstance(v.data, FTensor):
            got = v.data.data
            rng_ = max(float(np.abs(ov.a).max()), 1e-30)
            d = np.abs(got - ov.a)
            rows.append((name, op.kind, "F", float(d.max() / rng_), float(np.mean(d > 0)), tuple(ov.a.shape)))
        elif isinstance(ov, rg.Q) and isinstance(v.data, QTensor):
            d = np.abs(v.data.data - ov.a)
            rows.append((name, op.kind, "Q%d" % ov.bits, float(d.max()), float(np.mean(d > 0)), tuple(ov.a.shape)))
for r in rows:
    print("%-95s %-18s %-4s max %.3e  frac!=0 %.3e  %s" % r)
out_f = c.q([c.x], retain=False)[0]
print("fused logits dev (steps): max %.3f mean %.3f" % ((np.abs(out_f - c.want) / c.step).max(), (np.abs(out_f - c.want) / c.step).mean()))
for flag in ("fuse_attention", "fuse_softmax_epilogue", "fuse_gelu_epilogue", "fuse_layernorm_glue"):
    setattr(c.q, flag, False)
    o = c.q([c.x], retain=False)[0]
    print("fused with %s=False (cumulative): max %.3f mean %.3f" % (flag, (np.abs(o - c.want) / c.step).max(), (np.abs(o - c.want) / c.step).mean()))
