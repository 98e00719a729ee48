#!/usr/bin/env python
"""Tiny driver for `ncu --set full` captures of the tensor-core GEMM (one launch per case)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpy_quant_b200 import _lib, kernels as K  # noqa: E402

DEV = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
cases = sys.argv[1:] or ["qkv", "fc1", "qk", "pv", "raw4096"]
for case in cases:
    if case in ("qkv", "fc1", "fc2"):
        M, N, Kd = {"qkv": (50432, 768, 768), "fc1": (50432, 3072, 768), "fc2": (50432, 768, 3072)}[case]
        a = torch.randint(-128, 128, (1, M, Kd), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
        bias = torch.randn(N, device=DEV)
        for _ in range(2):
            out = K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias)
    elif case in ("qk", "pv"):
        bt, M, N, Kd = (3072, 197, 197, 64) if case == "qk" else (3072, 197, 64, 197)
        a = torch.randint(-128, 128, (bt, M, Kd), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (bt, Kd, N), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", True), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, -4, Kd, oa.rowsum, ob.rowsum, False)
        for _ in range(2):
            out = K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp)
    elif case in ("qkv_quant", "v_quant", "o_res", "fc2_res", "fc1_gelu"):
        M, N, Kd = (50432, 3072, 768) if case == "fc1_gelu" else (50432, 768, 3072) if case == "fc2_res" else (50432, 768, 768)
        a = torch.randint(-128, 128, (1, M, Kd), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
        bias = torch.randn(N, device=DEV)
        res = torch.randn(M, N, device=DEV)
        for _ in range(2):
            if case in ("o_res", "fc2_res"):
                out = K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias, residual=res)
            elif case == "fc1_gelu":
                out = K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, M, False,
                                         gelu=(1.4142135381698608, 1.0, 0.5)).data
            else:
                out = K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3,
                                         "split_rows" if case == "qkv_quant" else "split_cols", 12, 197, True).data
    elif case == "attention":
        bt, S, D = 3072, 197, 64
        q8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        k8 = torch.randint(-128, 128, (bt, D, S), generator=g, device=DEV, dtype=torch.int8)
        v8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        fq, fk, fv = K.operand_from_codes(q8, "A", True), K.operand_from_codes(k8, "B", True), K.operand_from_codes(v8, "B", True)
        for o in (fq, fk, fv):
            o.batch_shape = (bt // 12, 12)
        for _ in range(2):
            out = K.attention(fq, fk, fv, 1e-4, 3, -4, 8.0, 8, 1 / 255, -128, 1e-4, 9, 8, 0.05, -3, False).data
    elif case == "ln":
        x = torch.randn((50432, 768), generator=g, device=DEV)
        gm, bt = torch.randn(768, device=DEV), torch.randn(768, device=DEV)
        for _ in range(3):
            out = K.layernorm_quantize(x, gm, bt, 1e-12, 8, 0.03, -5, False, float_glue=True).data
    elif case == "conv":
        # BASELINE config 3 as an implicit GEMM: quantize -> padded NHWC, then nq_qconv2d_s8
        Bc, Cc, Hc, Wc, Oc, kh, kw = 1024, 64, 57, 58, 128, 3, 2
        xf = torch.randn((Bc, Cc, Hc, Wc), generator=g, device=DEV)
        w8 = torch.randint(-128, 128, (Oc, kh * kw * Cc), generator=g, device=DEV, dtype=torch.int8)
        wk = K.operand_from_codes(w8, "A", True)
        azp = K.AccZeroPoint(-5, None, kh * kw * Cc, None, wk.rowsum, True)
        bias = torch.randn(Oc, device=DEV)
        for _ in range(2):
            nhwc = K.nhwc_pad(xf, (0, 2, 2, 1), -5, quant=(8, 0.03, -5))
            out, _, _ = K.qconv2d(nhwc, wk, kh, kw, (2, 1), _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias)
    elif case in ("qk_softmax", "pv_merge"):
        bt, S, D = 3072, 197, 64
        if case == "qk_softmax":
            a = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
            b = torch.randint(-128, 128, (bt, D, S), generator=g, device=DEV, dtype=torch.int8)
        else:
            a = torch.randint(-128, 128, (bt, S, S), generator=g, device=DEV, dtype=torch.int8)
            b = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", True), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, -4, a.shape[-1], oa.rowsum, ob.rowsum, False)
        for _ in range(2):
            if case == "qk_softmax":
                out = K.qgemm_softmax_to_operand(oa, ob, 1e-4, azp, 8.0, 8, 1 / 255, -128, True).data
            else:
                out = K.qgemm_to_operand(oa, ob, 1e-4, azp, None, 8, 0.05, -3, "merge_heads", 12, S, False).data
    else:
        a = torch.randint(-128, 128, (1, 4096, 4096), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(a, "B", False)
        for _ in range(2):
            out = K.qgemm(oa, ob)
    torch.cuda.synchronize()
    print(case, "ok", tuple(out.shape))
