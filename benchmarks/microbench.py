#!/usr/bin/env python
"""Config 5 microbench: standalone qGEMM 4096^3 and the quantize / dequantize / requantize
kernels on 4096^2 (and the ViT-B b256 shapes), each timed alone with CUDA events after an L2
flush, reported against the measured peaks.  One JSON line per case on stdout.

    python benchmarks/microbench.py [--quick]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpy_quant_b200 import _lib, kernels as K  # noqa: E402

DEV = torch.device("cuda:0")
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
BF16 = PEAKS.get("bf16_tflops", 1590.0)
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.fill_(1)                                  # evict L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def gemm_case(M, N, Kd, batch=1, mode=_lib.EPI_RAW, label=""):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randint(-128, 128, (batch, M, Kd), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", True), K.operand_from_codes(b, "B", True)
    azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
    kw = dict(mode=mode, scale=1e-4, azp=azp)
    if mode == _lib.EPI_REQUANT:
        kw.update(out_bits=8, out_scale=0.05, out_zp=-3)
    med, best = timed(lambda: K.qgemm(oa, ob, **kw))
    ops = 2.0 * batch * M * N * Kd
    out_b = {0: 4, 1: 4, 2: 1}[mode]
    byts = batch * M * Kd + N * Kd + batch * M * N * out_b
    emit(case=f"qgemm {label or ''}".strip(), M=M, N=N, K=Kd, batch=batch, epilogue=["raw_s32", "dequant_f32", "requant_s8"][mode],
         ms_median=med, ms_best=best, tops=ops / med / 1e9, frac_of_2x_bf16_measured=ops / med / 1e9 / (2 * BF16),
         frac_of_nominal_4500=ops / med / 1e9 / 4500.0, hbm_gbs_implied=byts / med / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--gemm-only", action="store_true")
    args = ap.parse_args()
    emit(case="peaks", hbm_gbs=HBM, bf16_tflops=BF16, source="MEASURED_PEAKS.json" if PEAKS else "fallback")
    # library int8 GEMM (cuBLASLt through torch._int_mm) as a second, measured int8 reference point
    try:
        n = 8192
        a = torch.randint(-128, 128, (n, n), device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (n, n), device=DEV, dtype=torch.int8).t().contiguous().t()
        med, best = timed(lambda: torch._int_mm(a, b), iters=5)
        emit(case="library int8 gemm (torch._int_mm, cuBLASLt) 8192^3", ms_median=med, tops=2.0 * n ** 3 / med / 1e9)
        del a, b
    except Exception as e:  # noqa: BLE001
        emit(case="library int8 gemm", error=str(e)[:200])
    for mode in (_lib.EPI_RAW, _lib.EPI_DEQUANT, _lib.EPI_REQUANT):
        gemm_case(4096, 4096, 4096, mode=mode, label="4096^3")
    if not args.quick:
        gemm_case(8192, 8192, 8192, label="8192^3")
        for (M, N, Kd, lab) in ((50432, 768, 768, "ViT qkv/o"), (50432, 3072, 768, "ViT fc1"), (50432, 768, 3072, "ViT fc2")):
            for mode in (_lib.EPI_RAW, _lib.EPI_DEQUANT, _lib.EPI_REQUANT):
                gemm_case(M, N, Kd, mode=mode, label=lab)
        g = torch.Generator(device="cuda").manual_seed(0)
        for (bt, M, N, Kd, lab) in ((3072, 197, 197, 64, "ViT QK^T"), (3072, 197, 64, 197, "ViT PV")):
            a = torch.randint(-128, 128, (bt, M, Kd), generator=g, device=DEV, dtype=torch.int8)
            b = torch.randint(-128, 128, (bt, Kd, N), generator=g, device=DEV, dtype=torch.int8)
            oa, ob = K.operand_from_codes(a, "A", True), K.operand_from_codes(b, "B", True)
            azp = K.AccZeroPoint(3, -4, Kd, oa.rowsum, ob.rowsum, False)
            med, best = timed(lambda: K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp))
            ops = 2.0 * bt * M * N * Kd
            emit(case=f"qgemm {lab}", M=M, N=N, K=Kd, batch=bt, epilogue="dequant_f32", ms_median=med, tops=ops / med / 1e9,
                 hbm_gbs_implied=(bt * (M * Kd + N * Kd) + bt * M * N * 4) / med / 1e6)
    if not args.quick:
        # epilogue-fused attention pieces at ViT-B b256 shapes
        g = torch.Generator(device="cuda").manual_seed(1)
        bt, S, D = 3072, 197, 64
        q8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        k8 = torch.randint(-128, 128, (bt, D, S), generator=g, device=DEV, dtype=torch.int8)
        oq, ok_ = K.operand_from_codes(q8, "A", True), K.operand_from_codes(k8, "B", True)
        azp = K.AccZeroPoint(3, -4, D, oq.rowsum, ok_.rowsum, False)
        med, _ = timed(lambda: K.qgemm_softmax_to_operand(oq, ok_, 1e-4, azp, 8.0, 8, 1 / 255, -128, True))
        emit(case="qgemm ViT QK^T + softmax + quantize (epilogue)", M=S, N=S, K=D, batch=bt, ms_median=med,
             tops=2.0 * bt * S * S * D / med / 1e9)
        f = K.qgemm(oq, ok_, _lib.EPI_DEQUANT, 1e-4, azp)
        med2, _ = timed(lambda: K.softmax_quantize(f, 8.0, 8, 1 / 255, -128, True))
        emit(case="softmax+quantize kernel on stored scores", ms_median=med2)
        a8 = torch.randint(-128, 128, (1, 50432, 768), generator=g, device=DEV, dtype=torch.int8)
        w8 = torch.randint(-128, 128, (1, 768, 768), generator=g, device=DEV, dtype=torch.int8)
        oa, ow = K.operand_from_codes(a8, "A", False), K.operand_from_codes(w8, "B", True)
        azw = K.AccZeroPoint(3, None, 768, None, ow.rowsum, True)
        bias = torch.randn(768, device=DEV)
        for kind in ("split_rows", "split_cols"):
            med, _ = timed(lambda: K.qgemm_to_operand(oa, ow, 1e-4, azw, bias, 8, 0.05, -3, kind, 12, 197, True))
            emit(case=f"qgemm ViT qkv -> int8 operand ({kind})", ms_median=med, tops=2.0 * 50432 * 768 * 768 / med / 1e9)
        # fused attention kernel at the same shapes (B = 256, H = 12)
        kt8 = torch.randint(-128, 128, (bt, D, S), generator=g, device=DEV, dtype=torch.int8)
        vv8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        fq, fk, fv = K.operand_from_codes(q8, "A", True), K.operand_from_codes(kt8, "B", True), K.operand_from_codes(vv8, "B", True)
        for o in (fq, fk, fv):
            o.batch_shape = (bt // 12, 12)
        med, _ = timed(lambda: K.attention(fq, fk, fv, 1e-4, 3, -4, 8.0, 8, 1 / 255, -128, 1e-4, 9, 8, 0.05, -3, False))
        emit(case="fused attention ViT-B b256 (QK^T + softmax + P.V + merge heads, one kernel)", batch=bt, S=S, D=D, ms_median=med,
             tops=4.0 * bt * S * S * D / med / 1e9)
        p8 = torch.randint(-128, 128, (bt, S, S), generator=g, device=DEV, dtype=torch.int8)
        v8 = torch.randint(-128, 128, (bt, S, D), generator=g, device=DEV, dtype=torch.int8)
        op_, ov_ = K.operand_from_codes(p8, "A", True), K.operand_from_codes(v8, "B", True)
        azp2 = K.AccZeroPoint(-128, 9, S, op_.rowsum, ov_.rowsum, False)
        med, _ = timed(lambda: K.qgemm_to_operand(op_, ov_, 1e-4, azp2, None, 8, 0.05, -3, "merge_heads", 12, S, False))
        emit(case="qgemm ViT PV -> int8 operand (merge_heads)", ms_median=med, tops=2.0 * bt * S * S * D / med / 1e9)
    if not args.quick:
        # MLP-1: bias + GELU chain + quantize for MLP-2 in the epilogue
        a = torch.randint(-128, 128, (1, 50432, 768), generator=g, device=DEV, dtype=torch.int8)
        b = torch.randint(-128, 128, (1, 768, 3072), generator=g, device=DEV, dtype=torch.int8)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3, None, 768, None, ob.rowsum, True)
        bias = torch.randn(3072, device=DEV)
        med, _ = timed(lambda: K.qgemm_to_operand(oa, ob, 1e-4, azp, bias, 8, 0.05, -3, "rows", 1, 50432, False,
                                                  gelu=(1.4142135381698608, 1.0, 0.5)))
        emit(case="qgemm ViT MLP-1 + bias + GELU -> int8 operand", M=50432, N=3072, K=768, ms_median=med,
             tops=2.0 * 50432 * 3072 * 768 / med / 1e9)
        del a, b, oa, ob
        # output projection / MLP-2: dequantize + bias + residual add in the epilogue (float32 residual stream)
        for name, (M, N, Kd) in (("o-proj", (50432, 768, 768)), ("MLP-2", (50432, 768, 3072))):
            a = torch.randint(-128, 128, (1, M, Kd), generator=g, device=DEV, dtype=torch.int8)
            b = torch.randint(-128, 128, (1, Kd, N), generator=g, device=DEV, dtype=torch.int8)
            oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
            azp = K.AccZeroPoint(3, None, Kd, None, ob.rowsum, True)
            bias, res = torch.randn(N, device=DEV), torch.randn(M, N, device=DEV)
            med, _ = timed(lambda: K.qgemm(oa, ob, _lib.EPI_DEQUANT, 1e-4, azp, bias_f32=bias, residual=res))
            emit(case=f"qgemm ViT {name} + bias + residual (f32 out)", M=M, N=N, K=Kd, ms_median=med, tops=2.0 * M * N * Kd / med / 1e9)
            del a, b, oa, ob, res
        # BASELINE config 3: Conv2d block as im2col + int8 qGEMM, batch 1024 (test_conv2d geometry scaled up)
        Bc, Cc, Hc, Wc, Oc, kh, kw = 1024, 64, 57, 58, 128, 3, 2
        g = torch.Generator(device="cuda").manual_seed(3)
        xf = torch.randn((Bc, Cc, Hc, Wc), generator=g, device=DEV)
        w8 = torch.randint(-128, 128, (1, kh * kw * Cc, Oc), generator=g, device=DEV, dtype=torch.int8)
        ow_ = K.operand_from_codes(w8, "B", True)
        bias = torch.randn(Oc, device=DEV)
        med_q, _ = timed(lambda: K.quantize(xf, 8, 0.03, -5), iters=5)
        xq = K.quantize(xf, 8, 0.03, -5)
        med_i, _ = timed(lambda: K.im2col(xq, kh, kw, (0, 2, 2, 1), (2, 1), -5), iters=5)
        cols, oh, ow2 = K.im2col(xq, kh, kw, (0, 2, 2, 1), (2, 1), -5)
        oa = K.Operand(cols.view(1, cols.shape[0], cols.shape[1]), (), cols.shape[0], kh * kw * Cc, cols.shape[1], None)
        azc = K.AccZeroPoint(-5, None, kh * kw * Cc, None, ow_.rowsum, True)
        med_g, _ = timed(lambda: K.qgemm(oa, ow_, _lib.EPI_DEQUANT, 1e-4, azc, bias_f32=bias), iters=5)
        Mc = cols.shape[0]
        ops = 2.0 * Mc * Oc * kh * kw * Cc
        emit(case="conv block (config 3) b1024: quantize + im2col + qGEMM(dequant+bias)", M=Mc, N=Oc, K=kh * kw * Cc,
             out_hw=[oh, ow2], ms_quantize=med_q, ms_im2col=med_i, ms_qgemm=med_g, ms_total=med_q + med_i + med_g,
             images_per_s=Bc / ((med_q + med_i + med_g) * 1e-3), tops_qgemm=ops / med_g / 1e9,
             quantize_gbs=xf.numel() * 5 / med_q / 1e6, im2col_gbs=(xq.numel() + cols.numel()) / med_i / 1e6)
        del cols, oa
        # the same block as an implicit GEMM: NCHW -> padded NHWC relayout (1x1 im2col), then nq_qconv2d_s8
        pads_c = (0, 2, 2, 1)
        med_r, _ = timed(lambda: K.nhwc_pad(xq, pads_c, -5), iters=5)
        med_f, _ = timed(lambda: K.nhwc_pad(xf, pads_c, -5, quant=(8, 0.03, -5)), iters=5)
        nhwc = K.nhwc_pad(xf, pads_c, -5, quant=(8, 0.03, -5))
        wk = K.operand_from_codes(w8[0].t().contiguous(), "A", True)
        azi = K.AccZeroPoint(-5, None, kh * kw * Cc, None, wk.rowsum, True)
        med_c, _ = timed(lambda: K.qconv2d(nhwc, wk, kh, kw, (2, 1), _lib.EPI_DEQUANT, 1e-4, azi, bias_f32=bias), iters=5)
        emit(case="conv block (config 3) b1024, implicit GEMM: quantize->NHWC + qconv2d(dequant+bias)", M=Mc, N=Oc,
             K=kh * kw * Cc, ms_quantize_nhwc=med_f, ms_qconv=med_c, ms_total=med_f + med_c,
             images_per_s=Bc / ((med_f + med_c) * 1e-3), tops_qconv=ops / med_c / 1e9,
             quantize_nhwc_gbs=(4 * xf.numel() + nhwc.numel()) / med_f / 1e6,
             ms_relayout_codes=med_r, relayout_codes_gbs=(xq.numel() + nhwc.numel()) / med_r / 1e6)
        del xf, xq, nhwc
    if args.gemm_only:
        return
    # HBM-bound kernels: algorithmic bytes per element as fixed in SURVEY.md §8(d)
    shapes = [(4096, 4096)] if args.quick else [(4096, 4096), (50432, 768), (50432, 3072)]
    for shp in shapes:
        n = shp[0] * shp[1]
        x = torch.randn(shp, device=DEV)
        for asym in (True, False):
            zp = -7 if asym else None
            med, _ = timed(lambda: K.quantize(x, 8, 0.03, zp))
            emit(case="quantize f32->s8", shape=shp, asym=asym, ms=med, gbs=5.0 * n / med / 1e6, frac_hbm=5.0 * n / med / 1e6 / HBM,
                 gelem_s=n / med / 1e6)
        med, _ = timed(lambda: K.quantize_operand(x, "A", 8, 0.03, -7, True))
        emit(case="quantize f32->s8 operand A (+rowsum)", shape=shp, ms=med, gbs=5.0 * n / med / 1e6, frac_hbm=5.0 * n / med / 1e6 / HBM)
        q = K.quantize(x, 8, 0.03, -7)
        med, _ = timed(lambda: K.dequantize(q, 0.03, -7))
        emit(case="dequantize s8->f32", shape=shp, ms=med, gbs=5.0 * n / med / 1e6, frac_hbm=5.0 * n / med / 1e6 / HBM)
        acc = torch.randint(-2 ** 20, 2 ** 20, (1,) + shp, device=DEV, dtype=torch.int32)
        cs = torch.randint(-1000, 1000, (1, shp[1]), device=DEV, dtype=torch.int32)
        azp = K.AccZeroPoint(5, None, 768, None, cs, True)
        med, _ = timed(lambda: K.dequantize_acc(acc, 1e-4, azp))
        emit(case="dequantize s32 acc->f32 (factored zp)", shape=shp, ms=med, gbs=8.0 * n / med / 1e6, frac_hbm=8.0 * n / med / 1e6 / HBM)
        med, _ = timed(lambda: K.requantize_acc(acc, 1e-4, azp, None, 8, 0.05, -3))
        emit(case="requantize s32 acc->s8", shape=shp, ms=med, gbs=5.0 * n / med / 1e6, frac_hbm=5.0 * n / med / 1e6 / HBM)
        med, _ = timed(lambda: K.gelu_erf(x, 1.4142135, 1.0, 0.5))
        emit(case="gelu-erf chain f32", shape=shp, ms=med, gbs=8.0 * n / med / 1e6, frac_hbm=8.0 * n / med / 1e6 / HBM)
        gm = torch.ones(shp[1], device=DEV)
        med, _ = timed(lambda: K.layernorm(x, gm, gm, 1e-12))
        emit(case="layernorm f32", shape=shp, ms=med, gbs=8.0 * n / med / 1e6, frac_hbm=8.0 * n / med / 1e6 / HBM)
        if shp[1] <= 1024:
            for glue in (False, True):
                med, _ = timed(lambda: K.layernorm_quantize(x, gm, gm, 1e-12, 8, 0.03, -5, False, float_glue=glue))
                emit(case=f"layernorm -> quantize (int8 operand){', float glue' if glue else ''}", shape=shp, ms=med,
                     gbs=5.0 * n / med / 1e6, frac_hbm=5.0 * n / med / 1e6 / HBM)
        med, _ = timed(lambda: K.binary("add", x, gm))
        emit(case="bias add f32", shape=shp, ms=med, gbs=8.0 * n / med / 1e6, frac_hbm=8.0 * n / med / 1e6 / HBM)
        for bits in (4, 2):
            med, _ = timed(lambda: K.pack(q, bits))
            emit(case=f"pack s8->{bits}b", shape=shp, ms=med, gbs=(1 + bits / 8) * n / med / 1e6, frac_hbm=(1 + bits / 8) * n / med / 1e6 / HBM)
        del x, q, acc
    if not args.quick:
        x = torch.randn((256 * 12 * 197, 197), device=DEV)
        n = x.numel()
        med, _ = timed(lambda: K.softmax_lastdim(x))
        emit(case="softmax f32 rows of 197", shape=list(x.shape), ms=med, gbs=8.0 * n / med / 1e6, frac_hbm=8.0 * n / med / 1e6 / HBM)


if __name__ == "__main__":
    main()
