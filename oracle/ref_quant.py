"""ORACLE (test infrastructure, not product code) -- CPU restatement of the
reference's per-tensor affine quantization arithmetic.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package.  The product
(`numpy_quant_b200`) never does; it fails loudly if its CUDA library is missing.

Each function restates one routine of `/root/reference/numpy_quant/
numpy_quantization.py` (cited per function) under the NumPy >= 2 (NEP 50)
promotion rules that the reference actually runs under in the build container
(NumPy 2.3.5), with every dtype conversion written out explicitly instead of
being left to ufunc promotion.  Parity is PINNED: `tests/golden/make_golden.py`
executes the unmodified reference on seeded inputs and
`tests/test_oracle_golden.py` checks these restatements against the committed
vectors bit for bit, plus the known-answer vectors KA-1/KA-2 of SURVEY.md §8c.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64
I64 = np.int64


def qrange(bit_width: int) -> tuple[float, float]:
    """[lo, hi] of a signed `bit_width`-bit integer, as Python floats (`numpy_quantization.py:8`)."""
    return -(2.0 ** (bit_width - 1)), 2.0 ** (bit_width - 1) - 1.0


def quant_parameters(min_val, max_val, bit_width: int, asymmetric: bool):
    """scale / zero-point from a calibrated range (`numpy_quantization.py:7-21`).

    All arithmetic happens in the dtype of `min_val`/`max_val` (float32 for
    FTensor statistics; the Python-float range constants are "weak").  The
    zero-point is NOT clamped to the integer range.  Returns `(scale, zp)` with
    `scale` a 0-d float32 array and `zp` None (symmetric), an int64 scalar (when
    it rounds to 0 -- the reference's `zp and np.array(zp)` idiom) or a 0-d int64
    array.
    """
    lo, hi = qrange(bit_width)
    span = hi - lo
    with np.errstate(all="ignore"):
        if asymmetric:
            scale = (max_val - min_val) / span
            zp = np.rint(lo - min_val / scale).astype(I64)
            zp_out = np.array(zp, dtype=I64) if zp else zp
        else:
            scale = (2 * max(max_val, min_val)) / span
            zp_out = None
    return np.array(scale, dtype=F32), zp_out


def quantize(x: np.ndarray, bit_width: int, scale, zero_point) -> np.ndarray:
    """float32 -> int64 codes (`numpy_quantization.py:24-34`).

    t = x / scale in float32 (IEEE division); asymmetric adds the int64
    zero-point, which promotes to float64; clip, then round-half-even.
    """
    lo, hi = qrange(bit_width)
    with np.errstate(all="ignore"):
        t = np.asarray(x, dtype=F32) / F32(scale)
        if zero_point is not None:
            t = t.astype(F64) + F64(I64(zero_point))
        return np.rint(np.clip(t, lo, hi)).astype(I64)


def dequantize(q: np.ndarray, scale, zero_point) -> np.ndarray:
    """int64 codes -> float32 (`numpy_quantization.py:37-41`, `tensor.py:189-193`).

    (q - zp) stays int64; the product with the float32 scale is formed in
    float64 and rounded once to float32.
    """
    q = np.asarray(q, dtype=I64)
    if zero_point is not None:
        q = q - np.asarray(zero_point, dtype=I64)
    return (q.astype(F64) * F64(F32(scale))).astype(F32)


def q_matmul(a: np.ndarray, scale_a, zp_a, b: np.ndarray, scale_b, zp_b):
    """Exact integer contraction + zero-point bookkeeping (`numpy_quantization.py:44-61`).

    Returns (acc int64, scale float32, zp int64 array | None) where the real
    value is `scale * (acc - zp)`.
    """
    a = np.asarray(a, dtype=I64)
    b = np.asarray(b, dtype=I64)
    acc = np.matmul(a, b)
    scale = F32(scale_a) * F32(scale_b)
    if zp_a is None and zp_b is None:
        return acc, scale, None
    k = a.shape[-1]
    row = a.sum(axis=-1, keepdims=True)
    col = b.sum(axis=-2, keepdims=True)
    if zp_a is None:
        return acc, scale, row * I64(zp_b)
    if zp_b is None:
        return acc, scale, col * I64(zp_a)
    return acc, scale, row * I64(zp_b) + col * I64(zp_a) - I64(zp_a) * I64(zp_b) * k


def requantize(acc: np.ndarray, acc_scale, acc_zp, out_scale, out_zp, bit_width: int) -> np.ndarray:
    """wide accumulator -> `bit_width`-bit codes (`numpy_quantization.py:64-72`).

    d = dequantize(acc) (float32); t = (1/out_scale) * d in float32; asymmetric
    adds the int64 zero-point in float64; round-half-even, then clip.
    """
    lo, hi = qrange(bit_width)
    with np.errstate(all="ignore"):
        d = dequantize(acc, acc_scale, acc_zp)
        t = (F32(1) / F32(out_scale)) * d
        if out_zp is not None:
            t = t.astype(F64) + F64(I64(out_zp))
        return np.clip(np.rint(t), lo, hi).astype(I64)


def tensor_min_max(x: np.ndarray):
    """min/max widened to include 0 (`tensor.py:232-236`)."""
    zero = F32(0.0)
    return np.minimum(x.min(), zero), np.maximum(x.max(), zero)


def erf_poly(x: np.ndarray) -> np.ndarray:
    """Abramowitz & Stegun 7.1.26 in float32, op order of `numpy_helper.py:95-112`."""
    x = np.asarray(x, dtype=F32)
    sgn = np.sign(x)
    ax = np.abs(x)
    a1, a2, a3, a4, a5 = F32(0.254829592), F32(-0.284496736), F32(1.421413741), F32(-1.453152027), F32(1.061405429)
    p = F32(0.3275911)
    one = F32(1.0)
    t = one / (one + p * ax)
    poly = (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t
    return sgn * (one - poly * np.exp(-ax * ax))


def conv2d_nchw(x: np.ndarray, w: np.ndarray, b: np.ndarray, pads, strides) -> np.ndarray:
    """float im2col convolution with the reference's geometry (`tensor.py:256-264`,
    `numpy_helper.py:18-92`): pads = (ph0, pw0, ph1, pw1), output extent
    ceil((h - kh + ph0 + ph1 + 1) / sh), NHWC patches [kh, kw, c] against
    weights transposed to [kh, kw, c, o], bias added last.
    """
    n, c, h, wd = x.shape
    o, _, kh, kw = w.shape
    ph0, pw0, ph1, pw1 = (int(p) for p in pads)
    sh, sw = (int(s) for s in strides)
    oh = int(np.ceil((h - kh + ph0 + ph1 + 1) / sh))
    ow = int(np.ceil((wd - kw + pw0 + pw1 + 1) / sw))
    xp = np.pad(x.transpose(0, 2, 3, 1), ((0, 0), (ph0, ph1), (pw0, pw1), (0, 0)))
    cols = np.empty((n, oh, ow, kh, kw, c), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, :, i, j, :] = xp[:, i:i + sh * oh:sh, j:j + sw * ow:sw, :]
    wmat = w.transpose(2, 3, 1, 0).reshape(kh * kw * c, o)
    y = cols.reshape(n * oh * ow, kh * kw * c).dot(wmat).reshape(n, oh, ow, o)
    return y.transpose(0, 3, 1, 2) + b[None, :, None, None]
