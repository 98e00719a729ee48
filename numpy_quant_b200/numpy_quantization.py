"""Quantization routines with the names and signatures of the reference's
`numpy_quant/numpy_quantization.py`, executed by the sm_100a kernels of libnq_b200.so.

`quant_parameters` is scalar host arithmetic (two float32 operations per tensor) and
stays on the host, exactly as in the reference.  The array routines accept either host
`np.ndarray`s -- copied to the GPU, processed, copied back, returning the reference's
dtypes (int64 codes, float32 values) -- or CUDA `torch.Tensor`s, which stay resident.
"""
from __future__ import annotations

import numpy as np

__all__ = ["quant_parameters", "quantize", "dequantize", "q_matmul", "requantize"]


def quant_parameters(min_val: np.float32, max_val: np.float32, bit_width: int, asymmetric: bool):
    """Affine parameters from a calibrated range (reference numpy_quantization.py:7-21).

    asymmetric: scale = (max-min)/(2^b-1), zp = rint(lo - min/scale) (int64, not clamped);
    symmetric : scale = 2*max(max,min)/(2^b-1), zp = None.  Arithmetic stays in the dtype of
    the inputs (float32 statistics -> float32); the result scale is a 0-d float32 array.
    """
    lo = -(2.0 ** (bit_width - 1))
    hi = 2.0 ** (bit_width - 1) - 1.0
    levels = hi - lo
    zero_point = None
    with np.errstate(all="ignore"):
        if asymmetric:
            scale = (max_val - min_val) / levels
            zp0 = np.rint(lo - min_val / scale).astype(np.int64)
            # the reference's `zp and np.array(zp)`: a zero stays the int64 scalar itself
            zero_point = np.array(zp0, dtype=np.int64) if zp0 else zp0
        else:
            scale = (2 * max(max_val, min_val)) / levels
    return np.array(scale, dtype=np.float32), zero_point


def _dev(arr):
    import torch
    from .tensor import _to_device
    if isinstance(arr, torch.Tensor):
        return arr, False
    return _to_device(np.asarray(arr)), True


def _zp_int(zero_point):
    return None if zero_point is None else int(np.asarray(zero_point).reshape(-1)[0])


def quantize(data, bit_width: int, scale, zero_point):
    """float32 -> integer codes (reference numpy_quantization.py:24-34)."""
    from . import kernels as K
    t, was_host = _dev(np.asarray(data, dtype=np.float32) if not hasattr(data, "is_cuda") else data)
    if not 2 <= int(bit_width) <= 32:
        raise ValueError("quantize: bit_width must be in 2..32 on the B200 path")
    if bit_width > 8:
        q = K.quantize_i64(t, bit_width, float(scale), _zp_int(zero_point))
    else:
        q = K.quantize(t, bit_width, float(scale), _zp_int(zero_point))
    return q.cpu().numpy().astype(np.int64) if was_host else q


def dequantize(arr, scale, zero_point):
    """integer codes -> float32 (reference numpy_quantization.py:37-41); `zero_point` may be an array."""
    import torch
    from . import kernels as K
    t, was_host = _dev(arr)
    if zero_point is not None and np.asarray(zero_point).size != 1:
        zt, _ = _dev(np.broadcast_to(np.asarray(zero_point, dtype=np.int64), tuple(t.shape)).copy())
        out = K.dequantize(t.to(torch.int64) - zt, float(scale), None)
    else:
        out = K.dequantize(t, float(scale), _zp_int(zero_point))
    return out.cpu().numpy() if was_host else out


def q_matmul(arr_a, scale_a, zero_point_a, arr_b, scale_b, zero_point_b):
    """Exact integer matmul + zero-point bookkeeping (reference numpy_quantization.py:44-61).
    Returns (acc int64, scale float32, zero_point int64 array | None) for host inputs."""
    from .tensor import QTensor
    for name, arr in (("arr_a", arr_a), ("arr_b", arr_b)):
        h = np.asarray(arr)
        if h.size and (h.min() < -128 or h.max() > 127):
            raise ValueError(f"q_matmul: {name} holds codes outside the int8 range (the tensor-core path contracts "
                             "2..8-bit operands; the reference's NumPy int64 matmul has no such limit)")
    a = QTensor(np.asarray(arr_a, dtype=np.int64), 8, scale_a, zero_point_a)
    b = QTensor(np.asarray(arr_b, dtype=np.int64), 8, scale_b, zero_point_b)
    y = a.matmul(b)
    return y.data, y.scale, y.zero_point


def requantize(arr, arr_scale, arr_zero_points, res_scale, res_zero_point, bit_width: int):
    """Wide accumulator -> `bit_width`-bit codes (reference numpy_quantization.py:64-72)."""
    from . import kernels as K
    if not 2 <= int(bit_width) <= 8:
        raise ValueError("requantize: bit_width must be in 2..8 on the B200 path (int8 code storage)")
    d = dequantize(arr, arr_scale, arr_zero_points)
    t, was_host = _dev(d)
    q = K.requantize_f32(t, bit_width, float(res_scale), _zp_int(res_zero_point))
    return q.cpu().numpy().astype(np.int64) if was_host else q
