"""ctypes binding of libnq_b200.so (the C ABI declared in include/nq_b200.h).

There is no fallback: if the shared library is absent or does not load, importing a
kernel entry point raises.  PyTorch is used by callers only to own device memory and
streams; nothing in this file touches torch.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NQ_B200_LIB") or os.path.join(HERE, "libnq_b200.so")   # override: A/B runs of older builds

i64, i32, f32, vp = C.c_int64, C.c_int32, C.c_float, C.c_void_p


class AccZp(C.Structure):
    """struct nq_acc_zp"""
    _fields_ = [("has_zp_a", C.c_int), ("has_zp_b", C.c_int), ("zp_a", i64), ("zp_b", i64), ("k", i64),
                ("rowsum_a", vp), ("colsum_b", vp), ("colsum_batch_stride", i64)]


class Epilogue(C.Structure):
    """struct nq_epilogue"""
    _fields_ = [("mode", C.c_int), ("scale", f32), ("zp", AccZp), ("bias_f32", vp), ("bias_q", vp),
                ("out_bits", C.c_int), ("out_scale", f32), ("has_out_zp", C.c_int), ("out_zp", i64),
                ("residual", vp), ("ld_residual", i64), ("stride_residual", i64),
                ("c_batch_inner", i64), ("stride_c_inner", i64),
                ("q_rows_per_image", i64), ("q_cols_per_head", i64), ("q_off", i64 * 6), ("q_rs", i64 * 6),
                ("q_rowsum", vp), ("sm_has_div", C.c_int), ("sm_div", f32),
                ("q_rowsum_count", i64), ("gelu_div", f32), ("gelu_add", f32), ("gelu_mul", f32), ("reverse_tiles", C.c_int)]


class Attention(C.Structure):
    """struct nq_attention"""
    _fields_ = [("scale_qk", f32), ("has_div", C.c_int), ("div", f32), ("has_zq", C.c_int), ("has_zk", C.c_int),
                ("zq", i64), ("zk", i64),
                ("p_bits", C.c_int), ("p_scale", f32), ("has_p_zp", C.c_int), ("p_zp", i64),
                ("scale_pv", f32), ("has_zv", C.c_int), ("zv", i64),
                ("out_bits", C.c_int), ("out_scale", f32), ("has_out_zp", C.c_int), ("out_zp", i64),
                ("out", vp), ("out_rowsum", vp), ("p_dump", vp), ("ld_p_dump", i64)]


EPI_RAW, EPI_DEQUANT, EPI_REQUANT, EPI_QUANT, EPI_SOFTMAX_QUANT, EPI_GELU_QUANT = 0, 1, 2, 3, 4, 5
UN = dict(neg=0, exp=1, erf=2, tanh=3, sigmoid=4, relu=5, sqrt=6, inv=7, copy=8)
BIN = dict(add=0, mul=1, div=2)

# name -> argtypes (every entry point returns int unless listed in _RESTYPE)
_SIGNATURES = {
    "nq_version": [],
    "nq_last_error": [],
    "nq_device_info": [C.POINTER(C.c_int)],
    "nq_quantize_f32": [vp, i64, C.c_int, f32, C.c_int, i64, vp, vp],
    "nq_quantize_f32_i64": [vp, i64, C.c_int, f32, C.c_int, i64, vp, vp],
    "nq_transpose_s8": [vp, i64, i64, i64, i64, i64, vp, i64, i64, vp],
    "nq_quantize_patches_f32": [vp, i64, i64, i64, i64, i64, i64, C.c_int, f32, C.c_int, i64, vp, i64, vp],
    "nq_quantize_f32_4d": [vp, i64, i64, i64, i64, i64, i64, i64, i64, C.c_int, f32, C.c_int, i64, vp, i64, vp, vp],
    "nq_dequantize": [vp, C.c_int, i64, f32, C.c_int, i64, vp, vp],
    "nq_dequantize_acc": [vp, i64, i64, i64, i64, f32, C.POINTER(AccZp), vp, vp],
    "nq_requantize_acc": [vp, i64, i64, i64, i64, f32, C.POINTER(AccZp), vp, C.c_int, f32, C.c_int, i64, vp, vp],
    "nq_requantize_f32": [vp, i64, C.c_int, f32, C.c_int, i64, vp, vp],
    "nq_rowsum_s8": [vp, i64, i64, i64, vp, vp],
    "nq_qgemm_s8": [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, C.POINTER(Epilogue), vp],
    "nq_attention_s8": [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, C.POINTER(Attention), vp],
    "nq_qconv2d_s8": [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, C.POINTER(Epilogue), vp],
    "nq_qgemm_s8_simt": [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, vp],
    "nq_nhwc_pad": [vp, C.c_int, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                    i64, vp, vp],
    "nq_im2col": [vp, C.c_int, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_int, i32, vp, i64, vp],
    "nq_pack_s8": [vp, i64, C.c_int, vp, vp],
    "nq_unpack_s8": [vp, i64, C.c_int, vp, vp],
    "nq_minmax_init": [vp, i64, vp],
    "nq_minmax_f32": [vp, i64, vp, i64, vp],
    "nq_unary_f32": [C.c_int, vp, i64, vp, vp],
    "nq_binary_f32": [C.c_int, vp, C.POINTER(i64), vp, C.POINTER(i64), C.POINTER(i64), vp, vp],
    "nq_gelu_erf_f32": [vp, i64, f32, f32, f32, vp, vp],
    "nq_layernorm_f32": [vp, i64, i64, i64, vp, vp, f32, vp, vp],
    "nq_softmax_f32": [vp, i64, i64, i64, vp, vp],
    "nq_softmax_div_f32": [vp, i64, i64, i64, f32, vp, vp],
    "nq_layernorm_quantize_f32": [vp, i64, i64, i64, vp, vp, f32, C.c_int, f32, C.c_int, i64, vp, i64, vp, C.c_int, vp],
    "nq_softmax_quantize_f32": [vp, i64, i64, i64, C.c_int, f32, C.c_int, f32, C.c_int, i64, vp, i64, vp, vp],
    "nq_gelu_quantize_f32": [vp, i64, i64, i64, f32, f32, f32, C.c_int, f32, C.c_int, i64, vp, i64, vp, vp],
    "nq_reduce_rows_f32": [C.c_int, vp, i64, i64, vp, vp],
    "nq_memset_async": [vp, C.c_int, i64, vp],
    "nq_selftest_division": [i64, C.c_int, C.c_int, vp, vp],
    "nq_copy_4d": [vp, C.c_int, C.POINTER(i64), C.POINTER(i64), vp, C.POINTER(i64), vp],
}
_RESTYPE = {"nq_last_error": C.c_char_p}

EXPORTS = tuple(_SIGNATURES)
_lib = None


class NqError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libnq_b200.so (once) and attach prototypes; raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NqError(f"{LIB_PATH} not found: build it with `python -m numpy_quant_b200.build` "
                          "(nvcc, sm_100a). There is no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            if "NQ_B200_LIB" in os.environ and not hasattr(lib, name):
                continue                       # an older build under A/B test may lack newer entry points
            fn = getattr(lib, name)            # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Invoke an entry point and raise NqError(nq_last_error()) on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise NqError(f"{name} failed ({rc}): {lib.nq_last_error().decode()}")


def i64x4(values) -> C.Array:
    return (i64 * 4)(*[int(v) for v in values])
