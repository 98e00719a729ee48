"""Thin Python launchers over the C ABI: torch tensors in, torch tensors out.

torch is plumbing here (device allocation, the current CUDA stream); every byte of
arithmetic on the quantized path happens inside libnq_b200.so.  All launchers are
asynchronous on `torch.cuda.current_stream()`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import AccZp, Epilogue, call

LAUNCHES = 0          # kernels enqueued through this module (bench.py reports it as gpu_launches)
GEMM_TIMER = None     # optional list: (ops, start_event, end_event) per tensor-core GEMM launch


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, dtype=None) -> None:
    if not t.is_cuda:
        raise _lib.NqError("numpy_quant_b200 kernels need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# --------------------------------------------------------------------------- K1
def quantize(x: torch.Tensor, bits: int, scale, zp) -> torch.Tensor:
    """float32 (contiguous) -> int8 codes, same shape (numpy_quantization.py:24-34)."""
    _need_cuda(x, torch.float32)
    x = materialize(x)
    out = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    call("nq_quantize_f32", x.data_ptr(), x.numel(), bits, float(scale), int(zp is not None),
         0 if zp is None else int(zp), out.data_ptr(), _stream())
    _count()
    return out


def quantize_patches(x: torch.Tensor, kh: int, kw: int, bits: int, scale, zp) -> "Operand":
    """float32 NCHW image -> quantized patch matrix [(b, oh, ow)][(c, kh, kw)] of a Conv whose kernel == stride and
    pads == 0, written straight into the K-major A operand of the convolution GEMM (no im2col pass)."""
    _need_cuda(x, torch.float32)
    x = materialize(x)
    B, Cc, H, W = (int(v) for v in x.shape)
    k = Cc * kh * kw
    ld = round_up(k, 16)
    rows = B * (H // kh) * (W // kw)
    out = torch.empty((1, rows, ld), dtype=torch.int8, device=x.device) if ld == k else \
        torch.zeros((1, rows, ld), dtype=torch.int8, device=x.device)
    call("nq_quantize_patches_f32", x.data_ptr(), B, Cc, H, W, kh, kw, bits, float(scale), int(zp is not None),
         0 if zp is None else int(zp), out.data_ptr(), ld, _stream())
    _count()
    return Operand(out, (), rows, k, ld, None)


def can_quantize_patches(shape, kh: int, kw: int) -> bool:
    return len(shape) == 4 and kw % 4 == 0 and shape[2] % kh == 0 and shape[3] % kw == 0


def quantize_i64(x: torch.Tensor, bits: int, scale, zp=None) -> torch.Tensor:
    """Wide quantize (4*bit_width-bit biases; 9..32 bit codes) -> int64 codes; `zp` as in numpy_quantization.py:24-34."""
    _need_cuda(x, torch.float32)
    x = materialize(x)
    out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    call("nq_quantize_f32_i64", x.data_ptr(), x.numel(), bits, float(scale), int(zp is not None), 0 if zp is None else int(zp),
         out.data_ptr(), _stream())
    _count()
    return out


@dataclass
class Operand:
    """K-major int8 GEMM operand: data[batch, rows, ld] with `k` valid columns per row."""
    data: torch.Tensor
    batch_shape: tuple
    rows: int
    k: int
    ld: int
    rowsum: Optional[torch.Tensor] = None      # int32 [batch, rows]: sum over k of each row

    @property
    def batch(self) -> int:
        return int(np.prod(self.batch_shape)) if self.batch_shape else 1


def quantize_operand(x: torch.Tensor, role: str, bits: int, scale, zp, want_rowsum: bool) -> Operand:
    """Quantize a (possibly strided, <= 4-D) float32 matrix stack straight into the
    K-major layout the tensor-core GEMM reads.

    role 'A': x[..., M, K] -> rows = M;   role 'B': x[..., K, N] -> rows = N (transposed on the fly).
    """
    _need_cuda(x, torch.float32)
    if x.dim() < 2:
        raise ValueError("matmul operands need >= 2 dims")
    if x.dim() > 4:
        x = x.reshape(-1, *x.shape[-3:])
    lead = tuple(x.shape[:-2])
    st = list(x.stride())
    d = [1] * (4 - x.dim()) + list(x.shape)
    s = [0] * (4 - x.dim()) + st
    if role == "A":
        R, Cc, sr, sc = d[2], d[3], s[2], s[3]
    else:
        R, Cc, sr, sc = d[3], d[2], s[3], s[2]
    ld = round_up(Cc, 16)
    batch = d[0] * d[1]
    out = torch.empty((batch, R, ld), dtype=torch.int8, device=x.device)
    rs = torch.empty((batch, R), dtype=torch.int32, device=x.device) if want_rowsum else None
    call("nq_quantize_f32_4d", x.data_ptr(), d[0], d[1], R, Cc, s[0], s[1], sr, sc, bits, float(scale),
         int(zp is not None), 0 if zp is None else int(zp), out.data_ptr(), ld, _ptr(rs), _stream())
    _count(2 if (want_rowsum and sr == 1 and sc != 1) else 1)
    return Operand(out, lead, R, Cc, ld, rs)


def operand_from_codes(q: torch.Tensor, role: str, want_rowsum: bool) -> Operand:
    """int8 codes [..., M, K] (role A) / [..., K, N] (role B) -> K-major Operand (copy + pad)."""
    _need_cuda(q, torch.int8)
    if q.dim() > 4:
        q = q.reshape(-1, *q.shape[-3:])
    lead = tuple(q.shape[:-2])
    d = [1] * (4 - q.dim()) + list(q.shape)
    s = [0] * (4 - q.dim()) + list(q.stride())
    if role == "B":
        d[2], d[3] = d[3], d[2]
        s[2], s[3] = s[3], s[2]
    R, K = d[2], d[3]
    ld = round_up(K, 16)
    batch = d[0] * d[1]
    out = torch.zeros((batch, R, ld), dtype=torch.int8, device=q.device)
    so = [d[1] * R * ld, R * ld, ld, 1]
    call("nq_copy_4d", q.data_ptr(), 1, _lib.i64x4(d), _lib.i64x4(s), out.data_ptr(), _lib.i64x4(so), _stream())
    _count()
    op = Operand(out, lead, R, K, ld, None)
    if want_rowsum:
        op.rowsum = rowsum(op)
    return op


def rowsum(op: Operand) -> torch.Tensor:
    rs = torch.empty((op.batch, op.rows), dtype=torch.int32, device=op.data.device)
    call("nq_rowsum_s8", op.data.data_ptr(), op.batch * op.rows, op.k, op.ld, rs.data_ptr(), _stream())
    _count()
    return rs


# --------------------------------------------------------------------------- K2 / K3
def dequantize(q: torch.Tensor, scale, zp) -> torch.Tensor:
    _need_cuda(q)
    q = materialize(q)
    eb = {torch.int8: 1, torch.int32: 4, torch.int64: 8}[q.dtype]
    out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    call("nq_dequantize", q.data_ptr(), eb, q.numel(), float(scale), int(zp is not None),
         0 if zp is None else int(zp), out.data_ptr(), _stream())
    _count()
    return out


@dataclass
class AccZeroPoint:
    """Factored zero-point of a q_matmul accumulator (numpy_quantization.py:49-61)."""
    zp_a: Optional[int]
    zp_b: Optional[int]
    k: int
    rowsum_a: Optional[torch.Tensor]       # int32 [batch, M]
    colsum_b: Optional[torch.Tensor]       # int32 [batchB, N]
    colsum_shared: bool                    # B shared across the batch

    def c_struct(self, n: int) -> AccZp:
        z = AccZp()
        z.has_zp_a = int(self.zp_a is not None)
        z.has_zp_b = int(self.zp_b is not None)
        z.zp_a = 0 if self.zp_a is None else int(self.zp_a)
        z.zp_b = 0 if self.zp_b is None else int(self.zp_b)
        z.k = int(self.k)
        z.rowsum_a = _ptr(self.rowsum_a) if self.zp_b is not None else None
        z.colsum_b = _ptr(self.colsum_b) if self.zp_a is not None else None
        z.colsum_batch_stride = 0 if self.colsum_shared else n
        return z

    @property
    def is_none(self) -> bool:
        return self.zp_a is None and self.zp_b is None


def dequantize_acc(acc: torch.Tensor, scale, azp: AccZeroPoint) -> torch.Tensor:
    """int32 accumulator [batch, M, N] -> float32 with the factored zero-point."""
    _need_cuda(acc, torch.int32)
    b, m, n = acc.shape
    out = torch.empty((b, m, n), dtype=torch.float32, device=acc.device)
    z = azp.c_struct(n)
    call("nq_dequantize_acc", acc.data_ptr(), b, m, n, acc.stride(1), float(scale), C.byref(z), out.data_ptr(),
         _stream())
    _count()
    return out


def requantize_acc(acc: torch.Tensor, scale, azp: AccZeroPoint, bias_q: Optional[torch.Tensor], bits: int,
                   out_scale, out_zp) -> torch.Tensor:
    _need_cuda(acc, torch.int32)
    b, m, n = acc.shape
    out = torch.empty((b, m, n), dtype=torch.int8, device=acc.device)
    z = azp.c_struct(n)
    call("nq_requantize_acc", acc.data_ptr(), b, m, n, acc.stride(1), float(scale), C.byref(z), _ptr(bias_q), bits,
         float(out_scale), int(out_zp is not None), 0 if out_zp is None else int(out_zp), out.data_ptr(), _stream())
    _count()
    return out


def requantize_f32(d: torch.Tensor, bits: int, out_scale, out_zp) -> torch.Tensor:
    """clip(rint(zp + (1/s) * d)) on a dequantized float32 tensor -> int8 codes."""
    _need_cuda(d, torch.float32)
    d = d.contiguous()
    out = torch.empty(d.shape, dtype=torch.int8, device=d.device)
    call("nq_requantize_f32", d.data_ptr(), d.numel(), bits, float(out_scale), int(out_zp is not None),
         0 if out_zp is None else int(out_zp), out.data_ptr(), _stream())
    _count()
    return out


# --------------------------------------------------------------------------- K4 / K5
# Walk direction of the large row-ordered producers (L2 reuse between consecutive kernels).  An activation of the
# BASELINE.json batch sizes (155 MB float32 / int8) does not fit the 126 MB L2, but its most recently written part
# does: a consumer launched right after a producer that walked the rows front to back finds the LAST rows in L2 and
# therefore walks back to front (and the other way round).  Results do not depend on the direction.
# _WALK: data_ptr of a recently produced tensor -> True if its producer walked back to front.
_WALK: dict = {}
WALK_REUSE = os.environ.get("NQ_NO_L2_WALK") is None      # A/B switch
_WALK_MIN_BYTES = 48 << 20


def _walk_note(ptr: int, reverse: bool) -> None:
    if len(_WALK) > 256:
        _WALK.clear()
    _WALK[ptr] = bool(reverse)


def _walk_opposite(ptr: int, nbytes: int) -> bool:
    """Direction for a consumer of the tensor at `ptr`: the opposite of its producer's (unknown producer: front to back)."""
    if not WALK_REUSE or nbytes < _WALK_MIN_BYTES:
        return False
    return not _WALK.get(ptr, False)


def qgemm(a: Operand, b: Operand, mode: int = _lib.EPI_RAW, scale: float = 1.0,
          azp: Optional[AccZeroPoint] = None, bias_f32: Optional[torch.Tensor] = None,
          bias_q: Optional[torch.Tensor] = None, out_bits: int = 8, out_scale: float = 1.0, out_zp=None,
          simt: bool = False, residual: Optional[torch.Tensor] = None, heads: int = 0) -> torch.Tensor:
    """C[batch, M, N] = A[batch, M, K] . B[batch|1, N, K]^T on the tcgen05 tensor cores.

    heads=H (batch = B*H): write the result as [B, M, H, N] -- i.e. already Transpose(0,2,1,3)-ed."""
    assert a.k == b.k, f"contraction mismatch {a.k} vs {b.k}"
    batch = max(a.batch, b.batch)
    assert a.batch in (1, batch) and b.batch in (1, batch)
    M, N, K = a.rows, b.rows, a.k
    dtype = {_lib.EPI_RAW: torch.int32, _lib.EPI_DEQUANT: torch.float32, _lib.EPI_REQUANT: torch.int8}[mode]
    # 32-bit outputs get rows padded to a 16-byte multiple so the epilogue can use its vector-store
    # path for ragged N (attention scores, N = 197); callers see a [batch, M, N] view of it
    # (int8 REQUANT rows: padded to 16 bytes so that the thread-per-row epilogue applies to ragged N, e.g. 1000 classes)
    ldn = N if (simt or N % 16 == 0) else round_up(N, 16) if mode == _lib.EPI_REQUANT else N if N % 4 == 0 else round_up(N, 4)
    if heads:
        assert batch % heads == 0 and N % 4 == 0 and mode != _lib.EPI_REQUANT and not simt and residual is None
        out = torch.empty((batch // heads, M, heads, N), dtype=dtype, device=a.data.device)
    else:
        out = torch.empty((batch, M, ldn), dtype=dtype, device=a.data.device)
    sa = 0 if (a.batch == 1 and batch > 1) else M * a.ld
    sb = 0 if (b.batch == 1 and batch > 1) else N * b.ld
    if simt:
        assert mode == _lib.EPI_RAW
        call("nq_qgemm_s8_simt", a.data.data_ptr(), b.data.data_ptr(), out.data_ptr(), M, N, K, batch, a.ld, b.ld, N,
             sa, sb, M * N, _stream())
        _count()
        return out
    ep = Epilogue()
    ep.mode = mode
    ep.scale = float(scale)
    if azp is not None:
        ep.zp = azp.c_struct(N)
    ep.bias_f32 = _ptr(bias_f32)
    ep.bias_q = _ptr(bias_q)
    if residual is not None:
        # float32 [batch, M, N] rows with one uniform row stride (contiguous or row-padded)
        assert mode == _lib.EPI_DEQUANT and residual.dtype == torch.float32 and residual.numel() == batch * M * N
        residual, ldr = rows_layout(residual)
        ep.residual = residual.data_ptr()
        ep.ld_residual = ldr
        ep.stride_residual = M * ldr
    ep.out_bits = out_bits
    ep.out_scale = float(out_scale)
    ep.has_out_zp = int(out_zp is not None)
    ep.out_zp = 0 if out_zp is None else int(out_zp)
    if a.batch == 1 and mode == _lib.EPI_DEQUANT:
        # a large A written by the previous kernel: start where that kernel stopped
        ep.reverse_tiles = int(a.data.data_ptr() in _WALK and _walk_opposite(a.data.data_ptr(), M * a.ld))
        _walk_note(out.data_ptr(), bool(ep.reverse_tiles))
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    ldc, sc = ldn, M * ldn
    if heads:
        ep.c_batch_inner, ep.stride_c_inner = heads, N
        ldc, sc = heads * N, M * heads * N
    call("nq_qgemm_s8", a.data.data_ptr(), b.data.data_ptr(), out.data_ptr(), M, N, K, batch, a.ld, b.ld, ldc,
         sa, sb, sc, C.byref(ep), _stream())
    if timer is not None:
        e1.record()
        timer.append((2 * batch * M * N * K, e0, e1))
    _count()
    return out if (heads or ldn == N) else out[:, :, :N]


def qconv2d(x_nhwc: torch.Tensor, w: Operand, kh: int, kw: int, strides, mode: int = _lib.EPI_RAW, scale: float = 1.0,
            azp: Optional[AccZeroPoint] = None, bias_f32: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, int, int]:
    """Implicit-GEMM convolution (nq_qconv2d_s8): x_nhwc int8 [n, Hp, Wp, C] already padded (pad pixels = zero-point
    code), w the filter matrix operand [O, kh*kw*C] in (kh, kw, c) order.  Returns out[n*OH*OW, O], OH, OW."""
    _need_cuda(x_nhwc)
    assert x_nhwc.dtype == torch.int8 and x_nhwc.is_contiguous() and x_nhwc.dim() == 4
    n, Hp, Wp, Cc = (int(v) for v in x_nhwc.shape)
    sh, sw = (int(v) for v in strides)
    assert w.batch == 1 and w.k == kh * kw * Cc, f"filter matrix K {w.k} != {kh}*{kw}*{Cc}"
    OH, OW = (Hp - kh) // sh + 1, (Wp - kw) // sw + 1
    O = w.rows
    dtype = {_lib.EPI_RAW: torch.int32, _lib.EPI_DEQUANT: torch.float32}[mode]
    out = torch.empty((n * OH * OW, O), dtype=dtype, device=x_nhwc.device)
    ep = Epilogue()
    ep.mode = mode
    ep.scale = float(scale)
    if azp is not None:
        ep.zp = azp.c_struct(O)
    ep.bias_f32 = _ptr(bias_f32)
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("nq_qconv2d_s8", x_nhwc.data_ptr(), w.data.data_ptr(), out.data_ptr(), n, Hp, Wp, Cc, kh, kw, sh, sw, O, w.ld, O,
         C.byref(ep), _stream())
    if timer is not None:
        e1.record()
        timer.append((2 * n * OH * OW * O * w.k, e0, e1))
    _count()
    return out, OH, OW


SPLIT_COLS_VIA_TRANSPOSE = os.environ.get("NQ_SPLIT_COLS_EPILOGUE") is None     # A/B switch: the column-layout epilogue


def qgemm_to_operand(a: Operand, b: Operand, scale: float, azp: AccZeroPoint, bias_f32: Optional[torch.Tensor],
                     bits: int, out_scale, out_zp, kind: str, heads: int, seq: int, want_rowsum: bool,
                     gelu: Optional[tuple] = None) -> Operand:
    """GEMM whose epilogue quantizes the float result (bias + dequant) with the consumer's parameters
    and scatters the codes straight into the K-major operand of the NEXT MatMul (NQ_EPI_QUANT).
    gelu=(c1, c2, c3) applies the graph's Div/Erf/Add/Mul/Mul chain first (NQ_EPI_GELU_QUANT, kind 'rows').

      kind 'rows'        GEMM [batch, M, N] -> operand [batch, M, N] (plain left operand)

      kind 'split_rows'  GEMM [B*S, H*D] -> operand [B*H, S, D]      (attention Q as A, K^T as B)
      kind 'split_cols'  GEMM [B*S, H*D] -> operand [B*H, D, S]      (attention V as B)
      kind 'merge_heads' batched GEMM (B*H) x [S, D] -> operand [B, S, H*D] as one [B*S, H*D] matrix
    """
    assert a.k == b.k
    batch = max(a.batch, b.batch)
    M, N, Kd = a.rows, b.rows, a.k
    dev = a.data.device
    if kind == "split_cols" and SPLIT_COLS_VIA_TRANSPOSE and gelu is None and N % heads == 0 and (N // heads) % 16 == 0:
        # V of the attention: written as [head][S][D] through the row-layout epilogue (the byte-scattering column
        # layout costs 70 us per ViT-B layer against 37 us), then one 1-byte-per-element transpose into [head][D][S]
        rows_op = qgemm_to_operand(a, b, scale, azp, bias_f32, bits, out_scale, out_zp, "split_rows", heads, seq, False)
        BH, S, D = rows_op.batch, rows_op.rows, rows_op.k
        ld = round_up(S, 16)
        out = torch.empty((BH, D, ld), dtype=torch.int8, device=dev)
        call("nq_transpose_s8", rows_op.data.data_ptr(), BH, S, D, rows_op.ld, S * rows_op.ld, out.data_ptr(), ld, D * ld, _stream())
        _count()
        res = Operand(out, rows_op.batch_shape, D, S, ld, None)
        if want_rowsum:
            res.rowsum = rowsum(res)
        return res
    ep = Epilogue()
    ep.mode = _lib.EPI_QUANT
    ep.scale = float(scale)
    ep.zp = azp.c_struct(N)
    ep.bias_f32 = _ptr(bias_f32)
    ep.out_bits, ep.out_scale = bits, float(out_scale)
    ep.has_out_zp, ep.out_zp = int(out_zp is not None), 0 if out_zp is None else int(out_zp)
    if kind in ("split_rows", "split_cols"):
        assert batch == 1 and M % seq == 0 and N % heads == 0
        B, S, H, D = M // seq, seq, heads, N // heads
        if kind == "split_rows":
            ld = round_up(D, 16)
            out = torch.empty((B * H, S, ld), dtype=torch.int8, device=dev)
            res = Operand(out, (B, H), S, D, ld, None)
            off = [0, 0, H * S * ld, ld, S * ld, 1]
            rs = [0, 0, H * S, 1, S, 0]
            n_rs = B * H * S
        else:
            ld = round_up(S, 16)
            out = torch.empty((B * H, D, ld), dtype=torch.int8, device=dev)
            res = Operand(out, (B, H), D, S, ld, None)
            off = [0, 0, H * D * ld, 1, D * ld, ld]
            rs = [0, 0, H * D, 0, D, 1]
            n_rs = B * H * D
        ep.q_rows_per_image, ep.q_cols_per_head = S, D
    elif kind == "merge_heads":
        assert batch % heads == 0
        B, S, H, D = batch // heads, M, heads, N
        ld = H * D
        assert ld % 16 == 0
        out = torch.empty((1, B * S, ld), dtype=torch.int8, device=dev)
        res = Operand(out, (), B * S, ld, ld, None)
        off = [S * H * D, D, 0, H * D, 0, 1]
        rs = [S, 0, 0, 1, 0, 0]
        n_rs = B * S
        ep.q_rows_per_image, ep.q_cols_per_head = S, D
        ep.c_batch_inner = H
    elif kind == "rows":
        # plain [batch, M, N] row-major operand (left operand of the next MatMul contracting over N)
        ld = round_up(N, 16)
        assert ld == N
        out = torch.empty((batch, M, ld), dtype=torch.int8, device=dev)
        lead = a.batch_shape if a.batch == batch else b.batch_shape
        res = Operand(out, tuple(lead), M, N, ld, None)
        off = [M * ld, 0, 0, ld, 0, 1]
        rs = [M, 0, 0, 1, 0, 0]
        n_rs = batch * M
        ep.q_rows_per_image, ep.q_cols_per_head = M, N
    else:
        raise ValueError(kind)
    for i in range(6):
        ep.q_off[i], ep.q_rs[i] = off[i], rs[i]
    if gelu is not None:
        ep.mode = _lib.EPI_GELU_QUANT
        ep.gelu_div, ep.gelu_add, ep.gelu_mul = (float(c) for c in gelu)
    if want_rowsum:
        res.rowsum = torch.empty((n_rs,), dtype=torch.int32, device=dev).view(res.batch, res.rows)
        ep.q_rowsum, ep.q_rowsum_count = res.rowsum.data_ptr(), n_rs     # the library clears them when it needs to
    sa = 0 if (a.batch == 1 and batch > 1) else M * a.ld
    sb = 0 if (b.batch == 1 and batch > 1) else N * b.ld
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if kind == "rows" and batch == 1:
        _walk_note(out.data_ptr(), False)                                 # row-ordered producer of a plain [M, N] operand
    call("nq_qgemm_s8", a.data.data_ptr(), b.data.data_ptr(), out.data_ptr(), M, N, Kd, batch, a.ld, b.ld, N,
         sa, sb, M * N, C.byref(ep), _stream())
    if timer is not None:
        e1.record()
        timer.append((2 * batch * M * N * Kd, e0, e1))
    _count()
    return res


def qgemm_softmax_to_operand(a: Operand, b: Operand, scale: float, azp: AccZeroPoint, div_c, bits: int,
                             out_scale, out_zp, want_rowsum: bool) -> Operand:
    """Attention scores GEMM whose epilogue applies [Div ->] Softmax -> quantize and writes the int8
    left operand [batch, M, round_up(N, 16)] of the following P.V MatMul (NQ_EPI_SOFTMAX_QUANT)."""
    assert a.k == b.k
    batch = max(a.batch, b.batch)
    M, N, Kd = a.rows, b.rows, a.k
    dev = a.data.device
    ld = round_up(N, 16)
    out = torch.empty((batch, M, ld), dtype=torch.int8, device=dev)
    lead = a.batch_shape if a.batch == batch else b.batch_shape
    res = Operand(out, tuple(lead), M, N, ld, None)
    ep = Epilogue()
    ep.mode = _lib.EPI_SOFTMAX_QUANT
    ep.scale = float(scale)
    ep.zp = azp.c_struct(N)
    ep.out_bits, ep.out_scale = bits, float(out_scale)
    ep.has_out_zp, ep.out_zp = int(out_zp is not None), 0 if out_zp is None else int(out_zp)
    ep.sm_has_div, ep.sm_div = int(div_c is not None), 1.0 if div_c is None else float(div_c)
    if want_rowsum:
        res.rowsum = torch.empty((batch, M), dtype=torch.int32, device=dev)      # plain stores: no clearing needed
        ep.q_rowsum = res.rowsum.data_ptr()
    sa = 0 if (a.batch == 1 and batch > 1) else M * a.ld
    sb = 0 if (b.batch == 1 and batch > 1) else N * b.ld
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("nq_qgemm_s8", a.data.data_ptr(), b.data.data_ptr(), out.data_ptr(), M, N, Kd, batch, a.ld, b.ld, ld,
         sa, sb, M * ld, C.byref(ep), _stream())
    if timer is not None:
        e1.record()
        timer.append((2 * batch * M * N * Kd, e0, e1))
    _count()
    return res


def can_fuse_attention(q: Operand, k: Operand, v: Operand) -> bool:
    """Geometry the fused attention kernel is built for: per (image, head) Q [S, D], K^T [S, D], V^T [D, S]."""
    return (q.batch == k.batch == v.batch and q.rows == k.rows == v.k and q.k == k.k == v.rows and q.rows <= 208
            and q.k <= 64 and q.k % 16 == 0 and len(q.batch_shape) == 2)


def can_fuse_attention_qk(q: Operand, k: Operand, zq=None, zk=None) -> bool:
    """The part of `can_fuse_attention` that is known at the score MatMul (V has the same S and D).  zq: the kernel
    folds -zq * colsum(K) into the score MMA as int8 constant passes (|zq| <= 254) and converts x - max(x) exactly,
    which needs max|q - zq| * 128 * D < 2^21 (true for any zero-point inside the int8 range at D <= 64); the zk term is
    constant along a row and cancels in the softmax."""
    if not (q.batch == k.batch and q.rows == k.rows and q.k == k.k and q.rows <= 208 and q.k <= 64 and q.k % 16 == 0
            and len(q.batch_shape) == 2):
        return False
    z = int(zq or 0)
    ra = max(abs(-128 - z), abs(127 - z))
    return abs(z) <= 254 and ra * 128 * q.k < (1 << 21)


def attention(q: Operand, k: Operand, v: Operand, scale_qk: float, zq, zk, div_c, p_bits: int, p_scale, p_zp,
              scale_pv: float, zv, out_bits: int, out_scale, out_zp, want_rowsum: bool, dump_p: bool = False):
    """softmax(Q.K^T / c) . V for every (image, head) in one kernel (nq_attention_s8); returns the int8 left
    operand [1, B*S, H*D] of the output projection (merge heads), plus its row sums on request.  The zero-point
    terms of both MatMuls are accumulated by the tensor core (no row / column sums needed).  dump_p=True additionally
    returns the emitted P codes as an int8 tensor [B*H, S, S] (test hook: unsigned bytes code - lo, converted back)."""
    assert can_fuse_attention(q, k, v)
    B, H = (int(x) for x in q.batch_shape)
    S, D = q.rows, q.k
    dev = q.data.device
    out = torch.empty((1, B * S, H * D), dtype=torch.int8, device=dev)
    res = Operand(out, (), B * S, H * D, H * D, None)
    a = _lib.Attention()
    a.scale_qk = float(scale_qk)
    a.has_div, a.div = int(div_c is not None), 1.0 if div_c is None else float(div_c)
    a.has_zq, a.zq = int(zq is not None), 0 if zq is None else int(zq)
    a.has_zk, a.zk = int(zk is not None), 0 if zk is None else int(zk)
    a.p_bits, a.p_scale = int(p_bits), float(p_scale)
    a.has_p_zp, a.p_zp = int(p_zp is not None), 0 if p_zp is None else int(p_zp)
    a.scale_pv = float(scale_pv)
    a.has_zv, a.zv = int(zv is not None), 0 if zv is None else int(zv)
    a.out_bits, a.out_scale = int(out_bits), float(out_scale)
    a.has_out_zp, a.out_zp = int(out_zp is not None), 0 if out_zp is None else int(out_zp)
    a.out = out.data_ptr()
    if want_rowsum:
        res.rowsum = torch.empty((1, B * S), dtype=torch.int32, device=dev)
        a.out_rowsum = res.rowsum.data_ptr()
    pd = None
    if dump_p:
        ldp = round_up(S, 8)
        pd = torch.zeros((B * H, S, ldp), dtype=torch.uint8, device=dev)
        a.p_dump, a.ld_p_dump = pd.data_ptr(), ldp
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("nq_attention_s8", q.data.data_ptr(), k.data.data_ptr(), v.data.data_ptr(), B * H, H, S, D, q.ld, k.ld, v.ld,
         C.byref(a), _stream())
    if timer is not None:
        e1.record()
        timer.append((4 * B * H * S * S * D, e0, e1, "attention"))
    _count()
    if dump_p:
        lo = -(1 << (int(p_bits) - 1))
        return res, (pd[:, :, :S].to(torch.int16) + lo).to(torch.int8)
    return res


# --------------------------------------------------------------------------- K10 / K11
def minmax_slots(n_slots: int, device) -> torch.Tensor:
    mm = torch.empty((n_slots, 2), dtype=torch.float32, device=device)
    call("nq_minmax_init", mm.data_ptr(), n_slots, _stream())
    _count()
    return mm


def minmax_into(x: torch.Tensor, mm: torch.Tensor, slot: int) -> None:
    _need_cuda(x, torch.float32)
    x = materialize(x)
    call("nq_minmax_f32", x.data_ptr(), x.numel(), mm.data_ptr(), slot, _stream())
    _count()


def pack(q: torch.Tensor, bits: int) -> torch.Tensor:
    _need_cuda(q, torch.int8)
    q = materialize(q)
    out = torch.empty(((q.numel() * bits + 7) // 8,), dtype=torch.uint8, device=q.device)
    call("nq_pack_s8", q.data_ptr(), q.numel(), bits, out.data_ptr(), _stream())
    _count()
    return out


def unpack(packed: torch.Tensor, n: int, bits: int) -> torch.Tensor:
    _need_cuda(packed, torch.uint8)
    out = torch.empty((n,), dtype=torch.int8, device=packed.device)
    call("nq_unpack_s8", packed.data_ptr(), n, bits, out.data_ptr(), _stream())
    _count()
    return out


# --------------------------------------------------------------------------- K7-K9 float glue
def unary(op: str, x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, torch.float32)
    x = materialize(x)
    out = torch.empty_like(x)
    call("nq_unary_f32", _lib.UN[op], x.data_ptr(), x.numel(), out.data_ptr(), _stream())
    _count()
    return out


def gelu_erf(x: torch.Tensor, c_div: float, c_add: float, c_mul: float) -> torch.Tensor:
    _need_cuda(x, torch.float32)
    x = materialize(x)
    out = torch.empty_like(x)
    call("nq_gelu_erf_f32", x.data_ptr(), x.numel(), float(c_div), float(c_add), float(c_mul), out.data_ptr(),
         _stream())
    _count()
    return out


def _pad4(shape, strides):
    n = len(shape)
    return [1] * (4 - n) + list(shape), [0] * (4 - n) + list(strides)


def _collapse_to_4d(t: torch.Tensor, shape) -> torch.Tensor:
    """Broadcast `t` to `shape` (as a stride-0 view) and make it at most 4-D."""
    v = t.expand(shape)
    if v.dim() > 4:
        v = v.reshape(-1, *shape[-3:])          # may copy for exotic stride patterns (not on the hot path)
    return v


def binary(op: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """NumPy-style broadcasting float32 binary op (tensor.py:80-104)."""
    _need_cuda(a, torch.float32)
    _need_cuda(b, torch.float32)
    shape = torch.broadcast_shapes(a.shape, b.shape)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    if out.numel() == 0:
        return out
    va, vb = _collapse_to_4d(a, shape), _collapse_to_4d(b, shape)
    d, sa = _pad4(va.shape, va.stride())
    _, sb = _pad4(vb.shape, vb.stride())
    sa = [0 if dd == 1 else ss for dd, ss in zip(d, sa)]
    sb = [0 if dd == 1 else ss for dd, ss in zip(d, sb)]
    call("nq_binary_f32", _lib.BIN[op], va.data_ptr(), _lib.i64x4(sa), vb.data_ptr(), _lib.i64x4(sb), _lib.i64x4(d),
         out.data_ptr(), _stream())
    _count()
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    _need_cuda(x, torch.float32)
    x = materialize(x)
    cols = x.shape[-1]
    out = torch.empty_like(x)
    call("nq_layernorm_f32", x.data_ptr(), x.numel() // cols, cols, cols, gamma.contiguous().data_ptr(),
         beta.contiguous().data_ptr(), float(eps), out.data_ptr(), _stream())
    _count()
    return out


def softmax_lastdim(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, torch.float32)
    x, ldx = rows_layout(x)
    cols = x.shape[-1]
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call("nq_softmax_f32", x.data_ptr(), x.numel() // cols, cols, ldx, out.data_ptr(), _stream())
    _count()
    return out


def rows_layout(x: torch.Tensor):
    """(tensor, ldx): `x` seen as rows of its last axis with ONE uniform row stride (e.g. the padded
    GEMM output view); falls back to a contiguous copy when the leading dims do not collapse."""
    if x.is_contiguous():
        return x, int(x.shape[-1])
    ok = x.dim() >= 2 and x.stride(-1) == 1
    if ok:
        ld = x.stride(-2)
        ok = ld >= x.shape[-1] and all(x.shape[i] == 1 or x.stride(i) == x.stride(i + 1) * x.shape[i + 1]
                                       for i in range(x.dim() - 2))
    if ok:
        return x, int(ld)
    x = materialize(x)
    return x, int(x.shape[-1])


def softmax_div_lastdim(x: torch.Tensor, div_c: float) -> torch.Tensor:
    """softmax(x / c) over the last axis: the graph's Div + Softmax pair in one pass."""
    _need_cuda(x, torch.float32)
    x, ldx = rows_layout(x)
    cols = x.shape[-1]
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call("nq_softmax_div_f32", x.data_ptr(), x.numel() // cols, cols, ldx, float(div_c), out.data_ptr(), _stream())
    _count()
    return out


def _rows_operand(x: torch.Tensor, want_rowsum: bool):
    """Allocate the A-role operand for a contiguous [..., M, K] float tensor produced row by row."""
    lead = tuple(x.shape[:-2])
    M, Kd = int(x.shape[-2]), int(x.shape[-1])
    batch = int(np.prod(lead)) if lead else 1
    ld = round_up(Kd, 16)
    out = torch.empty((batch, M, ld), dtype=torch.int8, device=x.device)
    rs = torch.empty((batch, M), dtype=torch.int32, device=x.device) if want_rowsum else None
    return Operand(out, lead, M, Kd, ld, rs), batch * M, Kd, ld


def can_fuse_layernorm_quantize(x: torch.Tensor) -> bool:
    return x.dim() >= 2 and x.is_contiguous() and x.shape[-1] % 4 == 0 and x.shape[-1] <= 1024


def layernorm_quantize(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, bits: int, scale, zp,
                       want_rowsum: bool, float_glue: bool = False) -> Operand:
    """LayerNormalization -> quantize (operand A) in one kernel.  float_glue=False: exactly the codes of
    layernorm() followed by quantize_operand(); True: the normalised value is float glue (1e-5 contract)."""
    _need_cuda(x, torch.float32)
    op, rows, cols, ld = _rows_operand(x, want_rowsum)
    flags = int(float_glue) | (2 if float_glue and _walk_opposite(x.data_ptr(), x.numel() * 4) else 0)
    call("nq_layernorm_quantize_f32", x.data_ptr(), rows, cols, cols, gamma.contiguous().data_ptr(),
         beta.contiguous().data_ptr(), float(eps), bits, float(scale), int(zp is not None), 0 if zp is None else int(zp),
         op.data.data_ptr(), ld, _ptr(op.rowsum), flags, _stream())
    _count()
    return op


def can_fuse_softmax_quantize(x: torch.Tensor) -> bool:
    return x.dim() >= 2 and round_up(int(x.shape[-1]), 16) <= 1024


def softmax_quantize(x: torch.Tensor, div_c, bits: int, scale, zp, want_rowsum: bool) -> Operand:
    """[Div ->] Softmax -> quantize (operand A) in one kernel."""
    _need_cuda(x, torch.float32)
    x, ldx = rows_layout(x)
    op, rows, cols, ld = _rows_operand(x, want_rowsum)
    call("nq_softmax_quantize_f32", x.data_ptr(), rows, cols, ldx, int(div_c is not None),
         1.0 if div_c is None else float(div_c), bits, float(scale), int(zp is not None), 0 if zp is None else int(zp),
         op.data.data_ptr(), ld, _ptr(op.rowsum), _stream())
    _count()
    return op


def gelu_quantize(x: torch.Tensor, c_div: float, c_add: float, c_mul: float, bits: int, scale, zp,
                  want_rowsum: bool) -> Operand:
    """GELU chain -> quantize (operand A) in one kernel."""
    _need_cuda(x, torch.float32)
    x = materialize(x)
    op, rows, cols, ld = _rows_operand(x, want_rowsum)
    call("nq_gelu_quantize_f32", x.data_ptr(), rows, cols, cols, float(c_div), float(c_add), float(c_mul), bits,
         float(scale), int(zp is not None), 0 if zp is None else int(zp), op.data.data_ptr(), ld, _ptr(op.rowsum),
         _stream())
    _count()
    return op


def reduce_lastdim(op: str, x: torch.Tensor, keepdims: bool) -> torch.Tensor:
    _need_cuda(x, torch.float32)
    x = materialize(x)
    cols = x.shape[-1]
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    call("nq_reduce_rows_f32", {"max": 0, "sum": 1, "mean": 2}[op], x.data_ptr(), x.numel() // cols, cols,
         out.data_ptr(), _stream())
    _count()
    return out.unsqueeze(-1) if keepdims else out


def materialize(x: torch.Tensor) -> torch.Tensor:
    """Contiguous copy of a strided view (Transpose / Expand / Slice) with our own copy kernel."""
    _need_cuda(x)
    if x.is_contiguous():
        return x
    out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    if x.numel() == 0:
        return out
    v = x if x.dim() <= 4 else None
    if v is None:
        return x.contiguous()                    # > 4-D strided views never occur on the quantized path
    d, sx = _pad4(v.shape, v.stride())
    so = [d[1] * d[2] * d[3], d[2] * d[3], d[3], 1]
    call("nq_copy_4d", v.data_ptr(), x.element_size(), _lib.i64x4(d), _lib.i64x4(sx), out.data_ptr(),
         _lib.i64x4(so), _stream())
    _count()
    return out


def can_nhwc_pad(C_: int, W: int) -> bool:
    return C_ % 4 == 0 and W * (C_ + 4) <= 48 * 1024


def nhwc_pad(x: torch.Tensor, pads, pad_code: int, quant: Optional[tuple] = None) -> torch.Tensor:
    """x[B,C,H,W] int8 codes -- or float32 with quant=(bits, scale, zp|None), quantized on the way -- to the padded
    NHWC int8 image [B, H+ph0+ph1, W+pw0+pw1, C] that nq_qconv2d_s8 reads (pad pixels = pad_code)."""
    _need_cuda(x)
    x = materialize(x)
    B, Cc, H, W = (int(v) for v in x.shape)
    ph0, pw0, ph1, pw1 = (int(p) for p in pads)
    out = torch.empty((B, H + ph0 + ph1, W + pw0 + pw1, Cc), dtype=torch.int8, device=x.device)
    if quant is None:
        assert x.dtype == torch.int8
        call("nq_nhwc_pad", x.data_ptr(), 1, B, Cc, H, W, ph0, pw0, ph1, pw1, int(pad_code), 8, 1.0, 0, 0, out.data_ptr(), _stream())
    else:
        assert x.dtype == torch.float32
        bits, scale, zp = quant
        call("nq_nhwc_pad", x.data_ptr(), 4, B, Cc, H, W, ph0, pw0, ph1, pw1, int(pad_code), int(bits), float(scale),
             int(zp is not None), 0 if zp is None else int(zp), out.data_ptr(), _stream())
    _count()
    return out


def im2col(x: torch.Tensor, kh: int, kw: int, pads, strides, pad_value=0) -> tuple[torch.Tensor, int, int]:
    """x[B,C,H,W] (int8 or float32) -> patches [B*OH*OW, ld] in (kh, kw, c) order (numpy_helper.py:18-92)."""
    _need_cuda(x)
    x = materialize(x)
    B, Cc, H, W = x.shape
    ph0, pw0, ph1, pw1 = (int(p) for p in pads)
    sh, sw = (int(s) for s in strides)
    OH = -((H - kh + ph0 + ph1 + 1) // -sh)
    OW = -((W - kw + pw0 + pw1 + 1) // -sw)
    k = kh * kw * Cc
    if x.dtype == torch.int8:
        ld, eb, pv = round_up(k, 16), 1, int(pad_value)
    else:
        ld, eb = k, 4
        pv = int(np.float32(pad_value).view(np.int32))
    out = torch.empty((B * OH * OW, ld), dtype=x.dtype, device=x.device)
    call("nq_im2col", x.data_ptr(), eb, B, Cc, H, W, kh, kw, ph0, pw0, ph1, pw1, sh, sw, pv, out.data_ptr(), ld,
         _stream())
    _count()
    return out, OH, OW
