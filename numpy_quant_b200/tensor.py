"""Device-resident ITensor / FTensor / QTensor with the operator surface of the
reference's `numpy_quant/tensor.py` (same class, method and function names, same
argument meaning, same exceptions).

Storage lives in HBM (torch CUDA tensors are only the allocator); `.data` hands back a
NumPy array (a cached device->host copy; int64 for quantized tensors as in
tensor.py:158-166).  All arithmetic runs in libnq_b200.so:

* FTensor ops   -> float32 kernels with NumPy's op-for-op rounding (tensor.py:47-152)
* QTensor.matmul-> tcgen05 int8 GEMM; the result is a LAZY accumulator tensor: the
  GEMM is launched when the consumer is known, with that consumer fused into the
  epilogue -- `.dequantize()` (tensor.py:189-193), `.requantize()` (tensor.py:195-199,
  incl. the `+ bias` of tensor.py:183-187) -- or raw int32 when `.data` is read.
* float `FTensor.matmul` is the one library call: a plain fp32 cuBLAS GEMM through
  torch.matmul (TF32 off); it is only used by the float calibration pass.
"""
from __future__ import annotations

from typing import Any, Optional, Union

import numpy as np
import torch

from . import _lib, kernels as K
from .numpy_quantization import quant_parameters


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.NqError("numpy_quant_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(arr: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(arr)).to(_device(), non_blocking=False)


class ITensor:
    """int64 shape / index tensor; lives on the host (reference tensor.py:12-44)."""

    def __init__(self, data: np.ndarray):
        self._data = data

    @property
    def data(self):
        return self._data

    def expand_dims(self, axis: "ITensor"):
        return ITensor(np.expand_dims(self._data, axis=tuple(axis.data)))

    @property
    def shape(self):
        return ITensor(np.array(self._data.shape, dtype=np.int64))

    @property
    def size(self):
        return self._data.size

    def __eq__(self, other: "ITensor"):
        return ITensor(np.array(self._data == other.data, np.int64))

    __hash__ = None

    def __getitem__(self, ind):
        return ITensor(self._data.__getitem__(ind))

    def __mul__(self, other: "ITensor"):
        return ITensor(self._data * other.data)

    def reshape(self, shape: "ITensor"):
        return ITensor(self._data.reshape(shape.data))

    def take(self, indices: "ITensor", axis: int):
        return ITensor(self._data.take(np.atleast_1d(indices.data), axis))


class FTensor:
    """float32 tensor in HBM (reference tensor.py:47-152)."""

    def __init__(self, data: Union[np.ndarray, torch.Tensor]):
        if isinstance(data, torch.Tensor):
            if data.dtype != torch.float32:
                raise ValueError("User np.float32 for FTensor")
            if not data.is_cuda:
                data = data.to(_device())
            self._t = data
            self._host = None
        else:
            data = np.asarray(data)
            if not data.dtype == np.float32:
                raise ValueError("User np.float32 for FTensor")
            self._t = _to_device(data)
            self._host = None

    # -- host view ---------------------------------------------------------
    @property
    def data(self) -> np.ndarray:
        if self._host is None:
            self._host = self._t.detach().cpu().numpy()
        return self._host

    @property
    def device_tensor(self) -> torch.Tensor:
        return self._t

    @property
    def shape(self):
        return ITensor(np.array(tuple(self._t.shape), dtype=np.int64))

    @property
    def T(self):
        return FTensor(self._t.permute(*reversed(range(self._t.dim()))))

    def copy(self):
        return FTensor(K.unary("copy", self._t))

    def reshape(self, shape: ITensor):
        target = [int(s) for s in np.asarray(shape.data).reshape(-1)]
        t = self._t
        try:
            return FTensor(t.view(target))
        except RuntimeError:
            return FTensor(K.materialize(t).view(target))

    def take(self, indices: ITensor, axis: int):
        idx = np.asarray(indices.data)
        if idx.ndim == 0:
            return FTensor(self._t.select(axis, int(idx)))
        pieces = [FTensor(self._t.narrow(axis, int(i), 1)) for i in idx.reshape(-1)]
        out = concat(pieces, axis=axis)
        if idx.ndim > 1:
            shp = list(self._t.shape)
            shp[axis:axis + 1] = list(idx.shape)
            out = FTensor(out._t.view(shp))
        return out

    def transpose(self, *axes):
        if len(axes) == 1 and not isinstance(axes[0], (int, np.integer)):
            axes = tuple(axes[0])
        if not axes:
            return self.T
        return FTensor(self._t.permute(*[int(a) for a in axes]))

    def __neg__(self):
        return FTensor(K.unary("neg", self._t))

    def __mul__(self, other: "FTensor"):
        if isinstance(other, FTensor):
            return FTensor(K.binary("mul", self._t, other._t))
        raise ValueError(f"Value of type {type(other)} cannot be multiplied")

    def __add__(self, other):
        if isinstance(other, FTensor):
            return FTensor(K.binary("add", self._t, other._t))
        if isinstance(other, float):
            return FTensor(K.binary("add", self._t, _scalar(other)))
        raise ValueError(f"Value of type {type(other)} cannot be added")

    def __radd__(self, other):
        return self.__add__(other)

    def __getitem__(self, ind):
        return FTensor(self._t.__getitem__(ind))

    def matmul(self, other: "FTensor"):
        # plain library GEMM (fp32 cuBLAS, TF32 disabled) -- calibration pass only
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return FTensor(torch.matmul(self._t, other._t))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def div(self, other: "FTensor"):
        return FTensor(K.binary("div", self._t, other._t))

    def erf(self):
        return FTensor(K.unary("erf", self._t))

    def exp(self):
        return FTensor(K.unary("exp", self._t))

    def expand(self, shape: "ITensor"):
        # ONNX Expand semantics on top of broadcasting (tensor.py:112-119)
        curr_shape = self.shape.data
        new_shape = np.array(shape.data, dtype=np.int64).copy()
        adjust_loc = np.logical_and(new_shape < curr_shape, new_shape == 1)
        new_shape[adjust_loc] = curr_shape[adjust_loc]
        return FTensor(self._t.expand(*[int(s) for s in new_shape]))

    def inv(self):
        return FTensor(K.unary("inv", self._t))

    def _reduce(self, op: str, axis: int, keepdims: bool):
        nd = self._t.dim()
        axis = axis % nd
        t = self._t if axis == nd - 1 else self._t.movedim(axis, -1)
        r = K.reduce_lastdim(op, K.materialize(t), False)
        return FTensor(r.unsqueeze(axis) if keepdims else r)

    def max(self, axis: int, keepdims: bool):
        return self._reduce("max", axis, keepdims)

    def mean(self, axis: int, keepdims: bool):
        return self._reduce("mean", axis, keepdims)

    def relu(self):
        return FTensor(K.unary("relu", self._t))

    def sigmoid(self):
        return FTensor(K.unary("sigmoid", self._t))

    def sum(self, axis: int, keepdims: bool):
        return self._reduce("sum", axis, keepdims)

    def _softmax(self, axis: int):
        m = self + (-(self.max(axis=axis, keepdims=True)))
        e = m.exp()
        return m, e, e.sum(axis=axis, keepdims=True)

    def softmax(self, axis: int):
        nd = self._t.dim()
        if axis % nd == nd - 1:
            return FTensor(K.softmax_lastdim(self._t))
        t = self._t.movedim(axis, -1)
        return FTensor(K.softmax_lastdim(K.materialize(t)).movedim(-1, axis))

    def layernorm(self, gamma: "FTensor", beta: "FTensor", epsilon: float):
        """Fused form of the recipe in model.py:134-152 (last axis)."""
        return FTensor(K.layernorm(self._t, gamma._t, beta._t, epsilon))

    def gelu_erf(self, c_div: float, c_add: float, c_mul: float):
        """Div -> Erf -> Add -> Mul -> Mul chain of the ViT graph in one pass."""
        return FTensor(K.gelu_erf(self._t, c_div, c_add, c_mul))

    def sqrt(self):
        return FTensor(K.unary("sqrt", self._t))

    def tanh(self):
        return FTensor(K.unary("tanh", self._t))


_SCALARS: dict = {}


def _scalar(v: float) -> torch.Tensor:
    key = (float(np.float32(v)), torch.cuda.current_device())
    if key not in _SCALARS:
        _SCALARS[key] = torch.tensor(key[0], dtype=torch.float32, device=_device())
    return _SCALARS[key]


def _as_opt_int(zp) -> Optional[int]:
    return None if zp is None else int(np.asarray(zp).reshape(-1)[0])


class QTensor:
    """Quantized tensor (reference tensor.py:155-221): integer codes + bit_width + scale +
    zero_point.  Codes of bit_width <= 8 are int8 in HBM, matmul accumulators int32,
    wide (4*bit_width) biases int64; `.data` always returns int64 like the reference.

    Internal forms (all resolved on demand by `_codes()`):
      _q        codes in logical layout (torch int8 / int32 / int64), or None
      _oplayout (Operand, role, logical shape): codes that so far exist only in the K-major
                operand layout of the GEMM (written there directly by quantize_tensor)
      _nhwc     (padded NHWC image, pads, logical NCHW shape): codes of a Conv input that so far exist only in
                the layout nq_qconv2d_s8 reads (written there by quantize_tensor_nhwc)
      _patches  (patch-matrix Operand, (kh, kw), logical NCHW shape): codes of the input of a Conv whose patches tile
                the image (kernel == stride, no padding), written by quantize_tensor_patches in the layout the
                convolution GEMM reads
      _lazy     pending GEMM of a matmul result (operands + optional int bias)
      _packed   (role, packed bitstream, operand geometry, row sums): sub-byte storage of a weight -- the K-major
                operand buffer packed to bit_width bits per code (nq_pack_s8); unpacked into a transient int8
                operand right before each GEMM (the tensor pipe has no int4 / int2 kind)
    """

    def __init__(self, data, bit_width: int, scale, zero_point=None):
        self._lazy = None
        self._oplayout = None
        self._nhwc = None
        self._patches = None
        self._ops: dict = {}           # cached K-major GEMM operands by role
        self._packed = None
        self._src = None               # base QTensor when this is the 2-D transpose of it
        self._host = None
        self._deq = None
        if isinstance(data, np.ndarray):
            if data.dtype != np.int64:
                raise ValueError("Use np.int64 for quantized tensors")
            lo, hi = (int(data.min()), int(data.max())) if data.size else (0, 0)
            if bit_width <= 8 and -128 <= lo and hi <= 127:
                self._q = _to_device(data.astype(np.int8))
            elif -2 ** 31 <= lo and hi < 2 ** 31:
                self._q = _to_device(data.astype(np.int32))
            else:
                self._q = _to_device(data)
        elif isinstance(data, torch.Tensor) or data is None:
            self._q = data
        else:
            raise ValueError("Use np.int64 for quantized tensors")
        if (zero_point is not None) and not isinstance(zero_point, K.AccZeroPoint):
            if np.asarray(zero_point).dtype != np.int64:
                raise ValueError("Use np.int64 for zero_point of quantized tensors")
        self.bit_width = bit_width
        self.scale = scale
        self._zp = zero_point

    # -- zero point: scalar / host array / factored device form -------------------------
    @property
    def zero_point(self):
        z = self._zp
        if not isinstance(z, K.AccZeroPoint):
            return z
        if z.is_none:
            return None
        bshape = tuple(self.shape[:-2])
        out = 0
        if z.zp_b is not None:
            rs = z.rowsum_a.detach().cpu().numpy().astype(np.int64)
            out = out + rs.reshape(bshape + (-1, 1)) * np.int64(z.zp_b)
        if z.zp_a is not None:
            cs = z.colsum_b.detach().cpu().numpy().astype(np.int64)
            cshape = (1,) * len(bshape) if z.colsum_shared else bshape
            out = out + cs.reshape(cshape + (1, -1)) * np.int64(z.zp_a)
        if z.zp_a is not None and z.zp_b is not None:
            out = out - np.int64(z.zp_a) * np.int64(z.zp_b) * np.int64(z.k)
        return out

    @zero_point.setter
    def zero_point(self, value):
        self._zp = value

    # -- shape / views ---------------------------------------------------------------
    @property
    def shape(self):
        if self._q is not None:
            return tuple(self._q.shape)
        if self._oplayout is not None:
            return tuple(self._oplayout[2])
        if self._nhwc is not None:
            return tuple(self._nhwc[2])
        if self._patches is not None:
            return tuple(self._patches[2])
        if self._packed is not None:
            return tuple(self._packed["logical"])
        L = self._lazy
        return tuple(L["batch_shape"]) + (L["M"], L["N"])

    def _pending(self) -> bool:
        return self._q is None and self._lazy is not None

    def _codes(self) -> torch.Tensor:
        """Integer codes in logical layout (un-pads an operand / runs the pending GEMM raw)."""
        if self._q is None:
            if self._packed is not None and self._oplayout is None:
                P = self._packed                                # observability only: the GEMM path never comes here
                op = self._unpacked_operand()
                v = op.data[:, :, : op.k]
                if P["role"] == "B":
                    v = v.transpose(1, 2)
                return K.materialize(v).view(P["logical"])      # not cached: storage stays packed
            if self._oplayout is not None:
                op, role, shape = self._oplayout
                v = op.data[:, :, : op.k]
                if role == "B":
                    v = v.transpose(1, 2)
                self._q = K.materialize(v).view(shape)
            elif self._nhwc is not None:
                img, (ph0, pw0, _, _), shape = self._nhwc
                self._q = K.materialize(img[:, ph0:ph0 + shape[2], pw0:pw0 + shape[3], :].permute(0, 3, 1, 2))
            elif self._patches is not None:
                op, (kh, kw), (b, c, h, w) = self._patches
                v = op.data[0, :, : op.k].view(b, h // kh, w // kw, c, kh, kw).permute(0, 3, 1, 4, 2, 5)
                self._q = K.materialize(v).view(b, c, h, w)
            else:
                L = self._lazy
                acc = K.qgemm(L["a"], L["b"])
                if L.get("bias_q") is not None:
                    acc = acc.to(torch.int64) + L["bias_q"]
                self._q = acc.view(*L["batch_shape"], L["M"], L["N"])
        return self._q

    @property
    def T(self):
        q = self._codes()
        zp = self.zero_point
        out = QTensor(q.permute(*reversed(range(q.dim()))), self.bit_width, self.scale,
                      None if zp is None else np.asarray(zp).T)
        if q.dim() == 2:
            out._src = self
        return out

    def reshape(self, shape: ITensor):
        q = self._codes()
        target = [int(s) for s in np.asarray(shape.data).reshape(-1)]
        try:
            v = q.view(target)
        except RuntimeError:
            v = K.materialize(q).view(target)
        return QTensor(v, self.bit_width, self.scale, self.zero_point)

    def transpose(self, *axes):
        if len(axes) == 1 and not isinstance(axes[0], (int, np.integer)):
            axes = tuple(axes[0])
        return QTensor(self._codes().permute(*[int(a) for a in axes]), self.bit_width, self.scale, self.zero_point)

    def __add__(self, other: "QTensor"):
        """Integer add of the codes (tensor.py:183-187); keeps self's bit_width/scale/zero_point.
        On a pending accumulator with a per-column vector it becomes the GEMM's int bias."""
        if not isinstance(other, QTensor):
            raise ValueError(f"Cannot add QTensor with {other.__class__}")
        if self._pending() and self._lazy.get("bias_q") is None and len(other.shape) == 1 \
                and other.shape[0] == self._lazy["N"]:
            out = QTensor(None, self.bit_width, self.scale, self._zp)
            out._lazy = dict(self._lazy, bias_q=other._codes().to(torch.int64).contiguous())
            return out
        a, b = self._codes().to(torch.int64), other._codes().to(torch.int64)    # generic, off the hot path
        return QTensor(a + b, self.bit_width, self.scale, self.zero_point)

    # -- K2 ----------------------------------------------------------------------------
    def dequantize_heads_last(self) -> Optional[FTensor]:
        """dequantize().transpose(0, 2, 1, 3), materialised: a pending [B, H, S, D] accumulator is
        written by the GEMM epilogue directly as [B, S, H, D] (the attention context layout).
        Returns None when this accumulator does not qualify."""
        L = self._lazy
        if not self._pending() or L.get("bias_q") is not None or len(L["batch_shape"]) != 2 or L["b"].rows % 4 \
                or L["a"].batch != int(np.prod(L["batch_shape"])):
            return None
        out = K.qgemm(L["a"], L["b"], _lib.EPI_DEQUANT, float(self.scale), self._zp, heads=int(L["batch_shape"][1]))
        return FTensor(out)

    def quantize_into_operand(self, bias: Optional[FTensor], bit_width: int, scale, zero_point, kind: str,
                              heads: int, seq: int, want_rowsum: bool, role: str, logical_shape: tuple):
        """dequantize (+ bias) -> [Reshape / Transpose] -> quantize for the next MatMul, all inside the
        GEMM epilogue: the int8 K-major operand of the consumer is the only thing written.  Same float32
        and rounding steps as the separate kernels, so the codes are identical.  Returns None when this
        accumulator does not qualify (caller falls back to the unfused route)."""
        L = self._lazy
        zo = _as_opt_int(zero_point)
        if not self._pending() or L.get("bias_q") is not None or (zo is not None and abs(zo) >= (1 << 20)):
            return None
        M, N = L["a"].rows, L["b"].rows
        if kind == "merge_heads":
            if len(L["batch_shape"]) != 2 or int(L["batch_shape"][1]) != heads or N % 16 or (heads * N) % 16 \
                    or L["a"].batch != int(np.prod(L["batch_shape"])):
                return None
        else:
            if max(L["a"].batch, L["b"].batch) != 1 or N % heads or (N // heads) % 16 or M % seq:
                return None
        try:
            op = K.qgemm_to_operand(L["a"], L["b"], float(self.scale), self._zp,
                                    None if bias is None else bias.device_tensor.contiguous(), bit_width,
                                    float(scale), zo, kind, heads, seq, want_rowsum)
        except _lib.NqError:
            return None
        return qtensor_from_operand(op, role, logical_shape, bit_width, scale, zero_point)

    def gelu_into_operand(self, bias: Optional[FTensor], consts: tuple, bit_width: int, scale, zero_point,
                          want_rowsum: bool):
        """dequantize (+ bias) -> Div/Erf/Add/Mul/Mul (GELU) -> quantize for the next MatMul's left operand, all
        inside the epilogue of the pending GEMM (NQ_EPI_GELU_QUANT).  Float glue under the 1e-5 contract, so the
        codes may differ from the node-by-node route where the float value sits on a rounding boundary.
        Returns None when this accumulator does not qualify."""
        L = self._lazy
        zo = _as_opt_int(zero_point)
        if not self._pending() or L.get("bias_q") is not None or (zo is not None and abs(zo) >= (1 << 20)):
            return None
        nb = int(np.prod(L["batch_shape"] or (1,)))
        flat = L["a"].batch == 1 and L["b"].batch == 1 and nb > 1          # shared-weight GEMM run as one flat batch
        if L["N"] % 16 or L["a"].rows != (nb * L["M"] if flat else L["M"]):
            return None
        try:
            op = K.qgemm_to_operand(L["a"], L["b"], float(self.scale), self._zp,
                                    None if bias is None else bias.device_tensor.contiguous(), bit_width,
                                    float(scale), zo, "rows", 1, L["a"].rows, want_rowsum,
                                    gelu=tuple(float(c) for c in consts))
        except _lib.NqError:
            return None
        if flat:
            rs = None if op.rowsum is None else op.rowsum.view(nb, L["M"])
            op = K.Operand(op.data.view(nb, L["M"], op.ld), tuple(L["batch_shape"]), L["M"], op.k, op.ld, rs)
        else:
            op.batch_shape = tuple(L["batch_shape"])
        return qtensor_from_operand(op, "A", tuple(L["batch_shape"]) + (L["M"], L["N"]), bit_width, scale, zero_point)

    def attention_into_operand(self, div_c, v: "QTensor", bit_width: int, p_scale, p_zero_point, out_scale, out_zero_point,
                               want_rowsum: bool):
        """self = pending Q.K^T scores [B, H, S, S]; v = quantized V [B, H, S, D].  dequantize -> [/ c] -> softmax ->
        quantize(p) -> MatMul(., V) -> dequantize -> Transpose(0,2,1,3) -> Reshape -> quantize(out): the whole
        attention of the graph in one kernel (nq_attention_s8); returns the [B, S, H*D] left operand of the output
        projection.  Same codes as softmax_into_operand + quantize_into_operand(merge_heads)."""
        L = self._lazy
        qa, kb = L["a"], L["b"]
        vb = v._operand("B", False)
        if not K.can_fuse_attention(qa, kb, vb):
            raise ValueError("attention_into_operand: geometry not supported by the fused kernel")
        za, zk = _as_opt_int(self._zp.zp_a), _as_opt_int(self._zp.zp_b)
        zp_p, zv = _as_opt_int(p_zero_point), _as_opt_int(v._scalar_zp())
        scale_pv = np.float32(p_scale) * np.float32(v.scale)
        op = K.attention(qa, kb, vb, float(self.scale), za, zk, div_c, bit_width, float(p_scale), zp_p, float(scale_pv), zv,
                         bit_width, float(out_scale), _as_opt_int(out_zero_point), want_rowsum)
        B, H = (int(x) for x in qa.batch_shape)
        return qtensor_from_operand(op, "A", (B, qa.rows, H * qa.k), bit_width, out_scale, out_zero_point)

    def softmax_into_operand(self, div_c, bit_width: int, scale, zero_point, want_rowsum: bool):
        """dequantize -> [/ c] -> softmax(last axis) -> quantize for the next MatMul (as its left operand),
        all inside the epilogue of the pending attention-score GEMM.  None when not applicable."""
        L = self._lazy
        zo = _as_opt_int(zero_point)
        if not self._pending() or L.get("bias_q") is not None or (zo is not None and abs(zo) >= (1 << 20)):
            return None
        if L["N"] > 224 or L["a"].batch != int(np.prod(L["batch_shape"] or (1,))) or L["a"].rows != L["M"]:
            return None
        try:
            op = K.qgemm_softmax_to_operand(L["a"], L["b"], float(self.scale), self._zp, div_c, bit_width,
                                            float(scale), zo, want_rowsum)
        except _lib.NqError:
            return None
        op.batch_shape = tuple(L["batch_shape"])
        return qtensor_from_operand(op, "A", tuple(L["batch_shape"]) + (L["M"], L["N"]), bit_width, scale, zero_point)

    def dequantize(self, bias: Optional[FTensor] = None, residual: Optional[FTensor] = None) -> FTensor:
        """tensor.py:189-193. `bias` (float32 [N]) and `residual` (float32, result shape) optionally
        fuse the bias Add / residual Add that follow in the graph: (bias + dequant) + residual."""
        if self._pending() and self._lazy.get("bias_q") is None:
            if bias is None and residual is None and self._deq is not None:
                return self._deq
            L = self._lazy
            out = K.qgemm(L["a"], L["b"], _lib.EPI_DEQUANT, float(self.scale), self._zp,
                          bias_f32=None if bias is None else bias.device_tensor.contiguous(),
                          residual=None if residual is None else residual.device_tensor)
            res = FTensor(out.view(*L["batch_shape"], L["M"], L["N"]))
            if bias is None and residual is None:
                self._deq = res
            return res
        q = self._codes()
        z = self._zp
        if isinstance(z, K.AccZeroPoint):
            q3 = K.materialize(q).view(-1, q.shape[-2], q.shape[-1])
            if q3.dtype != torch.int32:
                raise ValueError("accumulator with a factored zero-point must be int32")
            res = FTensor(K.dequantize_acc(q3, float(self.scale), z).view(q.shape))
        elif z is None or np.asarray(z).size == 1:
            res = FTensor(K.dequantize(K.materialize(q), float(self.scale), _as_opt_int(z)))
        else:
            # array zero-point supplied by a caller: integer subtract, then scalar dequantize
            zt = _to_device(np.broadcast_to(np.asarray(z, dtype=np.int64), tuple(q.shape)).copy())
            res = FTensor(K.dequantize(K.materialize(q).to(torch.int64) - zt, float(self.scale), None))
        if bias is not None:
            res = bias + res
        return res if residual is None else res + residual

    # -- K3 ----------------------------------------------------------------------------
    def requantize(self, bit_width: int, scale: np.float32, zero_point: np.int64):
        """tensor.py:195-199; fused into the GEMM epilogue when the accumulator is pending."""
        zo = _as_opt_int(zero_point)
        if self._pending():
            L = self._lazy
            out = K.qgemm(L["a"], L["b"], _lib.EPI_REQUANT, float(self.scale), self._zp, bias_q=L.get("bias_q"),
                          out_bits=bit_width, out_scale=float(scale), out_zp=zo)
            return QTensor(out.view(*L["batch_shape"], L["M"], L["N"]), bit_width, scale, zero_point)
        d = self.dequantize().device_tensor
        return QTensor(K.requantize_f32(d, bit_width, float(scale), zo), bit_width, scale, zero_point)

    @property
    def data(self) -> np.ndarray:
        if self._host is None:
            self._host = K.materialize(self._codes()).detach().cpu().numpy().astype(np.int64)
        return self._host

    # -- K4 / K5 -----------------------------------------------------------------------
    # -- sub-byte storage ---------------------------------------------------------------
    def pack_storage(self, role: str = "B") -> int:
        """Keep only the bit_width-bit packed form of this tensor's GEMM operand (plus its row sums); returns the
        bytes now resident.  `.data` still works (unpacks on demand)."""
        if self._src is not None:
            return self._src.pack_storage("A" if role == "B" else "B")
        if self._packed is not None:
            return int(self._packed["data"].numel())
        op = self._operand(role, True)
        logical = tuple(self.shape)
        packed = K.pack(op.data.reshape(-1), self.bit_width)
        self._packed = dict(role=role, data=packed, n=int(op.data.numel()), dshape=tuple(op.data.shape),
                            batch_shape=tuple(op.batch_shape), rows=op.rows, k=op.k, ld=op.ld, rowsum=op.rowsum,
                            logical=logical)
        self._q, self._oplayout, self._deq, self._host = None, None, None, None
        self._ops.clear()
        return int(packed.numel())

    def _unpacked_operand(self) -> K.Operand:
        P = self._packed
        data = K.unpack(P["data"], P["n"], self.bit_width).view(P["dshape"])
        return K.Operand(data, P["batch_shape"], P["rows"], P["k"], P["ld"], P["rowsum"])

    def _operand(self, role: str, want_rowsum: bool) -> K.Operand:
        if self._src is not None:                       # transposed 2-D view: roles swap, cache on the base
            return self._src._operand("A" if role == "B" else "B", want_rowsum)
        if self._packed is not None and self._packed["role"] == role:
            return self._unpacked_operand()             # transient int8 operand, freed after the GEMM
        op = self._ops.get(role)
        if op is None:
            q = self._codes()
            if q.dtype != torch.int8:
                raise ValueError("integer matmul operands must be <= 8-bit codes")
            op = K.operand_from_codes(q, role, want_rowsum)
            self._ops[role] = op
        elif want_rowsum and op.rowsum is None:
            op.rowsum = K.rowsum(op)
        return op

    def _scalar_zp(self):
        z = self._zp
        if isinstance(z, K.AccZeroPoint) or (z is not None and np.asarray(z).size != 1):
            raise ValueError("this operation needs a scalar (or no) zero-point")
        return z

    def matmul(self, other: "QTensor"):
        """tensor.py:205-210 / numpy_quantization.py:44-61: exact integer contraction on the
        tensor cores; the result carries 4*bit_width bits, scale_a*scale_b and the factored
        zero-point, and stays PENDING until its consumer is known."""
        assert self.bit_width == other.bit_width, f"{self.bit_width} != {other.bit_width}"
        if self.bit_width > 8:
            raise ValueError("integer matmul operands must have bit_width <= 8")
        za, zb = _as_opt_int(self._scalar_zp()), _as_opt_int(other._scalar_zp())
        sa_shape, sb_shape = self.shape, other.shape
        if len(sa_shape) < 2 or len(sb_shape) < 2:
            raise ValueError("QTensor.matmul needs operands with >= 2 dims")
        ba, bb = tuple(sa_shape[:-2]), tuple(sb_shape[:-2])
        batch_shape = tuple(np.broadcast_shapes(ba, bb))
        a_q, b_q = self, other
        # the device GEMM takes one flat batch axis, each operand either batched or shared;
        # partial broadcasting (e.g. (2,1,..) x (1,2,..)) is expanded explicitly
        if ba != batch_shape and int(np.prod(ba or (1,))) != 1:
            a_q = QTensor(self._codes().expand(*batch_shape, *sa_shape[-2:]), self.bit_width, self.scale, self._zp)
        if bb != batch_shape and int(np.prod(bb or (1,))) != 1:
            b_q = QTensor(other._codes().expand(*batch_shape, *sb_shape[-2:]), other.bit_width, other.scale,
                          other._zp)
        if int(np.prod(ba or (1,))) == 1 and int(np.prod(bb or (1,))) != 1:
            a_q = QTensor(self._codes().expand(*batch_shape, *sa_shape[-2:]), self.bit_width, self.scale, self._zp)
        opa = a_q._operand("A", zb is not None)
        opb = b_q._operand("B", za is not None)
        if opa.k != opb.k:
            raise ValueError(f"matmul: shapes {sa_shape} and {sb_shape} not aligned")
        scale = np.float32(self.scale) * np.float32(other.scale)
        azp = K.AccZeroPoint(za, zb, opa.k, opa.rowsum if zb is not None else None,
                             opb.rowsum if za is not None else None, colsum_shared=(opb.batch == 1))
        out = QTensor(None, 4 * self.bit_width, scale, azp)
        M, N = int(sa_shape[-2]), int(sb_shape[-1])                    # logical extents (an operand may be flat)
        if opb.batch == 1 and opa.batch > 1:
            # one weight for the whole batch: run ONE [batch*M, K] x [K, N] GEMM instead of `batch` small ones
            # (no partially filled 128-row tiles per batch entry); the operand buffer is already contiguous
            rs = None if opa.rowsum is None else opa.rowsum.view(1, -1)
            opa = K.Operand(opa.data.view(1, opa.batch * opa.rows, opa.ld), (), opa.batch * opa.rows, opa.k, opa.ld, rs)
            if azp.rowsum_a is not None:
                azp = K.AccZeroPoint(za, zb, opa.k, rs, azp.colsum_b, azp.colsum_shared)
                out._zp = azp
        out._lazy = dict(a=opa, b=opb, batch_shape=batch_shape, M=M, N=N)
        return out

    def relu(self):
        # clamp codes at the zero-point (tensor.py:212-215); off the hot path
        q = self._codes().to(torch.int64)
        zp = _as_opt_int(self._scalar_zp())
        return QTensor(torch.clamp(q, min=zp), self.bit_width, self.scale, self.zero_point)

    def sigmoid(self):
        # dequantize -> sigmoid -> quantize with own params (tensor.py:217-221)
        act = self.dequantize().sigmoid()
        return quantize_tensor(act, self.bit_width, self.scale, self._scalar_zp())


Tensor = Union[ITensor, FTensor, QTensor]


def quantize_tensor(tensor: FTensor, bit_width: int, scale: np.float32, zero_point, role: Optional[str] = None,
                    want_rowsum: bool = False) -> QTensor:
    """tensor.py:227-229. With `role` ('A' / 'B') the codes are written directly in the
    K-major operand layout of the tensor-core GEMM (`.data` un-pads them on demand)."""
    if bit_width > 8:
        if bit_width > 32:
            raise ValueError("bit_width must be <= 32 on the B200 path")
        return QTensor(K.quantize_i64(tensor.device_tensor, bit_width, float(scale), _as_opt_int(zero_point)), bit_width,
                       scale=scale, zero_point=zero_point)
    if bit_width < 2:
        raise ValueError("bit_width must be in 2..8 on the B200 path")
    zi = _as_opt_int(zero_point)
    t = tensor.device_tensor
    if role is None or t.dim() < 2:
        return QTensor(K.quantize(t, bit_width, float(scale), zi), bit_width, scale=scale, zero_point=zero_point)
    op = K.quantize_operand(t, role, bit_width, float(scale), zi, want_rowsum)
    return qtensor_from_operand(op, role, tuple(t.shape), bit_width, scale, zero_point)


def quantize_tensor_nhwc(tensor: FTensor, bit_width: int, scale: np.float32, zero_point, pads) -> Optional[QTensor]:
    """quantize_tensor for a value consumed only by Conv nodes with these pads: float32 NCHW is quantized straight into
    the padded NHWC image of the implicit-GEMM convolution (pad pixels = the zero-point code, the quantized 0.0 of the
    reference's float padding).  None when the geometry is not served by that route."""
    t = tensor.device_tensor
    zi = _as_opt_int(zero_point)
    if t.dim() != 4 or not (2 <= bit_width <= 8) or not IMPLICIT_CONV:
        return None
    lo, hi = -(1 << (bit_width - 1)), (1 << (bit_width - 1)) - 1
    if int(t.shape[1]) % 64 != 0 or not K.can_nhwc_pad(int(t.shape[1]), int(t.shape[3])) or (zi is not None and not lo <= zi <= hi):
        return None
    pads = tuple(int(p) for p in pads)
    img = K.nhwc_pad(t, pads, 0 if zi is None else zi, quant=(bit_width, float(scale), zi))
    out = QTensor(None, bit_width, scale=scale, zero_point=zero_point)
    out._nhwc = (img, pads, tuple(int(v) for v in t.shape))
    return out


def quantize_tensor_patches(tensor: FTensor, bit_width: int, scale: np.float32, zero_point, kernel) -> Optional[QTensor]:
    """quantize_tensor for a value consumed only by Conv nodes whose patches tile the image (kernel == stride, no
    padding: the ViT patch embedding): the float32 image is quantized straight into the patch matrix the convolution
    GEMM reads, so the Conv is a reshape + MatMul and no im2col pass runs.  None when the geometry is not served."""
    t = tensor.device_tensor
    zi = _as_opt_int(zero_point)
    kh, kw = (int(v) for v in kernel)
    lo, hi = -(1 << (bit_width - 1)), (1 << (bit_width - 1)) - 1
    if t.dim() != 4 or not (2 <= bit_width <= 8) or not K.can_quantize_patches(tuple(t.shape), kh, kw) \
            or (zi is not None and not lo <= zi <= hi):
        return None
    op = K.quantize_patches(t, kh, kw, bit_width, float(scale), zi)
    out = QTensor(None, bit_width, scale=scale, zero_point=zero_point)
    out._patches = (op, (kh, kw), tuple(int(v) for v in t.shape))
    return out


def qtensor_from_operand(op: K.Operand, role: str, shape: tuple, bit_width: int, scale, zero_point) -> QTensor:
    """Wrap codes that exist only in the K-major GEMM operand layout (`.data` un-pads on demand)."""
    out = QTensor(None, bit_width, scale=scale, zero_point=zero_point)
    out._ops[role] = op
    out._oplayout = (op, role, tuple(shape))
    return out


def tensor_min_max(tensor: Tensor):
    """tensor.py:232-236: min/max widened to contain 0 (device reduction for F/Q tensors)."""
    zero_val = np.array(0.0).astype(np.float32)
    if isinstance(tensor, FTensor):
        mm = K.minmax_slots(1, tensor.device_tensor.device)
        K.minmax_into(tensor.device_tensor, mm, 0)
        lo, hi = mm.cpu().numpy()[0]
        return np.minimum(np.float32(lo), zero_val), np.maximum(np.float32(hi), zero_val)
    return np.minimum(tensor.data.min(), zero_val), np.maximum(tensor.data.max(), zero_val)


def quantize_tensor_min_max(tensor: Tensor, bit_width: int, asymmetric: bool):
    min_val, max_val = tensor_min_max(tensor)
    scale, zero_point = quant_parameters(min_val, max_val, bit_width, asymmetric)
    return quantize_tensor(tensor, bit_width, scale, zero_point)


def concat(x_list: list, axis: int):
    """tensor.py:245-248 (ITensor on the host, FTensor with the strided-copy kernel)."""
    assert all(x.__class__ == x_list[0].__class__ for x in x_list), (
        f"types {[x.__class__ for x in x_list]} of x_list entries do no match")
    if isinstance(x_list[0], ITensor):
        return ITensor(np.concatenate([x.data for x in x_list], axis=axis))
    if not isinstance(x_list[0], FTensor):
        raise ValueError("concat supports ITensor and FTensor")
    ts = [x.device_tensor for x in x_list]
    nd = ts[0].dim()
    axis = axis % nd
    shape = list(ts[0].shape)
    shape[axis] = sum(int(t.shape[axis]) for t in ts)
    out = torch.empty(shape, dtype=torch.float32, device=ts[0].device)
    off = 0
    for t in ts:
        n = int(t.shape[axis])
        dst = out.narrow(axis, off, n)
        _copy_into(t, dst)
        off += n
    return FTensor(out)


def _copy_into(src: torch.Tensor, dst: torch.Tensor) -> None:
    if src.numel() == 0:
        return
    if src.dim() > 4:
        src, dst = src.reshape(-1, *src.shape[-3:]), dst.view(-1, *dst.shape[-3:])
    d = [1] * (4 - src.dim()) + list(src.shape)
    sx = [0] * (4 - src.dim()) + list(src.stride())
    so = [0] * (4 - dst.dim()) + list(dst.stride())
    _lib.call("nq_copy_4d", src.data_ptr(), src.element_size(), _lib.i64x4(d), _lib.i64x4(sx), dst.data_ptr(),
              _lib.i64x4(so), torch.cuda.current_stream().cuda_stream)
    K._count()


def where(condition: ITensor, a: Tensor, b: Tensor):
    """tensor.py:251-253 (shape arithmetic: ITensor only on this path)."""
    assert a.__class__ == b.__class__, f"types {a.__class__} and {b.__class__} do not match"
    if isinstance(a, ITensor):
        return ITensor(np.where(condition.data, a.data, b.data))
    if isinstance(a, FTensor):
        # off the quantized hot path (shape arithmetic uses ITensor): an elementwise select, plain torch indexing
        cond = torch.from_numpy(np.ascontiguousarray(np.asarray(condition.data) != 0)).to(a.device_tensor.device)
        return FTensor(torch.where(cond, a.device_tensor, b.device_tensor))
    raise ValueError("where() cannot build a QTensor (the reference's tensor.py:251-253 fails for QTensor as well)")


IMPLICIT_CONV = True     # A/B switch: False materialises the patch matrix (nq_im2col + nq_qgemm_s8)


def fconv2d(x: FTensor, w: FTensor, b: FTensor, pads, strides):
    """Float im2col convolution (tensor.py:256-264, numpy_helper.py:18-92): used by the
    float calibration pass; GEMM through the fp32 library matmul."""
    xt, wt = x.device_tensor, w.device_tensor
    o, c, kh, kw = (int(s) for s in wt.shape)
    cols, oh, ow = K.im2col(xt, kh, kw, pads, strides, 0.0)
    wmat = FTensor(K.materialize(wt.permute(2, 3, 1, 0)).view(kh * kw * c, o))
    y = FTensor(cols).matmul(wmat)
    y = y + FTensor(b.device_tensor.view(1, o))
    n = int(xt.shape[0])
    return FTensor(y.device_tensor.view(n, oh, ow, o).permute(0, 3, 1, 2))


def qconv2d(x: QTensor, w: QTensor, b: FTensor, pads, strides) -> FTensor:
    """Integer im2col conv: im2col(q_x padded with zp_x) . q_w on the tensor cores, zero-point
    correction + dequantize + float bias in the epilogue.  Equals the reference's fake-quant
    float conv (model.py:95-100 on dequantized inputs) up to float32 summation rounding."""
    zx = _as_opt_int(x._scalar_zp())
    if w._scalar_zp() is not None:
        raise ValueError("qconv2d expects symmetric weights")
    if zx is not None and not (-128 <= zx <= 127):
        return fconv2d(x.dequantize(), w.dequantize(), b, pads, strides)
    wq = w._codes()
    o, c, kh, kw = (int(s) for s in wq.shape)
    xshape = tuple(int(v) for v in x.shape)
    k = kh * kw * c
    if x._patches is not None and x._patches[1] == (kh, kw) and tuple(int(s) for s in strides) == (kh, kw) \
            and not any(int(p) for p in pads):
        # the patches tile the image: the input was quantized straight into the patch matrix (rows (b, oh, ow), columns
        # (c, kh, kw)), the filters are used in their natural [O, C*KH*KW] order -- reshape + MatMul, nothing else
        opa = x._patches[0]
        opw = w._ops.get("conv_chw")
        if opw is None:
            opw = K.operand_from_codes(wq.reshape(o, k), "A", True)        # rows = O, rowsum = colsum(B)
            w._ops["conv_chw"] = opw
        azp_p = K.AccZeroPoint(zx, None, k, None, opw.rowsum, True)
        y = K.qgemm(opa, opw, _lib.EPI_DEQUANT, float(np.float32(x.scale) * np.float32(w.scale)), azp_p,
                    bias_f32=b.device_tensor.contiguous())
        n_img, oh, ow = xshape[0], xshape[2] // kh, xshape[3] // kw
        return FTensor(y.view(n_img, oh, ow, o).permute(0, 3, 1, 2))
    opb = w._ops.get("conv")
    if opb is None:
        wk = K.materialize(wq.permute(0, 2, 3, 1)).view(o, k)            # [O, kh*kw*c] == K-major B operand
        opb = K.operand_from_codes(wk, "A", True)                         # rows = O, rowsum = colsum(B)
        w._ops["conv"] = opb
    azp = K.AccZeroPoint(zx, None, k, None, opb.rowsum, True)
    scale = np.float32(x.scale) * np.float32(w.scale)
    n = xshape[0]
    sh, sw = (int(s) for s in strides)
    pads = tuple(int(p) for p in pads)
    hp, wp = xshape[2] + pads[0] + pads[2], xshape[3] + pads[1] + pads[3]
    if IMPLICIT_CONV and x.bit_width <= 8 and c % 64 == 0 and sh <= 8 and sw <= 8 and hp >= kh and wp >= kw \
            and K.can_nhwc_pad(c, xshape[3]):
        # implicit GEMM: the patch matrix is never written.  The codes live in (or are relaid to) a padded NHWC image
        # (pad pixels = zp_x, i.e. the quantized zero of the reference's float padding); the GEMM's im2col-mode TMA
        # reads the (kh, kw) taps of every output pixel straight from it.
        if x._nhwc is not None and x._nhwc[1] == pads:
            nhwc = x._nhwc[0]
        else:
            nhwc = K.nhwc_pad(x._codes(), pads, 0 if zx is None else zx)
        y, oh, ow = K.qconv2d(nhwc, opb, kh, kw, (sh, sw), _lib.EPI_DEQUANT, float(scale), azp,
                              bias_f32=b.device_tensor.contiguous())
        return FTensor(y.view(n, oh, ow, o).permute(0, 3, 1, 2))
    xq = x._codes()
    cols, oh, ow = K.im2col(xq, kh, kw, pads, strides, 0 if zx is None else zx)
    opa = K.Operand(cols.view(1, cols.shape[0], cols.shape[1]), (), cols.shape[0], k, cols.shape[1], None)
    y = K.qgemm(opa, opb, _lib.EPI_DEQUANT, float(scale), azp, bias_f32=b.device_tensor.contiguous())
    return FTensor(y.view(n, oh, ow, o).permute(0, 3, 1, 2))
