"""ORACLE (test infrastructure, not product code) -- CPU restatement of the
reference's graph import, float interpreter, calibration and quantized
interpreter (`/root/reference/numpy_quant/model.py`, `tensor.py`).

Written as a functional interpreter over tagged NumPy values rather than the
reference's tensor classes; every rule cites the reference lines it restates.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  Parity is PINNED against the unmodified
reference by `tests/golden/make_golden.py` -> `tests/test_oracle_golden.py`
(MLP at 2/4/8 bit, Gemm/MatMul/Conv graphs, a small ViT: every quantization
parameter and the outputs, bit for bit).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from time import perf_counter
from typing import Any, NamedTuple

import numpy as np

from . import ref_quant as rq

F32, I64 = np.float32, np.int64


class F(NamedTuple):          # float32 tensor            (reference FTensor, tensor.py:47-152)
    a: np.ndarray


class I(NamedTuple):          # int64 shape/index tensor  (reference ITensor, tensor.py:12-44)
    a: np.ndarray


class Q(NamedTuple):          # quantized tensor          (reference QTensor, tensor.py:155-221)
    a: np.ndarray             # int64 codes / accumulators
    bits: int
    scale: Any                # float32
    zp: Any                   # None | int64 scalar | int64 array


def q_dequant(q: Q) -> F:
    return F(rq.dequantize(q.a, q.scale, q.zp))


def q_quant(x: F, bits: int, scale, zp) -> Q:
    return Q(rq.quantize(x.a, bits, scale, zp), bits, scale, zp)


def q_matmul(x: Q, w: Q) -> Q:
    assert x.bits == w.bits, f"{x.bits} != {w.bits}"                      # tensor.py:206
    acc, scale, zp = rq.q_matmul(x.a, x.scale, x.zp, w.a, w.scale, w.zp)
    return Q(acc, 4 * x.bits, scale, zp)                                  # tensor.py:210


def q_transpose2(q: Q) -> Q:                                              # tensor.py:172-175
    return Q(q.a.T, q.bits, q.scale, None if q.zp is None else np.asarray(q.zp).T)


# --------------------------------------------------------------------------
# graph container
# --------------------------------------------------------------------------
@dataclass
class Op:
    name: str
    kind: str
    attrs: dict
    ins: list[str]
    outs: list[str]


@dataclass
class Graph:
    ops: list[Op]
    consts: dict[str, F]                  # initializers (float32 only, model.py:254-256)
    inputs: list[str]
    outputs: list[str]
    value_order: list[str]                # initializers, graph inputs, node outputs (model.py:292)
    env: dict[str, Any] = field(default_factory=dict)   # values of the last float run


def _attr_value(attr, to_array, get_attribute_value, tensor_type):
    v = get_attribute_value(attr)
    return to_array(v) if isinstance(v, tensor_type) else v              # model.py:57-62


def import_graph(model, onnx_api) -> Graph:
    """`Model.from_onnx` (model.py:249-292) for any object exposing `.graph`.

    `onnx_api` supplies `to_array`, `get_attribute_value`, `TensorProto`
    (`numpy_quant_b200.onnx_lite` or the real `onnx` helpers).
    """
    g = model.graph
    consts, order = {}, []
    for t in g.initializer:
        arr = np.array(onnx_api.to_array(t))
        if arr.dtype != F32:
            raise ValueError("User np.float32 for FTensor")               # tensor.py:49-50
        consts[t.name] = F(arr)
        order.append(t.name)
    inputs = [v.name for v in g.input]
    order += [n for n in inputs if n not in consts]
    ops = []
    for n in g.node:
        attrs = {a.name: _attr_value(a, onnx_api.to_array, onnx_api.get_attribute_value, onnx_api.TensorProto)
                 for a in n.attribute}
        ops.append(Op(n.name, n.op_type, attrs, list(n.input), list(n.output)))
        for o in n.output:
            if o not in order:
                order.append(o)
    return Graph(ops, consts, inputs, [v.name for v in g.output], order)


# --------------------------------------------------------------------------
# float operator table (model.py:65-213 + tensor.py FTensor/ITensor methods)
# --------------------------------------------------------------------------
def _add(a, b):
    if isinstance(a, F) and isinstance(b, F):
        return F(a.a + b.a)                                               # tensor.py:86-88
    if isinstance(a, Q) and isinstance(b, Q):
        return Q(a.a + b.a, a.bits, a.scale, a.zp)                        # tensor.py:183-185
    raise ValueError(f"cannot add {type(a)} and {type(b)}")


def _mul(a, b):
    if isinstance(a, I):
        return I(a.a * b.a)                                               # tensor.py:37-38
    return F(a.a * b.a)                                                   # tensor.py:80-82


def _softmax(x: F, axis: int) -> F:                                       # tensor.py:139-146
    m = x.a + (-(x.a.max(axis=axis, keepdims=True)))
    e = np.exp(m)
    return F(e / e.sum(axis=axis, keepdims=True))


def _layernorm(x: F, gamma: F, beta: F, axis: int, eps: float) -> F:      # model.py:134-152
    mean = x.a.mean(axis=axis, keepdims=True)
    d = x.a + (-mean)
    var = (d * d).mean(axis=axis, keepdims=True)
    inv = 1 / np.sqrt(var + eps)
    return F(d * inv * gamma.a + beta.a)


def _expand(x: F, shape: I) -> F:                                         # tensor.py:112-119
    cur = np.array(x.a.shape, dtype=I64)
    new = shape.a.copy()
    fix = np.logical_and(new < cur, new == 1)
    new[fix] = cur[fix]
    return F(np.broadcast_to(x.a, tuple(new)))


def _slice(x, starts, ends, axes):                                        # model.py:182-190
    sl = [slice(None)] * x.a.ndim
    for s, e, ax in zip(starts.a, ends.a, axes.a):
        sl[ax] = slice(s, e)
    return type(x)(x.a[tuple(sl)])


def _gemm(ins, attrs):                                                    # model.py:122-131
    x, w, b = ins
    if attrs.get("transA"):
        x = q_transpose2(x) if isinstance(x, Q) else F(x.a.T)
    if attrs.get("transB"):
        w = q_transpose2(w) if isinstance(w, Q) else F(w.a.T)
    return _add(_matmul(x, w), b)


def _matmul(a, b):                                                        # model.py:153-157
    if isinstance(a, Q):
        return q_matmul(a, b)
    return F(np.matmul(a.a, b.a))                                         # tensor.py:100-101


def _constant(value):                                                     # model.py:75-83
    if value.dtype == F32:
        return F(value)
    if value.dtype == I64:
        return I(value)
    raise ValueError(f"Constant value type {value.dtype} not supported.")


def _sigmoid(x: F) -> F:                                                  # tensor.py:133-134
    return F(1 / (1.0 + np.exp(-x.a)))


def apply_op(kind: str, ins: list, attrs: dict) -> list:
    """One node of `onnx_operator_implementation` (model.py:65-213)."""
    if kind == "Add":
        return [_add(ins[0], ins[1])]
    if kind == "Concat":                                                  # tensor.py:245-248
        assert all(type(x) is type(ins[0]) for x in ins)
        return [type(ins[0])(np.concatenate([x.a for x in ins], axis=attrs["axis"]))]
    if kind == "Constant":
        return [_constant(attrs["value"])]
    if kind == "ConstantOfShape":                                         # model.py:84-94
        v = attrs["value"]
        return [_constant(np.full(tuple(ins[0].a), fill_value=v, dtype=v.dtype))]
    if kind == "Conv":                                                    # model.py:95-100
        return [F(rq.conv2d_nchw(ins[0].a, ins[1].a, ins[2].a, tuple(attrs["pads"]), tuple(attrs["strides"])))]
    if kind == "Div":
        return [F(ins[0].a / ins[1].a)]
    if kind == "Equal":                                                   # tensor.py:31-32
        return [I(np.array(ins[0].a == ins[1].a, I64))]
    if kind == "Erf":
        return [F(rq.erf_poly(ins[0].a))]
    if kind == "Expand":
        return [_expand(ins[0], ins[1])]
    if kind == "Gather":                                                  # tensor.py:43-44, 71-72
        idx = np.atleast_1d(ins[1].a) if isinstance(ins[0], I) else ins[1].a
        return [type(ins[0])(ins[0].a.take(idx, attrs["axis"]))]
    if kind == "Gemm":
        return [_gemm(ins, attrs)]
    if kind == "Identity":
        return [F(ins[0].a.copy())]
    if kind == "LayerNormalization":
        return [_layernorm(ins[0], ins[1], ins[2], attrs["axis"], attrs["epsilon"])]
    if kind == "MatMul":
        return [_matmul(ins[0], ins[1])]
    if kind == "Mul":
        return [_mul(ins[0], ins[1])]
    if kind == "Relu":                                                    # tensor.py:130-131 (-0.0 for x<0)
        return [F((ins[0].a > 0) * ins[0].a)]
    if kind == "Reshape":
        x = ins[0]
        if isinstance(x, Q):                                              # tensor.py:177-178
            return [Q(x.a.reshape(ins[1].a), x.bits, x.scale, x.zp)]
        return [type(x)(x.a.reshape(ins[1].a))]
    if kind == "Sigmoid":
        return [_sigmoid(ins[0])]
    if kind == "Shape":
        return [I(np.array(ins[0].a.shape, dtype=I64))]
    if kind == "Slice":
        return [_slice(*ins[:4])]
    if kind == "Softmax":
        return [_softmax(ins[0], attrs["axis"])]
    if kind == "Tanh":
        return [F(np.tanh(ins[0].a))]
    if kind == "Transpose":
        x = ins[0]
        if isinstance(x, Q):                                              # tensor.py:180-181
            return [Q(x.a.transpose(attrs["perm"]), x.bits, x.scale, x.zp)]
        return [F(x.a.transpose(attrs["perm"]))]
    if kind == "Where":                                                   # tensor.py:251-253
        assert type(ins[1]) is type(ins[2])
        return [type(ins[1])(np.where(ins[0].a, ins[1].a, ins[2].a))]
    raise ValueError(f"ONNX operand {kind} not supported.")


def _wrap_input(arr: np.ndarray):
    if arr.dtype == F32:
        return F(arr)
    if arr.dtype == I64:
        return I(arr)
    raise ValueError(f"Array dtype {arr.dtype} not supported")            # model.py:300-305


def run_float(g: Graph, inputs: list[np.ndarray], profile: bool = False):
    """`Model.__call__` (model.py:294-326). Leaves every value in `g.env`."""
    env: dict[str, Any] = dict(g.consts)
    for name, arr in zip(g.inputs, inputs):
        env[name] = _wrap_input(arr.copy())
    times = {op.kind: 0.0 for op in g.ops}
    for op in g.ops:
        t0 = perf_counter()
        outs = apply_op(op.kind, [env[i] for i in op.ins], op.attrs)
        times[op.kind] += perf_counter() - t0
        for o, v in zip(op.outs, outs):
            env[o] = v
    g.env = env
    res = [env[o].a for o in g.outputs]
    return (res, times) if profile else res


# --------------------------------------------------------------------------
# calibration -> quantized plan (model.py:328-442)
# --------------------------------------------------------------------------
@dataclass
class QPlan:
    graph: Graph
    bits: int
    qparams: dict[str, tuple]             # value name -> (scale, zp)
    qconsts: dict[str, Q]                 # quantized initializers (b-bit, or 4b-bit biases)
    env: dict[str, Any] = field(default_factory=dict)


def _stat(arr: np.ndarray, fn):                                           # model.py:333-336
    flat = arr.reshape((arr.shape[0], -1) if arr.shape else (-1,))
    return np.mean(fn(flat))


def calibrate(g: Graph, calib_inputs: list[np.ndarray], bits: int = 8) -> QPlan:
    """`Model.quantize` (model.py:328-442): one float pass, global min/max per
    value, symmetric b-bit constants, asymmetric activations, 4b-bit biases."""
    run_float(g, calib_inputs)
    vmin = {n: _stat(g.env[n].a, np.min) for n in g.value_order}
    vmax = {n: _stat(g.env[n].a, np.max) for n in g.value_order}

    def params(name: str, asym: bool):
        return rq.quant_parameters(vmin[name], vmax[name], bits, asym)

    qp: dict[str, tuple] = {}
    qc: dict[str, Q] = {}
    for n in g.inputs:                                                    # model.py:349-355
        qp[n] = params(n, n not in g.consts)
    for n in g.value_order:                                               # model.py:357-365
        if n in g.consts:
            s, z = params(n, False)
            qc[n] = q_quant(g.consts[n], bits, s, z)
            qp[n] = (s, z)
    for op in g.ops:
        out = op.outs[0]
        if op.kind == "Gemm":                                             # model.py:374-394
            for n in op.ins[:2]:
                if n not in g.consts:
                    qp[n] = params(n, True)
            bias = op.ins[2]
            bscale = qp[op.ins[0]][0] * qp[op.ins[1]][0]
            qp[bias] = (bscale, None)
            qc[bias] = q_quant(g.consts[bias], 4 * bits, bscale, None)
            qp[out] = params(out, True)
        elif op.kind == "Add" and (op.ins[0] in g.consts or op.ins[1] in g.consts):   # model.py:395-415
            bi = 0 if op.ins[0] in g.consts else 1
            bias, other = op.ins[bi], op.ins[1 - bi]
            bscale = qp[other][0]
            qc[bias] = q_quant(g.consts[bias], 4 * bits, bscale, None)
            qp[bias] = (bscale, None)
            qp[out] = params(out, True)
        elif op.kind in ("Identity", "Relu"):                             # model.py:416-420
            qp[out] = qp[op.ins[0]]
        else:                                                             # model.py:368-373, 421-425
            qp[out] = params(out, True)
    return QPlan(g, bits, qp, qc)


def run_quant(p: QPlan, inputs: list[np.ndarray], profile: bool = False):
    """`QModel.__call__` (model.py:486-565)."""
    g = p.graph
    env: dict[str, Any] = dict(p.qconsts)
    for name, arr in zip(g.inputs, inputs):                               # model.py:488-495
        if arr.dtype == F32:
            env[name] = q_quant(F(arr), p.bits, *p.qparams[name])
        elif arr.dtype == I64:
            env[name] = I(arr)
        else:
            raise ValueError(f"Array dtype {arr.dtype} not supported")
    times = {op.kind: 0.0 for op in g.ops}
    times["TinyqQuant"] = 0.0
    times["TinyqDequant"] = 0.0
    for op in g.ops:
        ins = []
        if op.kind in ("MatMul", "Gemm"):                                 # model.py:503-527
            for n in op.ins:
                v = env[n]
                if isinstance(v, F):
                    t0 = perf_counter()
                    v = q_quant(v, p.bits, *p.qparams[n])
                    times["TinyqQuant"] += perf_counter() - t0
                ins.append(v)
        else:                                                             # model.py:528-538
            for n in op.ins:
                v = env[n]
                if isinstance(v, Q):
                    t0 = perf_counter()
                    v = q_dequant(v)
                    times["TinyqDequant"] += perf_counter() - t0
                ins.append(v)
        t0 = perf_counter()
        outs = apply_op(op.kind, ins, op.attrs)
        times[op.kind] += perf_counter() - t0
        for o, v in zip(op.outs, outs):
            if op.kind == "Gemm":                                         # model.py:544-548
                s, z = p.qparams[op.outs[0]]
                v = Q(rq.requantize(v.a, v.scale, v.zp, s, z, p.bits), p.bits, s, z)
            env[o] = v
    p.env = env
    res = []
    for o in g.outputs:                                                   # model.py:552-559
        v = env[o]
        if isinstance(v, Q):
            v = q_dequant(v)
        if not isinstance(v, F):
            raise ValueError
        res.append(v.a)
    return (res, times) if profile else res
