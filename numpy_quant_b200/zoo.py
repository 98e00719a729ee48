"""Synthetic graph builders (ONNX-lite protos) for the configs in BASELINE.json.

The reference ships its ViT only as a weight-less topology
(`models/vit/vit_image_classifier_no_weights.onnx`, external data absent) and
builds its small test graphs with the `onnx` package (`models/test.py:19-329`).
Neither is available on the GPU box, so the graphs are rebuilt here node by node
with the same operator vocabulary, attribute values and wiring as the torch-1.13
export, parameterised in batch / size so the
same builder yields ViT-B/16 (config 2) and the small test ViT.

All weights are synthetic: one `default_rng(seed)`
stream drawn in initializer order, weights/biases N(0, 0.02), LayerNorm gamma
1 + N(0, 0.02); no tensor is all-zero or all-negative (the reference's symmetric
scale `2*max/(2^b-1)` would be <= 0, `numpy_quantization.py:15`).
"""
from __future__ import annotations

import numpy as np

from . import onnx_lite as ol

_F32 = np.float32
_I64 = np.int64


class _G:
    """Tiny graph-building helper that keeps node / initializer order."""

    def __init__(self, name: str, seed: int):
        self.g = ol.GraphProto(name=name)
        self.rng = np.random.default_rng(seed)
        self._const_count: dict[str, int] = {}

    def init(self, name: str, shape, kind: str = "w", std: float = 0.02) -> str:
        x = self.rng.normal(size=shape) * std
        if kind == "gamma":
            x = 1.0 + x
        self.g.initializer.append(ol.from_array(x.astype(_F32), name))
        return name

    def node(self, op: str, name: str, inputs, n_out: int = 1, out_names=None, **attrs):
        outs = out_names or [f"{name}_output_{i}" for i in range(n_out)]
        self.g.node.append(ol.make_node(op, inputs, outs, name=name, **attrs))
        return outs[0] if len(outs) == 1 else outs

    def const(self, scope: str, value) -> str:
        """`Constant` node named like torch's exporter: <scope>/Constant[_k]."""
        k = self._const_count.get(scope, 0)
        self._const_count[scope] = k + 1
        name = f"{scope}/Constant" + (f"_{k}" if k else "")
        return self.node("Constant", name, [], value=np.asarray(value))


def vit_graph(batch: int = 1, image_size: int = 224, patch_size: int = 16, hidden: int = 768,
              heads: int = 12, intermediate: int = 3072, layers: int = 12, classes: int = 1000,
              channels: int = 3, seed: int = 0, with_classifier: bool = True) -> ol.ModelProto:
    """ViT image classifier with the topology of the reference's committed export.

    Node sequence per encoder layer (62 nodes) and the embedding prologue (21
    nodes) follow `models/vit/vit_image_classifier_no_weights.onnx`; with the
    defaults this is ViT-B/16: 516 nodes, 200 initializers.
    """
    assert hidden % heads == 0 and image_size % patch_size == 0
    b = _G("vit", seed)
    seq = (image_size // patch_size) ** 2 + 1
    hd = hidden // heads

    # ---- initializers (drawn in this order) --------------------------------
    cls = b.init("vit.embeddings.cls_token", (1, 1, hidden))
    pos = b.init("vit.embeddings.position_embeddings", (1, seq, hidden))
    pw = b.init("vit.embeddings.patch_embeddings.projection.weight", (hidden, channels, patch_size, patch_size))
    pb = b.init("vit.embeddings.patch_embeddings.projection.bias", (hidden,))
    L = []
    for i in range(layers):
        p = f"vit.encoder.layer.{i}."
        L.append(dict(
            qb=b.init(p + "attention.attention.query.bias", (hidden,)),
            kb=b.init(p + "attention.attention.key.bias", (hidden,)),
            vb=b.init(p + "attention.attention.value.bias", (hidden,)),
            ob=b.init(p + "attention.output.dense.bias", (hidden,)),
            ib=b.init(p + "intermediate.dense.bias", (intermediate,)),
            fb=b.init(p + "output.dense.bias", (hidden,)),
            g1=b.init(p + "layernorm_before.weight", (hidden,), "gamma"),
            b1=b.init(p + "layernorm_before.bias", (hidden,)),
            g2=b.init(p + "layernorm_after.weight", (hidden,), "gamma"),
            b2=b.init(p + "layernorm_after.bias", (hidden,)),
        ))
    gf = b.init("vit.layernorm.weight", (hidden,), "gamma")
    bf = b.init("vit.layernorm.bias", (hidden,))
    if with_classifier:
        cw = b.init("classifier.weight", (classes, hidden))
        cb = b.init("classifier.bias", (classes,))
    for i in range(layers):
        L[i].update(
            qw=b.init(f"onnx::MatMul_q{i}", (hidden, hidden)),
            kw=b.init(f"onnx::MatMul_k{i}", (hidden, hidden)),
            vw=b.init(f"onnx::MatMul_v{i}", (hidden, hidden)),
            ow=b.init(f"onnx::MatMul_o{i}", (hidden, hidden)),
            iw=b.init(f"onnx::MatMul_i{i}", (hidden, intermediate)),
            fw=b.init(f"onnx::MatMul_f{i}", (intermediate, hidden)),
        )

    # ---- embeddings ---------------------------------------------------------
    e = "/vit/embeddings"
    pe = e + "/patch_embeddings"
    idx0 = b.const(e, np.array(0, _I64))                                     # reused by the final Gather
    conv = b.node("Conv", pe + "/projection/Conv", ["inputs", pw, pb], dilations=[1, 1], group=1,
                  kernel_shape=[patch_size, patch_size], pads=[0, 0, 0, 0], strides=[patch_size, patch_size])
    shp = b.node("Shape", pe + "/Shape", [conv])
    ax = b.const(pe, np.array([0], _I64))
    st = b.const(pe, np.array([0], _I64))
    en = b.const(pe, np.array([2], _I64))
    sl = b.node("Slice", pe + "/Slice", [shp, st, en, ax])
    m1 = b.const(pe, np.array([-1], _I64))
    cc = b.node("Concat", pe + "/Concat", [sl, m1], axis=0)
    rs = b.node("Reshape", pe + "/Reshape", [conv, cc], allowzero=0)
    tr = b.node("Transpose", pe + "/Transpose", [rs], perm=[0, 2, 1])
    eshape = b.const(e, np.array([batch, -1, -1], _I64))
    three = b.const(e, np.array([3], _I64))
    ones = b.node("ConstantOfShape", e + "/ConstantOfShape", [three], value=np.array([1], _I64))
    neg1 = b.const(e, np.array(-1, _I64))
    mul = b.node("Mul", e + "/Mul", [ones, neg1])
    eq = b.node("Equal", e + "/Equal", [eshape, mul])
    wh = b.node("Where", e + "/Where", [eq, ones, eshape])
    ex = b.node("Expand", e + "/Expand", [cls, wh])
    cat = b.node("Concat", e + "/Concat", [ex, tr], axis=1)
    x = b.node("Add", e + "/Add", [cat, pos])

    eps = float(np.float32(1e-12))

    # ---- encoder ------------------------------------------------------------
    for i in range(layers):
        w = L[i]
        s = f"/vit/encoder/layer.{i}"
        a = s + "/attention/attention"
        ln1 = b.node("LayerNormalization", s + "/layernorm_before/LayerNormalization", [x, w["g1"], w["b1"]],
                     axis=-1, epsilon=eps)
        q = b.node("MatMul", a + "/query/MatMul", [ln1, w["qw"]])
        q = b.node("Add", a + "/query/Add", [w["qb"], q])
        k = b.node("MatMul", a + "/key/MatMul", [ln1, w["kw"]])
        k = b.node("Add", a + "/key/Add", [w["kb"], k])
        split = [batch, seq, heads, hd]
        k4 = b.node("Reshape", a + "/Reshape", [k, b.const(a, np.array(split, _I64))], allowzero=0)
        v = b.node("MatMul", a + "/value/MatMul", [ln1, w["vw"]])
        v = b.node("Add", a + "/value/Add", [w["vb"], v])
        v4 = b.node("Reshape", a + "/Reshape_1", [v, b.const(a, np.array(split, _I64))], allowzero=0)
        vt = b.node("Transpose", a + "/Transpose", [v4], perm=[0, 2, 1, 3])
        q4 = b.node("Reshape", a + "/Reshape_2", [q, b.const(a, np.array(split, _I64))], allowzero=0)
        qt = b.node("Transpose", a + "/Transpose_1", [q4], perm=[0, 2, 1, 3])
        kt = b.node("Transpose", a + "/Transpose_2", [k4], perm=[0, 2, 3, 1])
        sc = b.node("MatMul", a + "/MatMul", [qt, kt])
        sc = b.node("Div", a + "/Div", [sc, b.const(a, np.array(np.sqrt(hd), _F32))])
        pr = b.node("Softmax", a + "/Softmax", [sc], axis=-1)
        ctx = b.node("MatMul", a + "/MatMul_1", [pr, vt])
        ctx = b.node("Transpose", a + "/Transpose_3", [ctx], perm=[0, 2, 1, 3])
        ctx = b.node("Reshape", a + "/Reshape_3", [ctx, b.const(a, np.array([batch, seq, hidden], _I64))],
                     allowzero=0)
        o = b.node("MatMul", s + "/attention/output/dense/MatMul", [ctx, w["ow"]])
        o = b.node("Add", s + "/attention/output/dense/Add", [w["ob"], o])
        x1 = b.node("Add", s + "/Add", [o, x])
        ln2 = b.node("LayerNormalization", s + "/layernorm_after/LayerNormalization", [x1, w["g2"], w["b2"]],
                     axis=-1, epsilon=eps)
        h = b.node("MatMul", s + "/intermediate/dense/MatMul", [ln2, w["iw"]])
        h = b.node("Add", s + "/intermediate/dense/Add", [w["ib"], h])
        act = s + "/intermediate/intermediate_act_fn"
        d = b.node("Div", act + "/Div", [h, b.const(act, np.array(1.4142135381698608, _F32))])
        er = b.node("Erf", act + "/Erf", [d])
        ad = b.node("Add", act + "/Add", [er, b.const(act, np.array(1.0, _F32))])
        mu = b.node("Mul", act + "/Mul", [h, ad])
        ge = b.node("Mul", act + "/Mul_1", [mu, b.const(act, np.array(0.5, _F32))])
        f = b.node("MatMul", s + "/output/dense/MatMul", [ge, w["fw"]])
        f = b.node("Add", s + "/output/dense/Add", [w["fb"], f])
        x = b.node("Add", s + "/output/Add", [f, x1])

    ln = b.node("LayerNormalization", "/vit/layernorm/LayerNormalization", [x, gf, bf], axis=-1, epsilon=eps)
    if with_classifier:
        tok = b.node("Gather", "/Gather", [ln, idx0], axis=1)
        b.node("Gemm", "/classifier/Gemm", [tok, cw, cb], out_names=["logits"], alpha=1.0, beta=1.0, transB=1)
        out = ol.ValueInfoProto("logits", ol.FLOAT, [batch, classes])
    else:
        b.g.node[-1].output[0] = "last_hidden_state"
        out = ol.ValueInfoProto("last_hidden_state", ol.FLOAT, [batch, seq, hidden])
    b.g.input.append(ol.ValueInfoProto("inputs", ol.FLOAT, [batch, channels, image_size, image_size]))
    b.g.output.append(out)
    return ol.ModelProto(graph=b.g, opset=17, producer_name="numpy_quant_b200.zoo")


def gemm_graph(k: int, m: int, n: int, seed: int = 0) -> ol.ModelProto:
    """Single `Gemm` node, input [k, m] x weight [m, n] + bias [n] (cf. `models/test.py:19-57`)."""
    rng = np.random.default_rng(seed)
    g = ol.GraphProto(name="Gemm")
    g.initializer.append(ol.from_array(rng.normal(size=(m, n)).astype(_F32), "weight"))
    g.initializer.append(ol.from_array(rng.normal(size=n).astype(_F32), "bias"))
    g.node.append(ol.make_node("Gemm", ["input", "weight", "bias"], ["output"], name="Gemm"))
    g.input.append(ol.ValueInfoProto("input", ol.FLOAT, [k, m]))
    g.output.append(ol.ValueInfoProto("output", ol.FLOAT, [k, n]))
    return ol.ModelProto(graph=g, opset=13)


def matmul_graph(a_shape, b_shape) -> ol.ModelProto:
    """Single `MatMul` node over two graph inputs (cf. `models/test.py:60-96`)."""
    out_shape = tuple(np.broadcast_shapes(a_shape[:-2], b_shape[:-2])) + (a_shape[-2], b_shape[-1])
    g = ol.GraphProto(name="MatMul")
    g.node.append(ol.make_node("MatMul", ["input_a", "input_b"], ["output"], name="MatMul"))
    g.input.append(ol.ValueInfoProto("input_a", ol.FLOAT, list(a_shape)))
    g.input.append(ol.ValueInfoProto("input_b", ol.FLOAT, list(b_shape)))
    g.output.append(ol.ValueInfoProto("output", ol.FLOAT, list(out_shape)))
    return ol.ModelProto(graph=g, opset=13)


def conv_graph(batch: int, channels: int, inp_hw, out_channels: int, kernel, pads, strides,
               seed: int = 0) -> ol.ModelProto:
    """Single `Conv` node with weight + bias initializers (cf. `models/test.py:99-151`)."""
    rng = np.random.default_rng(seed)
    oh = (inp_hw[0] - kernel[0] + pads[0] + pads[2]) // strides[0] + 1
    ow = (inp_hw[1] - kernel[1] + pads[1] + pads[3]) // strides[1] + 1
    g = ol.GraphProto(name="Conv")
    g.initializer.append(ol.from_array(rng.normal(size=(out_channels, channels, *kernel)).astype(_F32), "weight"))
    g.initializer.append(ol.from_array(rng.normal(size=out_channels).astype(_F32), "bias"))
    g.node.append(ol.make_node("Conv", ["input", "weight", "bias"], ["output"], name="Conv",
                               kernel_shape=list(kernel), pads=list(pads), strides=list(strides)))
    g.input.append(ol.ValueInfoProto("input", ol.FLOAT, [batch, channels, *inp_hw]))
    g.output.append(ol.ValueInfoProto("output", ol.FLOAT, [batch, out_channels, oh, ow]))
    return ol.ModelProto(graph=g, opset=13)


def mlp_graph(widths=(2, 5, 2), seed: int = 0, std: float = 2.0) -> ol.ModelProto:
    """Gemm -> Relu -> Gemm -> Sigmoid MLP with the wiring of the reference's `models/mlp.onnx`."""
    rng = np.random.default_rng(seed)
    g = ol.GraphProto(name="mlp")
    cur = "input"
    n_fc = len(widths) - 1
    for i in range(n_fc):
        w = (rng.normal(size=(widths[i + 1], widths[i])) * std).astype(_F32)
        w.flat[0] = abs(w.flat[0]) + 0.5                      # keep max(w) > 0 for the symmetric scale
        bias = (rng.normal(size=widths[i + 1]) * std).astype(_F32)
        bias[0] = abs(bias[0]) + 0.5
        g.initializer.append(ol.from_array(w, f"fc{i + 1}.weight"))
        g.initializer.append(ol.from_array(bias, f"fc{i + 1}.bias"))
        last = i == n_fc - 1
        out = f"/fc{i + 1}/Gemm_output_0"
        g.node.append(ol.make_node("Gemm", [cur, f"fc{i + 1}.weight", f"fc{i + 1}.bias"], [out],
                                   name=f"/fc{i + 1}/Gemm", alpha=1.0, beta=1.0, transB=1))
        if last:
            g.node.append(ol.make_node("Sigmoid", [out], ["output"], name="/sigmoid/Sigmoid"))
        else:
            cur = f"/relu{i + 1}/Relu_output_0"
            g.node.append(ol.make_node("Relu", [out], [cur], name=f"/relu{i + 1}/Relu"))
    g.input.append(ol.ValueInfoProto("input", ol.FLOAT, ["batch_size", widths[0]]))
    g.output.append(ol.ValueInfoProto("output", ol.FLOAT, ["batch_size", widths[-1]]))
    return ol.ModelProto(graph=g, opset=10)
