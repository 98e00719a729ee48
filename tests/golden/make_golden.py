#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (`/root/reference/numpy_quant`) in the build container.

The reference imports the `onnx` package (model.py:8-10), which is not
installed; a stand-in module backed by `numpy_quant_b200.onnx_lite` is put in
`sys.modules` first (it only has to read protos -- no arithmetic lives there).
Run from the repo root:   python tests/golden/make_golden.py
The GPU box has no /root/reference, so the outputs (small .npz files) are
committed together with this script.
"""
import os
import shutil
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
# output directory: tests/golden/ by default; `python make_golden.py <dir>` writes elsewhere (the regeneration test)
OUT = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else HERE
sys.path.insert(0, ROOT)

from numpy_quant_b200 import onnx_lite as ol, zoo  # noqa: E402


def install_onnx_stub():
    onnx = types.ModuleType("onnx")
    onnx.TensorProto = ol.TensorProto
    onnx.ModelProto = ol.ModelProto
    onnx.mapping = types.ModuleType("onnx.mapping")
    onnx.numpy_helper = types.ModuleType("onnx.numpy_helper")
    onnx.numpy_helper.to_array = ol.to_array
    onnx.helper = types.ModuleType("onnx.helper")
    onnx.helper.get_attribute_value = ol.get_attribute_value
    for name in ("onnx", "onnx.mapping", "onnx.numpy_helper", "onnx.helper"):
        sys.modules[name] = onnx if name == "onnx" else getattr(onnx, name.split(".")[1])


install_onnx_stub()
sys.path.insert(0, REF)
from numpy_quant.model import Model, Constant  # noqa: E402  (the reference)
from numpy_quant.numpy_quantization import (quant_parameters, quantize, dequantize, q_matmul,  # noqa: E402
                                            requantize)
from numpy_quant.tensor import FTensor, QTensor, quantize_tensor_min_max, tensor_min_max, fconv2d  # noqa: E402
from numpy_quant import numpy_helper  # noqa: E402

warnings.simplefilter("ignore")
I64 = np.int64


def zp_arr(zp):
    """(flag, value) encoding of an optional zero-point for npz storage."""
    return (np.array(0 if zp is None else 1, I64), np.array(0 if zp is None else zp, I64))


def kernel_vectors():
    """quantize / dequantize / requantize / q_matmul on seeded inputs, every bit width."""
    rng = np.random.default_rng(1234)
    out = {}
    x = rng.normal(size=(37, 53)).astype(np.float32) * 3.0
    x[0, :8] = [0.0, -0.0, 1e-30, -1e-30, 1e30, -1e30, 2.5, -2.5]
    out["x"] = x
    for bits in range(2, 9):
        for asym in (False, True):
            tag = f"b{bits}_{'a' if asym else 's'}"
            scale, zp = quant_parameters(x.min(), x.max(), bits, asym)
            # half-integer multiples of the scale exercise round-half-even and the f64 zero-point add
            ties = (np.arange(-40, 40, dtype=np.float32) + 0.5) * np.float32(scale)
            xin = np.concatenate([x.ravel(), ties]).astype(np.float32)
            q = quantize(xin, bits, scale, zp)
            out[f"q_in_{tag}"] = xin
            out[f"q_scale_{tag}"] = np.asarray(scale)
            out[f"q_zpf_{tag}"], out[f"q_zp_{tag}"] = zp_arr(zp)
            out[f"q_out_{tag}"] = q
            out[f"dq_out_{tag}"] = dequantize(q, scale, zp)
    # wide accumulators with array zero-points -> dequantize / requantize
    acc = rng.integers(-2 ** 27, 2 ** 27, size=(29, 31)).astype(I64)
    acc[0, :4] = [2 ** 31 - 1, -2 ** 31, 2 ** 24 + 1, -(2 ** 24) - 1]
    zrow = rng.integers(-50000, 50000, size=(29, 1)).astype(I64)
    zcol = rng.integers(-50000, 50000, size=(1, 31)).astype(I64)
    sc = np.float32(3.1e-7)
    out["acc"], out["acc_zrow"], out["acc_zcol"], out["acc_scale"] = acc, zrow, zcol, np.asarray(sc)
    out["acc_dq_none"] = dequantize(acc, sc, None)
    out["acc_dq_row"] = dequantize(acc, sc, zrow)
    out["acc_dq_full"] = dequantize(acc, sc, zrow + zcol - 77)
    for bits in range(2, 9):
        for asym in (False, True):
            tag = f"b{bits}_{'a' if asym else 's'}"
            d = dequantize(acc, sc, zrow + zcol - 77)
            s_out, zp_out = quant_parameters(d.min(), d.max(), bits, asym)
            out[f"rq_scale_{tag}"] = np.asarray(s_out)
            out[f"rq_zpf_{tag}"], out[f"rq_zp_{tag}"] = zp_arr(zp_out)
            out[f"rq_out_{tag}"] = requantize(acc, sc, zrow + zcol - 77, s_out, zp_out, bits)
    # q_matmul: 4 sym/asym combos, broadcast batch dims (test_quantization.py:70-86)
    for bits in (2, 4, 8):
        lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
        a = rng.integers(lo, hi + 1, size=(2, 1, 9, 13)).astype(I64)
        b = rng.integers(lo, hi + 1, size=(1, 3, 13, 7)).astype(I64)
        out[f"mm_a_b{bits}"], out[f"mm_b_b{bits}"] = a, b
        for za, zb in ((None, None), (None, I64(-3)), (I64(5), None), (I64(5), I64(-3))):
            tag = f"b{bits}_{'n' if za is None else 'z'}{'n' if zb is None else 'z'}"
            y, s, z = q_matmul(a, np.float32(0.02), za, b, np.float32(0.5), zb)
            out[f"mm_acc_{tag}"] = y
            out[f"mm_scale_{tag}"] = np.asarray(s)
            out[f"mm_zpf_{tag}"] = np.array(0 if z is None else 1, I64)
            out[f"mm_zp_{tag}"] = np.array(0, I64) if z is None else np.broadcast_to(z, y.shape).copy()
    # erf polynomial + float conv (fake-quant Conv path)
    xe = np.linspace(-6, 6, 4001).astype(np.float32)
    out["erf_in"], out["erf_out"] = xe, numpy_helper.erf(xe)
    cx = rng.normal(size=(2, 3, 9, 10)).astype(np.float32)
    cw = rng.normal(size=(2, 3, 3, 2)).astype(np.float32)
    cb = rng.normal(size=2).astype(np.float32)
    out["conv_x"], out["conv_w"], out["conv_b"] = cx, cw, cb
    out["conv_y"] = fconv2d(FTensor(cx), FTensor(cw), FTensor(cb), (0, 2, 2, 1), (2, 1)).data
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **out)
    print("kernels.npz", len(out), "arrays")


def ka1():
    """KA-1: literal case of test_quantization.py:44-52 incl. requantize."""
    w = np.array([[1.3, 5.0, -0.3], [2.1, -3.4, -0.1], [-0.4, 4.0, 1.7]], np.float32)
    x = np.array([[2.2], [2.1], [-2.0]], np.float32)
    y = FTensor(w).matmul(FTensor(x))
    ys, yz = quant_parameters(*tensor_min_max(y), bit_width=8, asymmetric=True)
    out = {"w": w, "x": x, "y_scale": np.asarray(ys), "y_zp": np.asarray(yz)}
    for wa in (False, True):
        for xa in (False, True):
            tag = f"{int(wa)}{int(xa)}"
            qw = quantize_tensor_min_max(FTensor(w), 8, wa)
            qx = quantize_tensor_min_max(FTensor(x), 8, xa)
            r = qw.matmul(qx)
            out[f"qw_{tag}"], out[f"qx_{tag}"] = qw.data, qx.data
            out[f"ws_{tag}"], out[f"xs_{tag}"] = np.asarray(qw.scale), np.asarray(qx.scale)
            out[f"wzf_{tag}"], out[f"wz_{tag}"] = zp_arr(qw.zero_point)
            out[f"xzf_{tag}"], out[f"xz_{tag}"] = zp_arr(qx.zero_point)
            out[f"acc_{tag}"] = r.data
            out[f"accs_{tag}"] = np.asarray(r.scale)
            out[f"acczf_{tag}"] = np.array(0 if r.zero_point is None else 1, I64)
            out[f"accz_{tag}"] = np.array(0, I64) if r.zero_point is None else np.asarray(r.zero_point)
            out[f"rq_{tag}"] = r.requantize(8, ys, yz).data
            out[f"dq_{tag}"] = r.dequantize().data
    np.savez_compressed(os.path.join(OUT, "ka1.npz"), **out)
    print("ka1.npz")


def dump_qmodel(prefix, out, model, qmodel, inputs, keep_values=()):
    """Store every quantization parameter, every quantized constant and the outputs."""
    res = qmodel(inputs)
    fres = model(inputs)
    for i, (r, f) in enumerate(zip(res, fres)):
        out[f"{prefix}/out{i}"] = r
        out[f"{prefix}/fout{i}"] = f
    names = []
    for name, p in qmodel.quant_params.items():
        names.append(name)
        out[f"{prefix}/qp_scale/{name}"] = np.asarray(p.scale, dtype=np.float64)   # may be f64 for ITensor stats
        out[f"{prefix}/qp_zpf/{name}"], out[f"{prefix}/qp_zp/{name}"] = zp_arr(p.zero_point)
    out[f"{prefix}/qp_names"] = np.array(names)
    for v in qmodel.values:
        if isinstance(v, Constant) and v.data.data.size <= 4096:
            out[f"{prefix}/qconst/{v.name}"] = v.data.data
            out[f"{prefix}/qconst_bits/{v.name}"] = np.array(v.data.bit_width, I64)
    by_name = {v.name: v for v in qmodel.values}
    for name in keep_values:
        d = by_name[name].data
        if isinstance(d, QTensor):
            out[f"{prefix}/val_q/{name}"] = d.data
            out[f"{prefix}/val_qbits/{name}"] = np.array(d.bit_width, I64)
            out[f"{prefix}/val_qscale/{name}"] = np.asarray(d.scale)
            if d.zero_point is not None:
                out[f"{prefix}/val_qzp/{name}"] = np.broadcast_to(d.zero_point, d.data.shape).copy()
        else:
            out[f"{prefix}/val_f/{name}"] = d.data


def graphs():
    out = {}
    # ---- config 1: the reference's own models/mlp.onnx (KA-2) -------------------
    shutil.copyfile(os.path.join(REF, "models", "mlp.onnx"), os.path.join(OUT, "mlp.onnx"))
    proto = ol.load(os.path.join(OUT, "mlp.onnx"))
    X = np.array([[0.5, -0.25], [-1.0, 0.75], [0.1, 0.9]], np.float32)
    out["mlp/x"] = X
    for bits in (2, 3, 4, 5, 6, 7, 8):
        model = Model.from_onnx(proto)
        q = model.quantize([X], bit_width=bits)
        dump_qmodel(f"mlp/b{bits}", out, model, q, [X],
                    keep_values=["input", "/fc1/Gemm_output_0", "/relu/Relu_output_0", "/fc2/Gemm_output_0"])
    # ---- Gemm / MatMul / Conv graphs (test_quantization.py:151-188) --------------
    rng = np.random.default_rng(0)
    gp = zoo.gemm_graph(3, 4, 2, seed=0)
    gi = rng.normal(size=(3, 4)).astype(np.float32)
    out["gemm/x"] = gi
    model = Model.from_onnx(gp)
    dump_qmodel("gemm/b8", out, model, model.quantize([gi], 8), [gi], keep_values=["output"])
    rng = np.random.default_rng(0)
    a = rng.normal(size=(1, 2, 3, 4)).astype(np.float32)
    b = rng.normal(size=(2, 1, 4, 3)).astype(np.float32)
    out["matmul/a"], out["matmul/b"] = a, b
    model = Model.from_onnx(zoo.matmul_graph(a.shape, b.shape))
    dump_qmodel("matmul/b8", out, model, model.quantize([a, b], 8), [a, b], keep_values=["output"])
    cp = zoo.conv_graph(2, 3, (9, 10), 2, (3, 2), (0, 2, 2, 1), (2, 1), seed=0)
    ci = np.random.default_rng(0).normal(size=(2, 3, 9, 10)).astype(np.float32)
    out["conv/x"] = ci
    model = Model.from_onnx(cp)
    dump_qmodel("conv/b8", out, model, model.quantize([ci], 8), [ci])
    # ---- small ViT with the committed ViT-B topology ------------------------------
    cfg = dict(batch=2, image_size=32, patch_size=16, hidden=32, heads=4, intermediate=64, layers=2, classes=10)
    vp = zoo.vit_graph(seed=0, **cfg)
    vi = np.random.default_rng(1).normal(size=(2, 3, 32, 32)).astype(np.float32)
    out["vit/x"] = vi
    a0 = "/vit/encoder/layer.0/attention/attention"
    keep = ["/vit/embeddings/Add_output_0", "/vit/encoder/layer.0/layernorm_before/LayerNormalization_output_0",
            a0 + "/query/MatMul_output_0", a0 + "/MatMul_output_0", a0 + "/Softmax_output_0",
            a0 + "/MatMul_1_output_0", "/vit/encoder/layer.0/intermediate/intermediate_act_fn/Mul_1_output_0",
            "/vit/encoder/layer.1/output/Add_output_0", "logits"]
    for bits in (8, 4, 2):
        model = Model.from_onnx(vp)
        dump_qmodel(f"vit/b{bits}", out, model, model.quantize([vi], bits), [vi], keep_values=keep)
    np.savez_compressed(os.path.join(OUT, "graphs.npz"), **out)
    print("graphs.npz", len(out), "arrays")


def topology_check():
    """The zoo builder must reproduce the committed ViT-B export node for node."""
    ref = ol.load(os.path.join(REF, "models/vit/vit_image_classifier_no_weights.onnx"), load_external=False)   # the .data sidecar is not shipped
    mine = zoo.vit_graph()

    def val(a):
        v = ol.get_attribute_value(a)
        return ol.to_array(v).tolist() if isinstance(v, ol.TensorProto) else v

    assert len(ref.graph.node) == len(mine.graph.node) == 516
    for r, m in zip(ref.graph.node, mine.graph.node):
        assert (r.op_type, r.name, r.output) == (m.op_type, m.name, m.output), (r.name, m.name)
        assert all(x == y or x.startswith("onnx::MatMul") for x, y in zip(r.input, m.input))
        assert {a.name: val(a) for a in r.attribute} == {a.name: val(a) for a in m.attribute}, r.name
    print("zoo.vit_graph() == committed ViT-B/16 topology (516 nodes, 200 initializers)")


if __name__ == "__main__":
    topology_check()
    kernel_vectors()
    ka1()
    graphs()
