"""End-to-end parity of the drop-in API (Model.from_onnx / quantize / __call__) on the GPU
against the golden vectors of the unmodified reference and against the oracle."""
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from numpy_quant_b200 import onnx_lite as ol, zoo  # noqa: E402
from numpy_quant_b200.model import Constant, Model, QModel, QuantizationParams  # noqa: E402
from numpy_quant_b200.tensor import FTensor, QTensor, quantize_tensor, quantize_tensor_min_max, tensor_min_max  # noqa: E402
from numpy_quant_b200.numpy_quantization import quant_parameters  # noqa: E402
from oracle import ref_graph as rg, ref_quant as rq  # noqa: E402

warnings.simplefilter("ignore")
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gv():
    return np.load(os.path.join(G, "graphs.npz"))


def inject_oracle_params(qmodel: QModel, fmodel: Model, plan: rg.QPlan):
    """Give the product model the oracle's exact quantization parameters (parity
    'using the same qparams'): calibration min/max come from a float pass whose summation order
    differs by ulps, which is a float (1e-5) matter, not part of the integer contract."""
    fconst = {v.name: v for v in fmodel.values if isinstance(v, Constant)}
    for name, (s, z) in plan.qparams.items():
        qmodel.quant_params[name] = QuantizationParams(s, z)
    for v in qmodel.values:
        if isinstance(v, Constant):
            oq = plan.qconsts[v.name]
            v.data = quantize_tensor(fconst[v.name].data, oq.bits, oq.scale, oq.zp)
            np.testing.assert_array_equal(v.data.data, oq.a, err_msg=v.name)
    qmodel._const_deq.clear()


def check_qparams_close(qmodel, plan, rtol=2e-5):
    for name, (s, z) in plan.qparams.items():
        p = qmodel.quant_params[name]
        if not np.isfinite(s) or s == 0:
            continue
        np.testing.assert_allclose(np.float64(p.scale), np.float64(s), rtol=rtol, err_msg=name)
        assert (p.zero_point is None) == (z is None), name
        if z is not None:
            assert abs(int(p.zero_point) - int(z)) <= 1, name


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 7, 8])
def test_mlp_reference_onnx_file(gv, bits):
    proto = ol.load(os.path.join(G, "mlp.onnx"))
    x = gv["mlp/x"]
    model = Model.from_onnx(proto)
    fout = model([x])[0]
    np.testing.assert_allclose(fout, gv[f"mlp/b{bits}/fout0"], rtol=1e-5, atol=1e-7)
    q = model.quantize([x], bit_width=bits)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], bits)
    check_qparams_close(q, plan)
    inject_oracle_params(q, model, plan)
    out = q([x])[0]
    by = {v.name: v for v in q.values}
    pre = f"mlp/b{bits}"
    # integer values: bit-exact
    np.testing.assert_array_equal(by["input"].data.data, gv[f"{pre}/val_q/input"])
    np.testing.assert_array_equal(by["/fc1/Gemm_output_0"].data.data, gv[f"{pre}/val_q//fc1/Gemm_output_0"])
    np.testing.assert_array_equal(by["/fc2/Gemm_output_0"].data.data, gv[f"{pre}/val_q//fc2/Gemm_output_0"])
    assert by["/fc1/Gemm_output_0"].data.bit_width == bits
    np.testing.assert_array_equal(by["/relu/Relu_output_0"].data.data, gv[f"{pre}/val_f//relu/Relu_output_0"])
    np.testing.assert_allclose(out, gv[f"{pre}/out0"], rtol=1e-5, atol=1e-7)


def test_gemm_matmul_conv_graphs(gv):
    proto = zoo.gemm_graph(3, 4, 2, seed=0)
    x = gv["gemm/x"]
    m = Model.from_onnx(proto)
    q = m.quantize([x], 8)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], 8)
    check_qparams_close(q, plan)
    inject_oracle_params(q, m, plan)
    np.testing.assert_array_equal(q([x])[0], gv["gemm/b8/out0"])           # int path end to end: exact
    np.testing.assert_array_equal({v.name: v for v in q.values}["output"].data.data, gv["gemm/b8/val_q/output"])

    a, b = gv["matmul/a"], gv["matmul/b"]
    proto = zoo.matmul_graph(a.shape, b.shape)
    m = Model.from_onnx(proto)
    q = m.quantize([a, b], 8)
    plan = rg.calibrate(rg.import_graph(proto, ol), [a, b], 8)
    inject_oracle_params(q, m, plan)
    out = q([a, b])[0]
    np.testing.assert_array_equal(out, gv["matmul/b8/out0"])
    acc = {v.name: v for v in q.values}["output"].data
    np.testing.assert_array_equal(acc.data, gv["matmul/b8/val_q/output"])                 # raw accumulator
    np.testing.assert_array_equal(np.broadcast_to(acc.zero_point, acc.data.shape), gv["matmul/b8/val_qzp/output"])
    assert acc.bit_width == 32

    proto = zoo.conv_graph(2, 3, (9, 10), 2, (3, 2), (0, 2, 2, 1), (2, 1), seed=0)
    x = gv["conv/x"]
    m = Model.from_onnx(proto)
    np.testing.assert_allclose(m([x])[0], gv["conv/b8/fout0"], rtol=1e-5, atol=1e-5)
    q = m.quantize([x], 8)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], 8)
    inject_oracle_params(q, m, plan)
    # integer im2col conv == the reference's fake-quant float conv up to float32 summation rounding
    np.testing.assert_allclose(q([x])[0], gv["conv/b8/out0"], rtol=1e-5, atol=2e-5)


def test_conv_block_config3_geometry_vs_oracle():
    """BASELINE config 3 (Conv2d block, test_conv2d geometry scaled up: asymmetric pads (0,2,2,1), strides (2,1),
    kernel (3,2), 64 -> 128 channels, 57x58 inputs) at a batch the oracle finishes in seconds: the integer
    im2col qGEMM path equals the reference's fake-quant float conv within float32 summation rounding, and the
    calibrated quantization parameters are the oracle's."""
    proto = zoo.conv_graph(4, 64, (57, 58), 128, (3, 2), (0, 2, 2, 1), (2, 1), seed=0)
    x = np.random.default_rng(0).normal(size=(4, 64, 57, 58)).astype(np.float32)
    m = Model.from_onnx(proto)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], 8)
    fref = rg.run_float(rg.import_graph(proto, ol), [x])[0] if hasattr(rg, "run_float") else None
    fout = m([x])[0]
    assert fout.shape == (4, 128, 29, 60)
    if fref is not None:
        np.testing.assert_allclose(fout, fref, rtol=1e-4, atol=1e-3)
    q = m.quantize([x], 8)
    check_qparams_close(q, plan, rtol=1e-4)
    inject_oracle_params(q, m, plan)
    want = rg.run_quant(plan, [x])[0]
    got = q([x])[0]
    # K = 384 products of magnitude ~1e2 in float32 (reference) vs exact integers (here): summation rounding only
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-3)
    np.testing.assert_array_equal(q([x], retain=False)[0], got)


@pytest.mark.parametrize("bits", [8, 4])
def test_patch_embedding_conv_is_a_reshape(bits, monkeypatch):
    """Conv with kernel == stride and no padding (the ViT patch embedding, C = 3): the input is quantized straight
    into the patch matrix (nq_quantize_patches_f32) and the Conv runs as a MatMul against the filters in their natural
    order -- no im2col pass.  Quantized input codes are the oracle's bit for bit, the output equals the reference's
    fake-quant float conv within float32 summation rounding, and it is bit-identical to the im2col route."""
    from numpy_quant_b200 import kernels as K
    proto = zoo.conv_graph(3, 3, (32, 48), 40, (8, 16), (0, 0, 0, 0), (8, 16), seed=1)
    x = np.random.default_rng(2).normal(size=(3, 3, 32, 48)).astype(np.float32)
    m = Model.from_onnx(proto)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], bits)
    q = m.quantize([x], bits)
    inject_oracle_params(q, m, plan)
    before = K.LAUNCHES
    got = q([x])[0]
    launches = K.LAUNCHES - before
    want = rg.run_quant(plan, [x])[0]
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-3)
    xin = {v.name: v for v in q.values}["input"].data
    assert xin._patches is not None                                       # the new route ran
    np.testing.assert_array_equal(xin.data, plan_quantized_input(plan, x, bits))
    np.testing.assert_array_equal(q([x], retain=False)[0], got)
    monkeypatch.setattr(K, "can_quantize_patches", lambda *a, **k: False)  # the nq_im2col + nq_qgemm_s8 route
    before = K.LAUNCHES
    ref = q([x])[0]
    assert K.LAUNCHES - before > launches                                 # one pass (im2col) more
    np.testing.assert_array_equal(got, ref)


def plan_quantized_input(plan, x, bits):
    s, z = plan.qparams["input"]
    from oracle import ref_quant as rq
    return rq.quantize(x, bits, s, z)


VIT_CFG = dict(batch=2, image_size=32, patch_size=16, hidden=32, heads=4, intermediate=64, layers=2, classes=10)


def test_streaming_calibration_gives_the_same_parameters():
    """quantize(keep_values=False) -- statistics reduced on the fly, activations freed after their last consumer --
    derives bit-identical quantization parameters and the same quantized outputs as the reference-like pass that
    keeps every `Value.data` readable."""
    x = np.random.default_rng(5).normal(size=(2, 3, 32, 32)).astype(np.float32)
    qa = Model.from_onnx(zoo.vit_graph(seed=0, **VIT_CFG)).quantize([x], bit_width=8)
    mb = Model.from_onnx(zoo.vit_graph(seed=0, **VIT_CFG))
    qb = mb.quantize([x], bit_width=8, keep_values=False)
    assert set(qa.quant_params) == set(qb.quant_params)
    for name, pa in qa.quant_params.items():
        pb = qb.quant_params[name]
        assert np.float32(pa.scale) == np.float32(pb.scale), name
        assert (pa.zero_point is None) == (pb.zero_point is None) and (pa.zero_point is None or int(pa.zero_point) == int(pb.zero_point)), name
    freed = [v for v in mb.values if v.data is None]
    assert len(freed) > 50                                    # intermediates were released
    np.testing.assert_array_equal(qa([x])[0], qb([x])[0])
    # conv graph (model input feeding a Conv): same thing
    proto = zoo.conv_graph(2, 64, (9, 10), 16, (3, 2), (0, 2, 2, 1), (2, 1), seed=0)
    xc = np.random.default_rng(6).normal(size=(2, 64, 9, 10)).astype(np.float32)
    ya = Model.from_onnx(proto).quantize([xc], 8)([xc])[0]
    yb = Model.from_onnx(proto).quantize([xc], 8, keep_values=False)([xc])[0]
    np.testing.assert_array_equal(ya, yb)


@pytest.mark.parametrize("bits", [8, 4, 2])
def test_small_vit(gv, bits):
    proto = zoo.vit_graph(seed=0, **VIT_CFG)
    x = gv["vit/x"]
    model = Model.from_onnx(proto)
    fout = model([x])[0]
    np.testing.assert_allclose(fout, gv[f"vit/b{bits}/fout0"], rtol=1e-4, atol=1e-5)
    q = model.quantize([x], bit_width=bits)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], bits)
    check_qparams_close(q, plan, rtol=1e-4)
    inject_oracle_params(q, model, plan)
    out = q([x])[0]
    by = {v.name: v for v in q.values}
    pre = f"vit/b{bits}"
    np.testing.assert_array_equal(by["inputs"].data.data, rq.quantize(x, bits, *plan.qparams["inputs"]))
    # first encoder inputs are float-parity values
    emb = "/vit/embeddings/Add_output_0"
    np.testing.assert_allclose(by[emb].data.data, gv[f"{pre}/val_f/{emb}"], rtol=1e-5, atol=1e-5)
    ln = "/vit/encoder/layer.0/layernorm_before/LayerNormalization_output_0"
    np.testing.assert_allclose(by[ln].data.data, gv[f"{pre}/val_f/{ln}"], rtol=1e-4, atol=1e-5)
    # integer accumulators downstream: identical except where a float input sat within an ulp of a
    # rounding boundary (bounded fraction, |difference| explained by single-code flips)
    a0 = "/vit/encoder/layer.0/attention/attention"
    qm = by[a0 + "/query/MatMul_output_0"].data
    ref = gv[f"{pre}/val_q/{a0}/query/MatMul_output_0"]
    assert qm.bit_width == 4 * bits
    mism = np.mean(qm.data != ref)
    assert mism < 0.02, f"query accumulator mismatch fraction {mism}"
    # TEACHER-FORCED against the reference's own layer-0 intermediates (golden vectors of the unmodified reference):
    # integer stages fed the reference's values must reproduce the reference's accumulators bit for bit
    rg.run_quant(plan, [x])
    env = plan.env
    consts = {v.name: v.data for v in q.values if isinstance(v, Constant)}
    ln_codes = rq.quantize(gv[f"{pre}/val_f/{ln}"], bits, *plan.qparams[ln])
    acc_q = QTensor(ln_codes, bits, *plan.qparams[ln]).matmul(consts["onnx::MatMul_q0"])
    np.testing.assert_array_equal(acc_q.data, ref)                                       # query accumulator: exact
    np.testing.assert_array_equal(np.broadcast_to(acc_q.zero_point, ref.shape), gv[f"{pre}/val_qzp/{a0}/query/MatMul_output_0"])
    qn, kn, vn, pn = (a0 + "/Transpose_1_output_0", a0 + "/Transpose_2_output_0", a0 + "/Transpose_output_0", a0 + "/Softmax_output_0")
    qt = QTensor(rq.quantize(env[qn].a, bits, *plan.qparams[qn]), bits, *plan.qparams[qn])
    kt = QTensor(rq.quantize(env[kn].a, bits, *plan.qparams[kn]), bits, *plan.qparams[kn])
    sc = qt.matmul(kt)
    np.testing.assert_array_equal(sc.data, gv[f"{pre}/val_q/{a0}/MatMul_output_0"])     # score accumulator: exact
    np.testing.assert_array_equal(np.broadcast_to(sc.zero_point, sc.data.shape), np.broadcast_to(gv[f"{pre}/val_qzp/{a0}/MatMul_output_0"], sc.data.shape))
    pt = QTensor(rq.quantize(gv[f"{pre}/val_f/{pn}"], bits, *plan.qparams[pn]), bits, *plan.qparams[pn])
    vt = QTensor(rq.quantize(env[vn].a, bits, *plan.qparams[vn]), bits, *plan.qparams[vn])
    pv = pt.matmul(vt)
    np.testing.assert_array_equal(pv.data, gv[f"{pre}/val_q/{a0}/MatMul_1_output_0"])   # P.V accumulator: exact
    # float glue fed the reference's values: softmax and GELU within 1e-5 of the reference's float32 results
    np.testing.assert_allclose(sc.dequantize().div(FTensor(np.array(np.sqrt(8.0), np.float32))).softmax(axis=-1).data,
                               gv[f"{pre}/val_f/{pn}"], rtol=1e-5, atol=1e-8)
    gel = "/vit/encoder/layer.0/intermediate/intermediate_act_fn/Mul_1_output_0"
    hin = FTensor(env["/vit/encoder/layer.0/intermediate/dense/Add_output_0"].a)
    np.testing.assert_allclose(hin.gelu_erf(1.4142135381698608, 1.0, 0.5).data, gv[f"{pre}/val_f/{gel}"], rtol=1e-5, atol=1e-6)
    out_scale = float(plan.qparams["logits"][0])
    want = gv[f"{pre}/out0"]
    assert np.abs(out - want).max() <= 4 * out_scale and np.abs(out - want).mean() <= 0.5 * out_scale
    # retain=False (freed intermediates, fused kernels) is bit-identical to the retained run
    q.fuse_softmax_epilogue = q.fuse_gelu_epilogue = q.fuse_layernorm_glue = False
    out2 = q([x], retain=False)[0]
    np.testing.assert_array_equal(out2, out)
    assert by[a0 + "/query/MatMul_output_0"].data is None
    # CUDA-graph replay of the fused forward: identical bits, also for a second, different input
    out4 = q([x], graph=True)[0]
    np.testing.assert_array_equal(out4, out)
    x_b = (x[::-1] * np.float32(0.5)).copy()
    np.testing.assert_array_equal(q([x_b], graph=True)[0], q([x_b], retain=False)[0])
    np.testing.assert_array_equal(q([x], graph=True)[0], out)
    # profile contract: dict of op type -> seconds with the two extra buckets (model.py:497-499)
    out3, prof = q([x], profile=True)
    np.testing.assert_array_equal(out3[0], out)
    assert {"TinyqQuant", "TinyqDequant", "MatMul", "Gemm", "Softmax", "LayerNormalization"} <= set(prof)
    assert all(isinstance(v, float) and v >= 0 for v in prof.values())


@pytest.mark.parametrize("bits", [8, 4])
def test_fused_executor_is_bit_identical_with_epilogue_quantization(bits):
    """Head dim 16 enables every fusion (epilogue quantize-into-operand for Q/K/V and the context,
    residual / bias epilogues, LayerNorm / Softmax / GELU -> quantize): retain=False, graph replay and the
    node-by-node retained run must agree bit for bit, and all agree with the oracle within the ViT bound."""
    cfg = dict(batch=3, image_size=32, patch_size=16, hidden=64, heads=4, intermediate=128, layers=2, classes=10)
    proto = zoo.vit_graph(seed=3, **cfg)
    x = np.random.default_rng(5).normal(size=(3, 3, 32, 32)).astype(np.float32)
    model = Model.from_onnx(proto)
    q = model.quantize([x], bit_width=bits)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], bits)
    inject_oracle_params(q, model, plan)
    ref = q([x])[0]
    q.fuse_softmax_epilogue = q.fuse_gelu_epilogue = q.fuse_layernorm_glue = False
    fused = q([x], retain=False)[0]
    np.testing.assert_array_equal(fused, ref)
    np.testing.assert_array_equal(q([x], graph=True)[0], ref)
    assert q._plan["to_operand"] and q._plan["merge_heads"] and q._plan["residual"]
    want = rg.run_quant(plan, [x])[0]
    step = float(plan.qparams["logits"][0])
    assert np.abs(ref - want).max() <= 4 * step
    # softmax in the score-GEMM epilogue and GELU in the first MLP GEMM's epilogue: float glue under the 1e-5
    # contract (different summation order / fused multiply-adds) -> same contract vs the reference, not
    # necessarily the same bits as the node-by-node run
    q.fuse_softmax_epilogue = q.fuse_gelu_epilogue = q.fuse_layernorm_glue = True
    fused2 = q([x], retain=False)[0]
    assert np.abs(fused2 - want).max() <= 4 * step and np.abs(fused2 - ref).max() <= 2 * step
    np.testing.assert_array_equal(q([x], graph=True)[0], fused2)


def test_fused_attention_falls_back_when_parameters_leave_its_window():
    """A softmax range calibrated far from zero (near-uniform attention: min and max both ~1/S) gives a P zero-point
    far below the code range; nq_attention_s8 refuses it (host-checked window) and the executor must take the two-GEMM
    route instead of failing -- same bits as running with fuse_attention=False."""
    cfg = dict(batch=3, image_size=32, patch_size=16, hidden=64, heads=4, intermediate=128, layers=2, classes=10)
    proto = zoo.vit_graph(seed=3, **cfg)
    x = np.random.default_rng(5).normal(size=(3, 3, 32, 32)).astype(np.float32)
    q = Model.from_onnx(proto).quantize([x], bit_width=8)
    names = [n.outputs[0].name for n in q.nodes if n.op == "Softmax"]
    assert len(names) == 2
    for n in names:
        q.quant_params[n] = QuantizationParams(np.float32(7.8e-6), np.int64(-638))      # range ~[0.004, 0.006]
    fused = q([x], retain=False)[0]
    assert q._plan["attention"]
    q.fuse_attention = False
    two_gemm = q([x], retain=False)[0]
    np.testing.assert_array_equal(fused, two_gemm)
    assert np.isfinite(fused).all()


@pytest.mark.parametrize("bits", [4, 2])
def test_packed_weight_storage_config4(bits):
    """BASELINE config 4 (ViT int4 / 2-bit with sub-byte packing): weights kept as bit_width-bit bitstreams and
    unpacked into transient int8 operands per GEMM give bit-identical outputs, `.data` of a packed weight still
    returns the reference's int64 codes, and the resident bytes shrink by 8 / bit_width (up to row padding)."""
    cfg = dict(batch=2, image_size=32, patch_size=16, hidden=64, heads=4, intermediate=128, layers=2, classes=10)
    proto = zoo.vit_graph(seed=7, **cfg)
    x = np.random.default_rng(8).normal(size=(2, 3, 32, 32)).astype(np.float32)
    model = Model.from_onnx(proto)
    q = model.quantize([x], bit_width=bits)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], bits)
    inject_oracle_params(q, model, plan)
    ref = q([x])[0]
    ref_fused = q([x], retain=False)[0]
    wname = next(n.inputs[1].name for n in q.nodes if n.op == "MatMul" and isinstance(n.inputs[1].data, QTensor)
                 and len(n.inputs[1].data.shape) == 2)
    wq = {v.name: v for v in q.values}[wname].data
    codes_before = wq.data.copy()
    info = q.pack_weights()
    assert info["resident_bytes"] <= info["int8_bytes"] * bits / 8 * 1.35 + 64, info
    np.testing.assert_array_equal(wq.data, codes_before)
    assert wq.data.dtype == np.int64 and np.abs(codes_before).max() <= 2 ** (bits - 1)
    np.testing.assert_array_equal(q([x])[0], ref)
    np.testing.assert_array_equal(q([x], retain=False)[0], ref_fused)
    np.testing.assert_array_equal(q([x], graph=True)[0], ref_fused)
    want = rg.run_quant(plan, [x])[0]
    step = float(plan.qparams["logits"][0])
    assert np.abs(ref - want).max() <= 4 * step


def test_vit_b16_full_size_batch_invariance_and_replay():
    """BASELINE config 2 at its full size (ViT-B/16, batch 256, synthetic 224x224), through size-independent
    properties: quantization parameters are per-tensor and static, every kernel works row by row, and row sums are
    exact integer sums, so (a) the logits of an image do not depend on what else is in the batch -- with the batch
    made of 32 copies of the same 8 images, every copy equals a batch-8 forward with the same parameters
    bit for bit -- and (b) for 256 distinct images eager fused execution, CUDA-graph replay and the pipelined submit
    path agree bit for bit."""
    vit = dict(image_size=224, patch_size=16, hidden=768, heads=12, intermediate=3072, layers=12, classes=1000)
    rng = np.random.default_rng(1)
    x8 = torch.from_numpy(rng.normal(size=(8, 3, 224, 224)).astype(np.float32)).cuda()
    x_rep = x8.repeat(32, 1, 1, 1).contiguous()
    big = Model.from_onnx(zoo.vit_graph(batch=256, seed=0, **vit))
    qbig = big.quantize([x_rep], bit_width=8)
    big.release()
    small = Model.from_onnx(zoo.vit_graph(batch=8, seed=0, **vit))
    qsmall = small.quantize([x8], bit_width=8)
    small.release()
    # The float calibration pass runs fp32 cuBLAS GEMMs whose summation order depends on the batch size, so min / max
    # can differ by an ulp between the two graphs: give the batch-8 model the batch-256 model's parameters and
    # quantized constants ("identical inputs and scales").
    checked = 0
    for name, qp in qbig.quant_params.items():
        other = qsmall.quant_params[name]
        if any(k in name for k in ("MatMul", "LayerNorm", "Softmax", "bias", "weight")):
            np.testing.assert_allclose(np.float64(qp.scale), np.float64(other.scale), rtol=1e-5, err_msg=name)
            checked += 1
        qsmall.quant_params[name] = qp
    assert checked > 200
    big_consts = {v.name: v for v in qbig.values if isinstance(v, Constant)}
    for v in qsmall.values:
        if isinstance(v, Constant) and isinstance(v.data, QTensor):
            v.data = big_consts[v.name].data
    qsmall._const_deq.clear()
    out8 = qsmall([x8], retain=False, device_outputs=True)[0].clone()
    out_rep = qbig([x_rep], retain=False, device_outputs=True)[0].clone()
    assert tuple(out_rep.shape) == (256, 1000) and bool(torch.isfinite(out_rep).all())
    assert torch.equal(out_rep.view(32, 8, 1000), out8.unsqueeze(0).expand(32, 8, 1000))
    x = torch.from_numpy(rng.normal(size=(256, 3, 224, 224)).astype(np.float32)).cuda()
    out256 = qbig([x], retain=False, device_outputs=True)[0].clone()
    replay = qbig([x], retain=False, device_outputs=True, graph=True)[0]
    assert torch.equal(replay, out256)
    host = qbig.submit([x.cpu().pin_memory()]).result()[0]
    np.testing.assert_array_equal(host, out256.cpu().numpy())
    assert len(set(out256.argmax(-1).cpu().tolist())) > 4      # not degenerate


@pytest.mark.parametrize("bits", [8, 4])
def test_qmodel_save_load_roundtrip(tmp_path, bits):
    """a quantized model written with packed codes reloads without float weights or calibration
    and reproduces every output bit for bit (node-by-node, fused and graph replay); MLP (Gemm/int64 biases) and ViT."""
    proto = ol.load(os.path.join(G, "mlp.onnx"))
    xm = np.random.default_rng(2).normal(size=(16, 2)).astype(np.float32)
    qm = Model.from_onnx(proto).quantize([xm], bit_width=bits)
    qm.save(str(tmp_path / "mlp.npz"))
    qm2 = QModel.load(str(tmp_path / "mlp.npz"))
    np.testing.assert_array_equal(qm2([xm])[0], qm([xm])[0])
    cfg = dict(batch=2, image_size=32, patch_size=16, hidden=64, heads=4, intermediate=128, layers=2, classes=10)
    x = np.random.default_rng(3).normal(size=(2, 3, 32, 32)).astype(np.float32)
    q = Model.from_onnx(zoo.vit_graph(seed=5, **cfg)).quantize([x], bit_width=bits)
    path = str(tmp_path / "vit.npz")
    q.save(path)
    q2 = QModel.load(path)
    assert q2.bit_width == bits and set(q2.quant_params) == set(q.quant_params)
    np.testing.assert_array_equal(q2([x])[0], q([x])[0])
    np.testing.assert_array_equal(q2([x], retain=False)[0], q([x], retain=False)[0])
    np.testing.assert_array_equal(q2([x], graph=True)[0], q([x], graph=True)[0])
    # packed codes: the file is much smaller than int8 weights would be at 4 bit
    wbytes = sum(int(np.prod(v.data.shape)) for v in q.values if isinstance(v, Constant) and isinstance(v.data, QTensor)
                 and v.data.bit_width <= 8)
    assert os.path.getsize(path) < wbytes * bits / 8 * 1.6 + 200000


def test_tensor_surface_off_the_hot_path():
    """the parts of the FTensor / QTensor surface that QModel never reaches on its hot path
    behave like the reference's NumPy expressions (tensor.py:60-153, 212-221, 245-253)."""
    from numpy_quant_b200.tensor import ITensor, concat, where
    rng = np.random.default_rng(4)
    x = rng.normal(size=(2, 3, 4, 5)).astype(np.float32)
    f = FTensor(x)
    np.testing.assert_array_equal(f.T.data, x.T)
    np.testing.assert_array_equal(f.transpose(0, 2, 1, 3).data, x.transpose(0, 2, 1, 3))
    np.testing.assert_array_equal(f.reshape(ITensor(np.array([6, 20], np.int64))).data, x.reshape(6, 20))
    np.testing.assert_array_equal(f.take(ITensor(np.array([2, 0], np.int64)), 1).data, x.take([2, 0], 1))
    np.testing.assert_array_equal(f[:, 1:3, ::2].data, x[:, 1:3, ::2])
    np.testing.assert_array_equal(f.copy().data, x)
    np.testing.assert_array_equal((-f).data, -x)
    np.testing.assert_array_equal((f + f).data, x + x)
    np.testing.assert_array_equal((f * f).data, x * x)
    np.testing.assert_array_equal(f.div(FTensor(np.float32(3.0) + np.abs(x))).data, x / (np.float32(3.0) + np.abs(x)))
    np.testing.assert_array_equal(f.relu().data, (x > 0) * x)
    np.testing.assert_array_equal(f.max(axis=-1, keepdims=True).data, x.max(-1, keepdims=True))
    np.testing.assert_allclose(f.mean(axis=-1, keepdims=False).data, x.mean(-1), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(f.sum(axis=1, keepdims=True).data, x.sum(1, keepdims=True), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(f.sigmoid().data, 1 / (1 + np.exp(-x)), rtol=1e-6)
    np.testing.assert_allclose(f.softmax(axis=-1).data, np.exp(x - x.max(-1, keepdims=True)) / np.exp(x - x.max(-1, keepdims=True)).sum(-1, keepdims=True), rtol=1e-5)
    e = FTensor(x[:, :1]).expand(ITensor(np.array([2, 3, 4, 5], np.int64)))
    np.testing.assert_array_equal(e.data, np.broadcast_to(x[:, :1], (2, 3, 4, 5)))
    np.testing.assert_array_equal(concat([f, f], axis=2).data, np.concatenate([x, x], axis=2))
    cond = ITensor((rng.random(size=(2, 3, 4, 5)) > 0.5).astype(np.int64))
    np.testing.assert_array_equal(where(cond, f, -f).data, np.where(cond.data, x, -x))
    np.testing.assert_array_equal(where(cond, ITensor(np.arange(120).reshape(2, 3, 4, 5)), ITensor(np.zeros((2, 3, 4, 5), np.int64))).data,
                                  np.where(cond.data, np.arange(120).reshape(2, 3, 4, 5), 0))
    with pytest.raises(ValueError):
        FTensor(x.astype(np.float64))
    with pytest.raises(ValueError):
        QTensor(np.zeros(3, np.int32), 8, np.float32(1.0))
    # QTensor.relu: clamp the codes at the zero-point; QTensor.sigmoid: dequantize -> sigmoid -> quantize with own params
    q = quantize_tensor_min_max(f, 8, True)
    codes, zp, sc = q.data, int(q.zero_point), np.float32(q.scale)
    want = codes.copy()
    want[want < zp] = zp
    np.testing.assert_array_equal(q.relu().data, want)
    act = (1 / (1 + np.exp(-rq.dequantize(codes, sc, np.int64(zp))))).astype(np.float32)
    qs = q.sigmoid()
    assert np.abs(qs.data - rq.quantize(act, 8, sc, np.int64(zp))).max() <= 1      # exp differs by <= 2 ulp from NumPy's
    assert qs.bit_width == 8 and np.float32(qs.scale) == sc and int(qs.zero_point) == zp
    with pytest.raises(AssertionError):
        q.matmul(quantize_tensor_min_max(FTensor(x.transpose(0, 1, 3, 2).copy()), 4, True))


def test_pipelined_submit_matches_graph_replay():
    """QModel.submit keeps two forwards in flight (H2D / kernels / D2H on separate streams): results are the
    graph-replay results, in submission order, for alternating inputs."""
    cfg = dict(batch=3, image_size=32, patch_size=16, hidden=64, heads=4, intermediate=128, layers=1, classes=10)
    proto = zoo.vit_graph(seed=4, **cfg)
    rng = np.random.default_rng(6)
    xs = [rng.normal(size=(3, 3, 32, 32)).astype(np.float32) for _ in range(3)]
    q = Model.from_onnx(proto).quantize([xs[0]], bit_width=8)
    want = [q([x], graph=True)[0].copy() for x in xs]
    pinned = [torch.from_numpy(x).pin_memory() for x in xs]
    got, pending = [], None
    for i in range(7):
        nxt = q.submit([pinned[i % 3]])
        if pending is not None:
            got.append(pending.result()[0].copy())
        pending = nxt
    got.append(pending.result()[0].copy())
    for i, g in enumerate(got):
        np.testing.assert_array_equal(g, want[i % 3])


def test_quantized_matmul_surface_known_answers():
    """test_quantization.py:40-149 restated on the device tensors, plus KA-1 literals."""
    k = np.load(os.path.join(G, "ka1.npz"))
    w, x = FTensor(k["w"]), FTensor(k["x"])
    y = w.matmul(x)
    ys, yz = quant_parameters(*tensor_min_max(y), bit_width=8, asymmetric=True)
    assert np.float32(ys) == k["y_scale"] and int(yz) == int(k["y_zp"])
    for wa in (False, True):
        for xa in (False, True):
            tag = f"{int(wa)}{int(xa)}"
            qw = quantize_tensor_min_max(w, 8, wa)
            qx = quantize_tensor_min_max(x, 8, xa)
            np.testing.assert_array_equal(qw.data, k[f"qw_{tag}"])
            np.testing.assert_array_equal(qx.data, k[f"qx_{tag}"])
            r = qw.matmul(qx)
            assert r.bit_width == 32 and np.float32(r.scale) == k[f"accs_{tag}"]
            np.testing.assert_array_equal(r.dequantize().data, k[f"dq_{tag}"])
            np.testing.assert_array_equal(r.requantize(8, ys, yz).data, k[f"rq_{tag}"])
            np.testing.assert_array_equal(r.data, k[f"acc_{tag}"])
            if int(k[f"acczf_{tag}"]):
                np.testing.assert_array_equal(np.broadcast_to(r.zero_point, r.data.shape),
                                              np.broadcast_to(k[f"accz_{tag}"], r.data.shape))
            else:
                assert r.zero_point is None
            np.testing.assert_allclose(r.dequantize().data, k["w"] @ k["x"], rtol=0.5)
    # broadcast batch dims (2,1,4,3) x (1,2,3,4), all sym/asym combos, vs the oracle
    rng = np.random.default_rng(0)
    wd, xd = rng.random((2, 1, 4, 3)).astype(np.float32), rng.random((1, 2, 3, 4)).astype(np.float32)
    for wa in (False, True):
        for xa in (False, True):
            qw = quantize_tensor_min_max(FTensor(wd), 8, wa)
            qx = quantize_tensor_min_max(FTensor(xd), 8, xa)
            acc, s, z = rq.q_matmul(qw.data, qw.scale, qw.zero_point, qx.data, qx.scale, qx.zero_point)
            r = qw.matmul(qx)
            np.testing.assert_array_equal(r.data, acc)
            np.testing.assert_array_equal(r.dequantize().data, rq.dequantize(acc, s, z))


def test_api_errors_match_reference():
    with pytest.raises(ValueError):
        FTensor(np.zeros(3, np.float64))                       # tensor.py:49-50
    with pytest.raises(ValueError):
        QTensor(np.zeros(3, np.int32), 8, np.float32(1.0))     # tensor.py:158-159
    m = Model.from_onnx(zoo.gemm_graph(3, 4, 2))
    with pytest.raises(ValueError):
        m([np.zeros((3, 4), np.float64)])                      # model.py:300-305
    with pytest.raises(AssertionError):
        a = QTensor(np.zeros((2, 2), np.int64), 8, np.float32(1.0))
        b = QTensor(np.zeros((2, 2), np.int64), 4, np.float32(1.0))
        a.matmul(b)                                            # tensor.py:206
    from numpy_quant_b200.model import onnx_operator_implementation
    with pytest.raises(ValueError, match="not supported"):
        onnx_operator_implementation("Foo", [], {})            # model.py:213


def test_no_device_memory_leak_over_repeated_quantize():
    """test/long_running/test_delete.py: 100x quantize must not accumulate memory."""
    proto = zoo.vit_graph(seed=0, **VIT_CFG)
    x = np.random.default_rng(1).normal(size=(2, 3, 32, 32)).astype(np.float32)
    model = Model.from_onnx(proto)
    model.quantize([x], 8)
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    for _ in range(30):
        model.quantize([x], 8)
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() <= base * 1.5 + (1 << 20)
