"""Oracle parity at the BENCHMARKED geometry (BASELINE config 2 / 4: ViT-B/16 -- hidden 768, 12 heads, head dim 64,
197 tokens, intermediate 3072) with the executor's default fusion flags, i.e. exactly what bench.py times.

Three layers of evidence, all against the oracle's `QModel.__call__` (oracle/ref_graph.py, a restatement of
/root/reference/numpy_quant/model.py:486-565 pinned bit for bit by tests/test_oracle_golden.py):

  (a) stem + 1 encoder layer + head, batch 2, int8 and int4: fused forward (retain=False, the bench path), CUDA-graph
      replay and the node-by-node retained run against `rg.run_quant`, with every quantization parameter and quantized
      constant taken from the oracle ("identical inputs and scales");
  (b) TEACHER-FORCED per-stage checks: each fused GPU stage is fed the ORACLE's own intermediate values (codes or
      float32) and must reproduce the oracle's next value -- bit-exact wherever the stage is integer / single-IEEE-op
      arithmetic (Q / K / V projections with bias and split heads, P.V given P, output projection + bias + residual,
      MLP-2 + bias + residual, classifier Gemm + requantize), inside the 1e-5 float-glue contract with a MEASURED flip
      rate where a float reduction or transcendental sits in front of the quantizer (LayerNorm, softmax, GELU);
  (c) one full 12-layer image against the oracle (about a minute of NumPy): top-1 equality as in
      /root/reference/test/long_running/test_vit.py:167 plus the logits bound.

Measured rates are appended to gpurun_out/parity_rates.jsonl; the asserted bounds are ~3x the measured values
(recorded in DESIGN.md 3).
"""
import json
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from numpy_quant_b200 import kernels as K, onnx_lite as ol, zoo  # noqa: E402
from numpy_quant_b200.model import Constant, Model, QuantizationParams  # noqa: E402
from numpy_quant_b200.tensor import FTensor, QTensor, quantize_tensor  # noqa: E402
from oracle import ref_graph as rg, ref_quant as rq  # noqa: E402

warnings.simplefilter("ignore")
VITB = dict(image_size=224, patch_size=16, hidden=768, heads=12, intermediate=3072, classes=1000)
B, S, H, D, HID = 2, 197, 12, 64, 768
L0 = "/vit/encoder/layer.0"
A0 = L0 + "/attention/attention"


def record(name, **values):
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_rates.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **{k: (float(v) if not isinstance(v, str) else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass


def inject_oracle_params(qmodel, fmodel, plan):
    """The oracle's exact quantization parameters and quantized constants ('using the same qparams')."""
    fconst = {v.name: v for v in fmodel.values if isinstance(v, Constant)}
    for name, (s, z) in plan.qparams.items():
        qmodel.quant_params[name] = QuantizationParams(s, z)
    for v in qmodel.values:
        if isinstance(v, Constant):
            oq = plan.qconsts[v.name]
            v.data = quantize_tensor(fconst[v.name].data, oq.bits, oq.scale, oq.zp)
            np.testing.assert_array_equal(v.data.data, oq.a, err_msg=v.name)
    qmodel._const_deq.clear()


class Case:
    """Stem + 1 layer + head of ViT-B/16 at batch 2: oracle plan / environment and the product model with the
    oracle's parameters."""

    def __init__(self, bits):
        self.bits = bits
        proto = zoo.vit_graph(batch=B, seed=0, layers=1, **VITB)
        self.x = np.random.default_rng(1).normal(size=(B, 3, 224, 224)).astype(np.float32)
        g = rg.import_graph(proto, ol)
        self.plan = rg.calibrate(g, [self.x], bits)
        self.want = rg.run_quant(self.plan, [self.x])[0]
        self.env = self.plan.env
        self.model = Model.from_onnx(proto)
        self.q = self.model.quantize([self.x], bit_width=bits)
        self.calibrated = {n: (np.float32(p.scale), p.zero_point) for n, p in self.q.quant_params.items()}
        inject_oracle_params(self.q, self.model, self.plan)
        self.model.release()
        self.by = {v.name: v for v in self.q.values}
        self.step = float(self.plan.qparams["logits"][0])
        self._noise = None

    def reference_self_sensitivity(self):
        """|logits(x * (1 + 1e-6 noise)) - logits(x)| of the ORACLE itself, in quantization steps: how far the reference
        algorithm's own output moves when its float inputs move by float32 rounding noise.  A quantizer turns any
        perturbation into code flips with probability ~ perturbation / step, and the flips cascade through the
        integer MatMuls, so free-running agreement of two correct implementations is bounded by THIS, not by 1e-5."""
        if self._noise is None:
            env = self.plan.env
            x2 = (self.x * (1 + 1e-6 * np.random.default_rng(2).normal(size=self.x.shape))).astype(np.float32)
            out2 = rg.run_quant(self.plan, [x2])[0]
            self.plan.env = env                                # keep the unperturbed environment for the other tests
            self._noise = np.abs(out2 - self.want) / self.step
        return self._noise

    def qp(self, name):
        s, z = self.plan.qparams[name]
        return np.float32(s), (None if z is None else int(z))

    def codes(self, name):
        """The oracle's quantization of a float value for its MatMul consumer (model.py:503-527)."""
        s, z = self.plan.qparams[name]
        return rq.quantize(self.env[name].a, self.bits, s, z)

    def qtensor(self, name, role="A"):
        s, z = self.plan.qparams[name]
        return QTensor(self.codes(name), self.bits, s, z)

    def const(self, name):
        return self.by[name].data

    def bias(self, name):
        return self.const(name).dequantize()


_CASES = {}


def case(bits):
    if bits not in _CASES:
        _CASES[bits] = Case(bits)
    return _CASES[bits]


def contract_excess(codes, value64, scale, zp, bits):
    lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
    t = np.clip(value64 / float(scale) + (0 if zp is None else zp), lo, hi)
    tol = 0.5 + (1e-5 * np.abs(value64) + 1e-6 * np.abs(value64).max()) / float(scale) + 1e-9
    return float((np.abs(codes - t) - tol).max())


# ------------------------------------------------------------------------------------------------ (a)
@pytest.mark.parametrize("bits", [8, 4])
def test_vitb_one_layer_fused_forward_vs_oracle(bits):
    """The bench path (retain=False, default fuse flags: attention kernel, GELU epilogue, LayerNorm glue) and its
    CUDA-graph replay against `rg.run_quant` on stem + 1 layer + head of ViT-B/16."""
    c = case(bits)
    q = c.q
    assert q.fuse_attention and q.fuse_softmax_epilogue and q.fuse_gelu_epilogue and q.fuse_layernorm_glue
    out = q([c.x], retain=False)[0]
    dev = np.abs(out - c.want) / c.step
    noise = c.reference_self_sensitivity()
    record("vitb_1layer_fused_logits", bits=bits, max_steps=dev.max(), mean_steps=dev.mean(), frac_exact=np.mean(dev == 0),
           oracle_noise_max_steps=noise.max(), oracle_noise_mean_steps=noise.mean(), oracle_noise_frac_exact=np.mean(noise == 0))
    # free-running bound: no further from the oracle than the oracle is from itself under 1e-6 input noise (x1.5 + 0.1 step)
    assert dev.max() <= noise.max() + 1.0 and dev.mean() <= 1.5 * noise.mean() + 0.1, (float(dev.max()), float(dev.mean()), float(noise.max()), float(noise.mean()))
    np.testing.assert_array_equal(out.argmax(-1), c.want.argmax(-1))
    np.testing.assert_array_equal(q([c.x], graph=True)[0], out)
    assert q._plan["attention"] and q._plan["to_operand"] and q._plan["gelu_in"] and q._plan["residual"]
    # the node-by-node retained run (every Value.data observable, as in the reference) against the oracle's environment
    out_r = q([c.x])[0]
    devr = np.abs(out_r - c.want) / c.step
    record("vitb_1layer_retained_logits", bits=bits, max_steps=devr.max(), mean_steps=devr.mean())
    assert devr.max() <= noise.max() + 1.0 and devr.mean() <= 1.5 * noise.mean() + 0.1
    worst_f, worst_q = 0.0, 0.0
    for name, ov in c.env.items():
        v = c.by.get(name)
        if v is None or v.data is None:
            continue
        if isinstance(ov, rg.F) and isinstance(v.data, FTensor):
            got = v.data.data
            scale = max(float(np.abs(ov.a).max()), 1e-30)
            worst_f = max(worst_f, float(np.abs(got - ov.a).max() / scale))
        elif isinstance(ov, rg.Q) and isinstance(v.data, QTensor) and ov.bits <= 8:
            worst_q = max(worst_q, float(np.abs(v.data.data - ov.a).max()))
    record("vitb_1layer_retained_values", bits=bits, worst_float_dev_rel_to_range=worst_f, worst_code_step=worst_q)
    # free-running values drift by the quantization steps of upstream single-code flips, not by float error: bounded by
    # a few steps here; the sharp statements are the teacher-forced tests below
    assert worst_f < 5.0 / (2 ** bits - 1) and worst_q <= noise.max() + 1, (worst_f, worst_q)


@pytest.mark.parametrize("bits", [8, 4])
def test_vitb_calibration_matches_oracle(bits):
    """Calibration statistics at ViT-B geometry: scales within float32 reduction noise of the oracle's, zero-points
    within one step (the float pass sums in a different order; min / max themselves are order-independent)."""
    c = case(bits)
    n = 0
    for name, (s, z) in c.plan.qparams.items():
        if not np.isfinite(s) or s == 0 or name not in c.calibrated:
            continue
        ps, pz = c.calibrated[name]
        np.testing.assert_allclose(np.float64(ps), np.float64(s), rtol=5e-5, err_msg=name)
        assert (pz is None) == (z is None), name
        if z is not None:
            assert abs(int(pz) - int(z)) <= 1, name
        n += 1
    assert n > 60


# ------------------------------------------------------------------------------------------------ (b)
@pytest.mark.parametrize("bits", [8, 4])
def test_teacher_forced_stem_and_head(bits):
    c = case(bits)
    # input quantization: bit-exact
    q_in = quantize_tensor(FTensor(c.x), bits, *c.plan.qparams["inputs"])
    np.testing.assert_array_equal(q_in.data, rq.quantize(c.x, bits, *c.plan.qparams["inputs"]))
    # patch embedding: the reference runs a float conv on the dequantized codes (model.py:95-100); here it is an integer
    # GEMM with exact accumulation -> float32 summation rounding of the reference only
    conv = "/vit/embeddings/patch_embeddings/projection/Conv_output_0"
    got = c.by[conv].data.data if c.by[conv].data is not None else None
    if got is None:
        c.q([c.x])
        got = c.by[conv].data.data
    ref = c.env[conv].a
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    record("vitb_stem_conv", bits=bits, max_err_rel_to_range=err)
    assert err < 3e-6                                                     # measured 6e-7
    # classifier: quantize(token) -> Gemm -> requantize, all integer / single-op arithmetic: bit-exact logits codes
    tok = c.env["/Gather_output_0"]
    s_t, z_t = c.plan.qparams["/Gather_output_0"]
    qt = quantize_tensor(FTensor(tok.a), bits, s_t, z_t)
    np.testing.assert_array_equal(qt.data, rq.quantize(tok.a, bits, s_t, z_t))
    acc = qt.matmul(c.const("classifier.weight").T) + c.const("classifier.bias")
    s_o, z_o = c.plan.qparams["logits"]
    logits_q = acc.requantize(bits, s_o, z_o)
    np.testing.assert_array_equal(logits_q.data, c.env["logits"].a)
    np.testing.assert_array_equal(logits_q.dequantize().data, c.want)


def _layernorm64(x, g, b, eps):
    x = x.astype(np.float64)
    d = x - x.mean(-1, keepdims=True)
    return d / np.sqrt((d * d).mean(-1, keepdims=True) + eps) * g.astype(np.float64) + b.astype(np.float64)


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("which", ["before", "after"])
def test_teacher_forced_layernorm_quantize(bits, which):
    """LayerNorm -> quantize (float glue): fed the oracle's float32 residual stream, every code within the 1e-5
    contract of the float64 evaluation; measured flip rate against the oracle's float32 route."""
    c = case(bits)
    xin = "/vit/embeddings/Add_output_0" if which == "before" else L0 + "/Add_output_0"
    out = f"{L0}/layernorm_{which}/LayerNormalization_output_0"
    g = rq.dequantize(c.plan.qconsts[f"vit.encoder.layer.0.layernorm_{which}.weight"].a, *c.plan.qparams[f"vit.encoder.layer.0.layernorm_{which}.weight"])
    b = rq.dequantize(c.plan.qconsts[f"vit.encoder.layer.0.layernorm_{which}.bias"].a, *c.plan.qparams[f"vit.encoder.layer.0.layernorm_{which}.bias"])
    x = c.env[xin].a
    s, z = c.qp(out)
    want = c.codes(out).reshape(B * S, HID)
    v64 = _layernorm64(x, g, b, float(np.float32(1e-12))).reshape(B * S, HID)
    xd, gd, bd = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, g, b))
    for glue in (True, False):
        op = K.layernorm_quantize(xd, gd, bd, float(np.float32(1e-12)), bits, float(s), z, True, float_glue=glue)
        got = op.data.cpu().numpy().astype(np.int64).reshape(B * S, -1)[:, :HID]
        d = np.abs(got - want)
        record("layernorm_quantize_flips", bits=bits, which=which, float_glue=str(glue), flip_fraction=np.mean(d != 0), max_step=d.max())
        assert d.max() <= 1 and np.mean(d != 0) < 3e-5, (int(d.max()), float(np.mean(d != 0)))   # measured <= 6.6e-6 (2 codes)
        assert contract_excess(got, v64, s, z, bits) <= 0
        np.testing.assert_array_equal(op.rowsum.cpu().numpy().reshape(-1), got.sum(-1))


@pytest.mark.parametrize("bits", [8, 4])
def test_teacher_forced_qkv_projections_bit_exact(bits):
    """Q / K / V projections: oracle LayerNorm codes in, MatMul + bias Add + Reshape + Transpose + quantize of the
    attention operands inside the GEMM epilogue -- integer accumulation and single IEEE operations only, so the codes
    (and the raw accumulator) must equal the oracle's bit for bit."""
    c = case(bits)
    ln = L0 + "/layernorm_before/LayerNormalization_output_0"
    a = c.qtensor(ln)
    # raw accumulator of the query MatMul
    acc = a.matmul(c.const("onnx::MatMul_q0"))
    np.testing.assert_array_equal(acc.data.reshape(B, S, HID), c.env[A0 + "/query/MatMul_output_0"].a)
    for tag, tr, kind, role, perm in (("query", "Transpose_1", "split_rows", "A", (0, 2, 1, 3)),
                                      ("key", "Transpose_2", "split_rows", "B", (0, 2, 3, 1)),
                                      ("value", "Transpose", "split_cols", "B", (0, 2, 1, 3))):
        out = f"{A0}/{tr}_output_0"
        s, z = c.plan.qparams[out]
        acc = a.matmul(c.const(f"onnx::MatMul_{tag[0]}0"))
        logical = tuple([B, S, H, D][p] for p in perm)
        qt = acc.quantize_into_operand(c.bias(f"vit.encoder.layer.0.attention.attention.{tag}.bias"), bits, s, z, kind, H, S, True,
                                       role, logical)
        assert qt is not None, tag
        np.testing.assert_array_equal(qt.data, c.codes(out), err_msg=tag)


@pytest.mark.parametrize("bits", [8, 4])
def test_teacher_forced_attention(bits):
    """The fused attention kernel fed the oracle's Q / K^T / V codes: P within the float-glue contract (measured flip
    rate vs the oracle's Softmax codes), and GIVEN the emitted P the context codes are the oracle arithmetic's, bit for
    bit; the end-to-end context codes differ from the oracle's only through flipped P codes (measured)."""
    c = case(bits)
    names = dict(q=A0 + "/Transpose_1_output_0", k=A0 + "/Transpose_2_output_0", v=A0 + "/Transpose_output_0",
                 p=A0 + "/Softmax_output_0", o=A0 + "/Reshape_3_output_0")
    qc, kc, vc = c.codes(names["q"]), c.codes(names["k"]), c.codes(names["v"])       # [B,H,S,D], [B,H,D,S], [B,H,S,D]
    dev8 = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int8))).cuda()
    oq, ok, ov = K.operand_from_codes(dev8(qc), "A", False), K.operand_from_codes(dev8(kc), "B", False), K.operand_from_codes(dev8(vc), "B", False)
    (sq, zq), (sk, zk), (sv, zv), (sp, zp), (so, zo) = (c.qp(names[k]) for k in "qkvpo")
    s1 = float(np.float32(sq) * np.float32(sk))
    s2 = float(np.float32(sp) * np.float32(sv))
    got, P = K.attention(oq, ok, ov, s1, zq, zk, 8.0, bits, sp, zp, s2, zv, bits, so, zo, True, dump_p=True)
    Pc = P.cpu().numpy().astype(np.int64).reshape(B, H, S, S)
    # oracle chain on the same codes
    acc1, sc1, z1 = rq.q_matmul(qc, sq, None if zq is None else np.int64(zq), kc, sk, None if zk is None else np.int64(zk))
    np.testing.assert_array_equal(acc1, c.env[A0 + "/MatMul_output_0"].a)          # the oracle's own score accumulator
    y64 = (acc1 - (0 if z1 is None else z1)).astype(np.float64) * np.float64(np.float32(sc1)) / 8.0
    e64 = np.exp(y64 - y64.max(-1, keepdims=True))
    p64 = e64 / e64.sum(-1, keepdims=True)
    assert contract_excess(Pc, p64, sp, zp, bits) <= 0
    P_ref = c.codes(names["p"])
    d = np.abs(Pc - P_ref)
    record("vitb_attention_P_flips", bits=bits, flip_fraction=np.mean(d != 0), max_step=d.max(), p_zp=str(zp))
    assert d.max() <= 1 and np.mean(d != 0) < 2e-5, (int(d.max()), float(np.mean(d != 0)))       # measured 5.4e-6 (5 codes)
    acc2, sc2, z2 = rq.q_matmul(Pc, sp, None if zp is None else np.int64(zp), vc, sv, None if zv is None else np.int64(zv))
    ctx = rq.dequantize(acc2, sc2, z2).transpose(0, 2, 1, 3).reshape(B, S, H * D)
    want = rq.quantize(ctx, bits, so, None if zo is None else np.int64(zo))
    gc = got.data.cpu().numpy().astype(np.int64).reshape(B, S, H * D)
    np.testing.assert_array_equal(gc, want)
    np.testing.assert_array_equal(got.rowsum.cpu().numpy().reshape(B, S), want.sum(-1))
    ref = c.codes(names["o"])
    dd = np.abs(gc - ref)
    record("vitb_attention_context_flips", bits=bits, flip_fraction=np.mean(dd != 0), max_step=dd.max())
    assert dd.max() <= 1 and np.mean(dd != 0) < 1e-4, (int(dd.max()), float(np.mean(dd != 0)))   # measured 3.3e-5 (10 codes)
    # the two-GEMM route of the executor's fallback gives the oracle's P.V arithmetic as well
    qP = QTensor(P_ref, bits, sp, None if zp is None else np.int64(zp))
    qV = QTensor(vc, bits, sv, None if zv is None else np.int64(zv))
    acc = qP.matmul(qV)
    np.testing.assert_array_equal(acc.data, c.env[A0 + "/MatMul_1_output_0"].a)
    q2 = qP.matmul(qV).quantize_into_operand(None, bits, so, None if zo is None else np.int64(zo), "merge_heads", H, S, True, "A", (B, S, H * D))
    np.testing.assert_array_equal(q2.data, ref)


@pytest.mark.parametrize("bits", [8, 4])
def test_teacher_forced_residual_projections_bit_exact(bits):
    """Output projection and MLP-2: oracle codes in, (bias + dequantize) + residual in the GEMM epilogue: three IEEE
    float32 operations on exact integers -> the float32 residual stream equals the oracle's bit for bit."""
    c = case(bits)
    for a_name, w, bias, resid, out in (
            (A0 + "/Reshape_3_output_0", "onnx::MatMul_o0", "vit.encoder.layer.0.attention.output.dense.bias",
             "/vit/embeddings/Add_output_0", L0 + "/Add_output_0"),
            (L0 + "/intermediate/intermediate_act_fn/Mul_1_output_0", "onnx::MatMul_f0", "vit.encoder.layer.0.output.dense.bias",
             L0 + "/Add_output_0", L0 + "/output/Add_output_0")):
        acc = c.qtensor(a_name).matmul(c.const(w))
        got = acc.dequantize(bias=c.bias(bias), residual=FTensor(c.env[resid].a)).data
        np.testing.assert_array_equal(got.reshape(B, S, HID), c.env[out].a, err_msg=out)


@pytest.mark.parametrize("bits", [8, 4])
def test_teacher_forced_mlp1_gelu(bits):
    """MLP-1 + bias + GELU chain + quantize in one epilogue (float glue): oracle LayerNorm codes in; codes within the 1e-5
    contract of the float64 evaluation of the reference chain, measured flip rate vs the oracle's float32 route."""
    c = case(bits)
    ln = L0 + "/layernorm_after/LayerNormalization_output_0"
    out = L0 + "/intermediate/intermediate_act_fn/Mul_1_output_0"
    s, z = c.plan.qparams[out]
    acc = c.qtensor(ln).matmul(c.const("onnx::MatMul_i0"))
    np.testing.assert_array_equal(acc.data.reshape(B, S, -1), c.env[L0 + "/intermediate/dense/MatMul_output_0"].a)
    acc = c.qtensor(ln).matmul(c.const("onnx::MatMul_i0"))
    qg = acc.gelu_into_operand(c.bias("vit.encoder.layer.0.intermediate.dense.bias"), (1.4142135381698608, 1.0, 0.5), bits, s, z, True)
    assert qg is not None
    got = qg.data.reshape(B * S, -1)
    want = c.codes(out).reshape(B * S, -1)
    d = np.abs(got - want)
    record("vitb_gelu_flips", bits=bits, flip_fraction=np.mean(d != 0), max_step=d.max())
    assert d.max() <= 1 and np.mean(d != 0) < 4e-5, (int(d.max()), float(np.mean(d != 0)))       # measured 1.07e-5 (13 codes)
    h = c.env[L0 + "/intermediate/dense/Add_output_0"].a.astype(np.float64).reshape(B * S, -1)      # exact input of the chain
    g64 = (rq.erf_poly((h / 1.4142135381698608).astype(np.float32)).astype(np.float64) + 1.0) * h * 0.5
    assert contract_excess(got, g64, s, None if z is None else int(z), bits) <= 1e-3


# ------------------------------------------------------------------------------------------------ (c)
def test_vitb_full_depth_one_image_vs_oracle():
    """All 12 layers, one image, int8, the bench path against the oracle (reference test/long_running/test_vit.py:167
    asserts top-1 equality of the quantized model; here it is asserted against the reference algorithm itself)."""
    proto = zoo.vit_graph(batch=1, seed=0, layers=12, **VITB)
    x = np.random.default_rng(7).normal(size=(1, 3, 224, 224)).astype(np.float32)
    plan = rg.calibrate(rg.import_graph(proto, ol), [x], 8)
    want = rg.run_quant(plan, [x])[0]
    model = Model.from_onnx(proto)
    q = model.quantize([x], bit_width=8)
    inject_oracle_params(q, model, plan)
    model.release()
    out = q([x], retain=False)[0]
    step = float(plan.qparams["logits"][0])
    dev = np.abs(out - want) / step
    # the oracle's own sensitivity to float32-rounding-sized input noise (see Case.reference_self_sensitivity)
    x2 = (x * (1 + 1e-6 * np.random.default_rng(2).normal(size=x.shape))).astype(np.float32)
    noise = np.abs(rg.run_quant(plan, [x2])[0] - want) / step
    record("vitb_12layer_fused_logits", bits=8, max_steps=dev.max(), mean_steps=dev.mean(), frac_exact=np.mean(dev == 0),
           oracle_noise_max_steps=noise.max(), oracle_noise_mean_steps=noise.mean())
    assert int(out.argmax()) == int(want.argmax())
    assert dev.max() <= noise.max() + 2.0 and dev.mean() <= 1.5 * noise.mean() + 0.25, (float(dev.max()), float(dev.mean()), float(noise.max()), float(noise.mean()))
    np.testing.assert_array_equal(q([x], graph=True)[0], out)
