"""Pin the oracle (oracle/ref_quant.py, oracle/ref_graph.py) against vectors produced
by the UNMODIFIED reference (tests/golden/make_golden.py) -- bit for bit."""
import os
import warnings

import numpy as np
import pytest

from numpy_quant_b200 import onnx_lite as ol, zoo
from oracle import ref_graph as rg, ref_quant as rq

warnings.simplefilter("ignore")
G = os.path.join(os.path.dirname(__file__), "golden")


def _zp(flag, val):
    return None if int(flag) == 0 else np.int64(val)


@pytest.fixture(scope="module")
def kv():
    return np.load(os.path.join(G, "kernels.npz"))


@pytest.fixture(scope="module")
def gv():
    return np.load(os.path.join(G, "graphs.npz"))


@pytest.mark.parametrize("bits", range(2, 9))
@pytest.mark.parametrize("asym", [False, True])
def test_quantize_dequantize_bit_exact(kv, bits, asym):
    tag = f"b{bits}_{'a' if asym else 's'}"
    scale, zp = kv[f"q_scale_{tag}"], _zp(kv[f"q_zpf_{tag}"], kv[f"q_zp_{tag}"])
    q = rq.quantize(kv[f"q_in_{tag}"], bits, scale, zp)
    np.testing.assert_array_equal(q, kv[f"q_out_{tag}"])
    np.testing.assert_array_equal(rq.dequantize(q, scale, zp), kv[f"dq_out_{tag}"])
    # quant_parameters itself
    x = kv["x"]
    s2, z2 = rq.quant_parameters(x.min(), x.max(), bits, asym)
    assert s2.dtype == np.float32 and s2 == scale
    assert (z2 is None) == (zp is None) and (zp is None or int(z2) == int(zp))


def test_dequantize_array_zero_points(kv):
    acc, sc = kv["acc"], kv["acc_scale"]
    np.testing.assert_array_equal(rq.dequantize(acc, sc, None), kv["acc_dq_none"])
    np.testing.assert_array_equal(rq.dequantize(acc, sc, kv["acc_zrow"]), kv["acc_dq_row"])
    np.testing.assert_array_equal(rq.dequantize(acc, sc, kv["acc_zrow"] + kv["acc_zcol"] - 77), kv["acc_dq_full"])


@pytest.mark.parametrize("bits", range(2, 9))
@pytest.mark.parametrize("asym", [False, True])
def test_requantize_bit_exact(kv, bits, asym):
    tag = f"b{bits}_{'a' if asym else 's'}"
    zp_arr = kv["acc_zrow"] + kv["acc_zcol"] - 77
    got = rq.requantize(kv["acc"], kv["acc_scale"], zp_arr, kv[f"rq_scale_{tag}"],
                        _zp(kv[f"rq_zpf_{tag}"], kv[f"rq_zp_{tag}"]), bits)
    np.testing.assert_array_equal(got, kv[f"rq_out_{tag}"])


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_q_matmul_bit_exact(kv, bits):
    a, b = kv[f"mm_a_b{bits}"], kv[f"mm_b_b{bits}"]
    for za, zb in ((None, None), (None, np.int64(-3)), (np.int64(5), None), (np.int64(5), np.int64(-3))):
        tag = f"b{bits}_{'n' if za is None else 'z'}{'n' if zb is None else 'z'}"
        acc, s, z = rq.q_matmul(a, np.float32(0.02), za, b, np.float32(0.5), zb)
        np.testing.assert_array_equal(acc, kv[f"mm_acc_{tag}"])
        assert np.float32(s) == kv[f"mm_scale_{tag}"]
        if int(kv[f"mm_zpf_{tag}"]):
            np.testing.assert_array_equal(np.broadcast_to(z, acc.shape), kv[f"mm_zp_{tag}"])
        else:
            assert z is None


def test_erf_and_conv(kv):
    np.testing.assert_array_equal(rq.erf_poly(kv["erf_in"]), kv["erf_out"])
    y = rq.conv2d_nchw(kv["conv_x"], kv["conv_w"], kv["conv_b"], (0, 2, 2, 1), (2, 1))
    np.testing.assert_allclose(y, kv["conv_y"], rtol=0, atol=2e-6)      # BLAS summation order only


def test_known_answer_ka1():
    """SURVEY.md §8c KA-1 (test_quantization.py:44-52), literal numbers + generated file."""
    k = np.load(os.path.join(G, "ka1.npz"))
    w, x = k["w"], k["x"]
    expect = {"00": ([20613, -3408, 6186], [126, -128, -26]), "01": ([20582, -3285, 5568], [127, -128, -26]),
              "10": ([21216, -7415, 3881], [126, -128, -28]), "11": ([21497, -6951, 3461], [127, -128, -28])}
    for tag, (acc_e, rq_e) in expect.items():
        wa, xa = tag[0] == "1", tag[1] == "1"
        ws, wz = rq.quant_parameters(*rq.tensor_min_max(w), 8, wa)
        xs, xz = rq.quant_parameters(*rq.tensor_min_max(x), 8, xa)
        qw, qx = rq.quantize(w, 8, ws, wz), rq.quantize(x, 8, xs, xz)
        np.testing.assert_array_equal(qw, k[f"qw_{tag}"])
        np.testing.assert_array_equal(qx, k[f"qx_{tag}"])
        acc, s, z = rq.q_matmul(qw, ws, wz, qx, xs, xz)
        assert acc.ravel().tolist() == acc_e
        assert np.float32(s) == k[f"accs_{tag}"]
        got = rq.requantize(acc, s, z, k["y_scale"], np.int64(k["y_zp"]), 8)
        assert got.ravel().tolist() == rq_e
        np.testing.assert_array_equal(rq.dequantize(acc, s, z), k[f"dq_{tag}"])
    assert float(k["y_scale"]) == pytest.approx(0.06384314, rel=1e-6) and int(k["y_zp"]) == -92


def _check_graph(gv, prefix, plan, outs, keep=()):
    names = [str(n) for n in gv[f"{prefix}/qp_names"]]
    assert set(names) == set(plan.qparams), set(names) ^ set(plan.qparams)
    for n in names:
        s, z = plan.qparams[n]
        np.testing.assert_array_equal(np.asarray(s, dtype=np.float64), gv[f"{prefix}/qp_scale/{n}"], err_msg=n)
        flag = int(gv[f"{prefix}/qp_zpf/{n}"])
        assert (z is not None) == bool(flag), n
        if flag:
            assert int(z) == int(gv[f"{prefix}/qp_zp/{n}"]), n
    for key in gv.files:
        if key.startswith(f"{prefix}/qconst/"):
            n = key[len(prefix) + 8:]
            np.testing.assert_array_equal(plan.qconsts[n].a, gv[key], err_msg=n)
            assert plan.qconsts[n].bits == int(gv[f"{prefix}/qconst_bits/{n}"])
    for i, o in enumerate(outs):
        np.testing.assert_array_equal(o, gv[f"{prefix}/out{i}"])
    for n in keep:
        v = plan.env[n]
        if isinstance(v, rg.Q):
            np.testing.assert_array_equal(v.a, gv[f"{prefix}/val_q/{n}"], err_msg=n)
            assert v.bits == int(gv[f"{prefix}/val_qbits/{n}"])
            if v.zp is not None:
                np.testing.assert_array_equal(np.broadcast_to(v.zp, v.a.shape), gv[f"{prefix}/val_qzp/{n}"])
        else:
            np.testing.assert_array_equal(v.a, gv[f"{prefix}/val_f/{n}"], err_msg=n)


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 7, 8])
def test_mlp_reference_file(gv, bits):
    g = rg.import_graph(ol.load(os.path.join(G, "mlp.onnx")), ol)
    x = gv["mlp/x"]
    plan = rg.calibrate(g, [x], bits)
    outs = rg.run_quant(plan, [x])
    _check_graph(gv, f"mlp/b{bits}", plan, outs,
                 keep=["input", "/fc1/Gemm_output_0", "/relu/Relu_output_0", "/fc2/Gemm_output_0"])
    np.testing.assert_array_equal(rg.run_float(g, [x])[0], gv[f"mlp/b{bits}/fout0"])


def test_mlp_known_answer_ka2(gv):
    """SURVEY.md §8c KA-2 literal numbers (8-bit)."""
    g = rg.import_graph(ol.load(os.path.join(G, "mlp.onnx")), ol)
    x = gv["mlp/x"]
    plan = rg.calibrate(g, [x], 8)
    rg.run_quant(plan, [x])
    assert plan.env["input"].a.ravel().tolist() == [73, -28, -128, 107, 19, 127]
    assert plan.qconsts["fc1.weight"].a.ravel().tolist() == [127, -30, -86, 102, -128, -44, 68, 125, 2, -123]
    assert plan.qconsts["fc1.bias"].a.ravel().tolist() == [-7861, -9507, -10640, -10551, -7753]
    assert plan.qconsts["fc1.bias"].bits == 32
    assert plan.env["/fc1/Gemm_output_0"].a.ravel().tolist() == [60, -69, -63, -16, 27, -127, 127, 62, 5, -80, -13,
                                                                  60, -62, 84, -94]
    assert plan.env["/fc2/Gemm_output_0"].a.ravel().tolist() == [-57, 61, 127, -126, 26, -24]
    assert int(plan.qparams["/relu/Relu_output_0"][1]) == -128


def test_gemm_matmul_conv_graphs(gv):
    g = rg.import_graph(zoo.gemm_graph(3, 4, 2, seed=0), ol)
    plan = rg.calibrate(g, [gv["gemm/x"]], 8)
    _check_graph(gv, "gemm/b8", plan, rg.run_quant(plan, [gv["gemm/x"]]), keep=["output"])
    a, b = gv["matmul/a"], gv["matmul/b"]
    g = rg.import_graph(zoo.matmul_graph(a.shape, b.shape), ol)
    plan = rg.calibrate(g, [a, b], 8)
    _check_graph(gv, "matmul/b8", plan, rg.run_quant(plan, [a, b]), keep=["output"])
    g = rg.import_graph(zoo.conv_graph(2, 3, (9, 10), 2, (3, 2), (0, 2, 2, 1), (2, 1), seed=0), ol)
    plan = rg.calibrate(g, [gv["conv/x"]], 8)
    outs = rg.run_quant(plan, [gv["conv/x"]])
    np.testing.assert_allclose(outs[0], gv["conv/b8/out0"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("bits", [8, 4, 2])
def test_small_vit(gv, bits):
    cfg = dict(batch=2, image_size=32, patch_size=16, hidden=32, heads=4, intermediate=64, layers=2, classes=10)
    g = rg.import_graph(zoo.vit_graph(seed=0, **cfg), ol)
    x = gv["vit/x"]
    plan = rg.calibrate(g, [x], bits)
    outs = rg.run_quant(plan, [x])
    a0 = "/vit/encoder/layer.0/attention/attention"
    keep = ["/vit/embeddings/Add_output_0", "/vit/encoder/layer.0/layernorm_before/LayerNormalization_output_0",
            a0 + "/query/MatMul_output_0", a0 + "/MatMul_output_0", a0 + "/Softmax_output_0",
            a0 + "/MatMul_1_output_0", "/vit/encoder/layer.0/intermediate/intermediate_act_fn/Mul_1_output_0",
            "/vit/encoder/layer.1/output/Add_output_0", "logits"]
    _check_graph(gv, f"vit/b{bits}", plan, outs, keep=keep)


@pytest.mark.skipif(not os.path.isdir("/root/reference/numpy_quant"), reason="the reference checkout is only present in the build container")
def test_golden_recipe_regenerates_the_committed_vectors(tmp_path):
    """The committed recipe must run: tests/golden/make_golden.py imports the UNMODIFIED reference and rewrites every
    fixture into a scratch directory; the result must equal the committed files array for array (and mlp.onnx byte
    for byte)."""
    import filecmp
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(G, "make_golden.py"), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for f in ("kernels.npz", "ka1.npz", "graphs.npz"):
        new, old = np.load(tmp_path / f), np.load(os.path.join(G, f))
        assert sorted(new.files) == sorted(old.files), f
        for k in new.files:
            assert new[k].dtype == old[k].dtype and new[k].shape == old[k].shape, (f, k)
            assert np.array_equal(new[k], old[k], equal_nan=new[k].dtype.kind == "f"), (f, k)
    assert filecmp.cmp(tmp_path / "mlp.onnx", os.path.join(G, "mlp.onnx"), shallow=False)
