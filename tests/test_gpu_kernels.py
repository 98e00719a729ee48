"""Kernel-level parity: every C-ABI entry point of libnq_b200.so against the oracle /
the golden vectors of the unmodified reference.  Integer results bit-exact; float results
within 1e-5 relative (tolerance stated per test)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from numpy_quant_b200 import _lib, kernels as K  # noqa: E402
from oracle import ref_quant as rq  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = torch.device("cuda:0")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def record(name, **values):
    """Measured rates (flip fractions, worst deviations) of the float-glue stages: appended to
    gpurun_out/parity_rates.jsonl so that the asserted bounds can be kept at ~3x what is measured."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_rates.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **{k: (float(v) if not isinstance(v, str) else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass


def _zp(flag, val):
    return None if int(flag) == 0 else int(val)


@pytest.fixture(scope="module")
def kv():
    return np.load(os.path.join(G, "kernels.npz"))


def test_library_is_the_native_one():
    lib = _lib.load()
    assert lib.nq_version() >= 100
    info = (_lib.C.c_int * 4)()
    _lib.call("nq_device_info", info)
    assert info[1] == 10, f"expected a compute-capability 10.x device, got {info[1]}.{info[2]}"


# ----------------------------------------------------------------------------- K1
@pytest.mark.parametrize("bits", range(2, 9))
@pytest.mark.parametrize("asym", [False, True])
def test_quantize_golden(kv, bits, asym):
    tag = f"b{bits}_{'a' if asym else 's'}"
    scale, zp = kv[f"q_scale_{tag}"], _zp(kv[f"q_zpf_{tag}"], kv[f"q_zp_{tag}"])
    q = K.quantize(dev(kv[f"q_in_{tag}"]), bits, scale, zp)
    np.testing.assert_array_equal(host(q).astype(np.int64), kv[f"q_out_{tag}"])
    np.testing.assert_array_equal(host(K.dequantize(q, scale, zp)), kv[f"dq_out_{tag}"])


@pytest.mark.parametrize("bits,asym", [(8, True), (8, False), (4, True), (2, False)])
def test_quantize_large_vs_oracle(bits, asym):
    rng = np.random.default_rng(bits * 2 + asym)
    x = (rng.normal(size=(1 << 20) + 13) * 2).astype(np.float32)
    scale, zp = rq.quant_parameters(x.min(), x.max(), bits, asym)
    # plant exact ties: (k + 0.5) * scale and neighbours
    k = np.arange(-300, 300, dtype=np.float32)
    x[: k.size] = (k + np.float32(0.5)) * np.float32(scale)
    ref = rq.quantize(x, bits, scale, zp)
    got = host(K.quantize(dev(x), bits, scale, zp)).astype(np.int64)
    np.testing.assert_array_equal(got, ref)
    # unaligned view -> scalar kernel path
    got2 = host(K.quantize(dev(x)[3:], bits, scale, zp)).astype(np.int64)
    np.testing.assert_array_equal(got2, ref[3:])


@pytest.mark.parametrize("zp", [None, 0, -7, -8, 5, 130, -131, (1 << 20) - 1, 1 << 21, -(1 << 22) - 1])
def test_quantize_exact_ties_and_zero_point_parity(zp):
    """x / scale lands exactly on k + 0.5 (power-of-two scale): the float32-exact route must
    break ties like the reference's float64 `rint(zp + t)` for even and odd, small and huge
    zero-points (|zp| >= 2^20 switches to the float64 kernel)."""
    scale = np.float32(0.25)
    k = np.arange(-70000, 70000, dtype=np.float32)
    centre = 0.0 if zp is None else -float(zp)
    t = np.concatenate([k + 0.5, k, k + 0.25, centre + (k[69000:71000] + 0.5), centre + k[69900:70100] * 0.5,
                        [1e30, -1e30, 3e9, -3e9, 1e-30, -1e-30, 0.0, -0.0, 0.49999997, -0.49999997, 0.50000006]])
    x = (t.astype(np.float32) * scale).astype(np.float32)
    for bits in (8, 5, 2):
        ref = rq.quantize(x, bits, scale, None if zp is None else np.int64(zp))
        got = host(K.quantize(dev(x), bits, scale, zp)).astype(np.int64)
        np.testing.assert_array_equal(got, ref, err_msg=f"bits={bits} zp={zp}")
    # non-power-of-two scale: ordinary data plus values straddling the rounding boundaries
    rng = np.random.default_rng(abs(hash(zp)) % 1000)
    s2 = np.float32(0.0371)
    base = ((np.arange(-300, 300) + 0.5) * s2).astype(np.float32)
    x2 = np.concatenate([np.nextafter(base, np.float32(np.inf)), base, np.nextafter(base, np.float32(-np.inf)),
                         (rng.normal(size=100000) * 3).astype(np.float32)]).astype(np.float32)
    if zp is not None:
        x2 = np.concatenate([x2, (x2 - np.float32(zp) * s2).astype(np.float32)])
    ref = rq.quantize(x2, 8, s2, None if zp is None else np.int64(zp))
    op = K.quantize_operand(dev(x2).view(1, -1), "A", 8, s2, zp, True)
    np.testing.assert_array_equal(host(op.data)[0, 0, : x2.size].astype(np.int64), ref)
    assert int(op.rowsum[0, 0]) == int(ref.sum())


def test_hoisted_reciprocal_division_equals_ieee_division():
    """The kernels divide by per-tensor constants through a hoisted correctly-rounded reciprocal and
    two fused residual corrections; it must give the bits of IEEE division (np.float32 '/') always."""
    cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
    for mode, n in ((0, 1 << 28), (1, 1 << 28)):
        for seed in (1, 2):
            _lib.call("nq_selftest_division", n, seed, mode, cnt.data_ptr(), None)
    assert int(cnt.item()) == 0, f"{int(cnt.item())} quotients differ from __fdiv_rn"
    # and end to end through a kernel that uses it: Div by a scalar vs NumPy
    rng = np.random.default_rng(4)
    x = (rng.normal(size=1 << 20) * rng.choice([1e-3, 1.0, 1e3], size=1 << 20)).astype(np.float32)
    for c in (8.0, 1.4142135381698608, 0.0371, 3.0, 1e-7, 123456.7):
        np.testing.assert_array_equal(host(K.binary("div", dev(x), dev(np.array(c, np.float32)))), x / np.float32(c))


def test_quantize_operand_layouts():
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 3, 37, 50)).astype(np.float32)
    scale, zp = rq.quant_parameters(x.min(), x.max(), 8, True)
    ref = rq.quantize(x, 8, scale, zp)
    xd = dev(x)
    a = K.quantize_operand(xd, "A", 8, scale, zp, True)                     # rows = 37, k = 50
    assert a.data.shape == (6, 37, 64) and a.ld == 64
    np.testing.assert_array_equal(host(a.data)[:, :, :50].astype(np.int64), ref.reshape(6, 37, 50))
    assert not host(a.data)[:, :, 50:].any()
    np.testing.assert_array_equal(host(a.rowsum).astype(np.int64), ref.reshape(6, 37, 50).sum(-1))
    b = K.quantize_operand(xd, "B", 8, scale, zp, True)                     # rows = 50 (N), k = 37
    np.testing.assert_array_equal(host(b.data)[:, :, :37].astype(np.int64), ref.reshape(6, 37, 50).transpose(0, 2, 1))
    np.testing.assert_array_equal(host(b.rowsum).astype(np.int64), ref.reshape(6, 37, 50).sum(-2))
    # strided view (the attention K^T pattern): [B, S, H, D] -> permute(0, 2, 3, 1) as operand B
    base = rng.normal(size=(2, 11, 3, 8)).astype(np.float32)
    view = dev(base).permute(0, 2, 3, 1)                                   # [2, 3, 8, 11] = [.., K=8, N=11]
    refv = rq.quantize(base.transpose(0, 2, 3, 1), 8, scale, zp)
    ob = K.quantize_operand(view, "B", 8, scale, zp, True)
    np.testing.assert_array_equal(host(ob.data)[:, :, :8].astype(np.int64), refv.reshape(6, 8, 11).transpose(0, 2, 1))
    np.testing.assert_array_equal(host(ob.rowsum).astype(np.int64), refv.reshape(6, 8, 11).sum(-2))
    # short rows (attention head dim): sub-warp rows, strided head view [B, S, H, D] -> [B, H, S, D]
    for D in (64, 32, 20):
        hx = rng.normal(size=(3, 197, 4, D)).astype(np.float32)
        hv = dev(hx).permute(0, 2, 1, 3)
        oh = K.quantize_operand(hv, "A", 8, scale, zp, True)
        refh = rq.quantize(hx.transpose(0, 2, 1, 3), 8, scale, zp).reshape(12, 197, D)
        np.testing.assert_array_equal(host(oh.data)[:, :, :D].astype(np.int64), refh)
        np.testing.assert_array_equal(host(oh.rowsum).astype(np.int64), refh.sum(-1))
    # symmetric 2-D weight [K, N] as operand B
    w = rng.normal(size=(70, 33)).astype(np.float32)
    ws, _ = rq.quant_parameters(w.min(), w.max(), 4, False)
    ow = K.quantize_operand(dev(w), "B", 4, ws, None, True)
    refw = rq.quantize(w, 4, ws, None)
    np.testing.assert_array_equal(host(ow.data)[0, :, :70].astype(np.int64), refw.T)
    np.testing.assert_array_equal(host(ow.rowsum)[0].astype(np.int64), refw.sum(0))


def test_quantize_wide_bias():
    rng = np.random.default_rng(7)
    b = rng.normal(size=1000).astype(np.float32)
    for bits, scale in ((32, np.float32(2.5e-4)), (16, np.float32(0.07)), (8, np.float32(3.4)), (32, np.float32(1e-12))):
        ref = rq.quantize(b, bits, scale, None)
        np.testing.assert_array_equal(host(K.quantize_i64(dev(b), bits, scale)), ref)
    # asymmetric wide codes through the public module function (numpy_quantization.py:24-34 at any bit width):
    # int64 zero-point + float32 quotient -> float64 sum, float64 clip bounds
    from numpy_quant_b200 import numpy_quantization as nqz
    b[:4] = [1e30, -1e30, 0.0, -0.0]
    for bits, scale, zp in ((16, np.float32(0.07), 1234), (12, np.float32(0.5), -2047), (32, np.float32(2.5e-4), -77),
                            (9, np.float32(0.011), 255)):
        ref = rq.quantize(b, bits, scale, np.int64(zp))
        got = nqz.quantize(b, bits, scale, np.int64(zp))
        assert got.dtype == np.int64
        np.testing.assert_array_equal(got, ref)
    with pytest.raises(ValueError):
        nqz.quantize(b, 40, np.float32(1.0), None)
    with pytest.raises(ValueError, match="int8 range"):
        nqz.q_matmul(np.full((2, 2), 300, np.int64), np.float32(1), None, np.ones((2, 2), np.int64), np.float32(1), None)
    with pytest.raises(ValueError):
        nqz.requantize(np.ones((2, 2), np.int64), np.float32(1), None, np.float32(1), None, 12)


# ----------------------------------------------------------------------------- K2 / K3
def test_dequantize_wide_and_acc(kv):
    acc, sc = kv["acc"], kv["acc_scale"]
    np.testing.assert_array_equal(host(K.dequantize(dev(acc), sc, None)), kv["acc_dq_none"])
    np.testing.assert_array_equal(host(K.dequantize(dev(acc.astype(np.int32)), sc, None)), kv["acc_dq_none"])
    np.testing.assert_array_equal(host(K.dequantize(dev(acc), sc, -77)), rq.dequantize(acc, sc, np.int64(-77)))
    # factored zero-point: rowsum*zp_b + colsum*zp_a - zp_a*zp_b*k  == zrow + zcol - 77
    zrow, zcol = kv["acc_zrow"], kv["acc_zcol"]
    azp = K.AccZeroPoint(zp_a=1, zp_b=1, k=77, rowsum_a=dev(zrow.astype(np.int32).reshape(1, -1)),
                         colsum_b=dev(zcol.astype(np.int32).reshape(1, -1)), colsum_shared=True)
    got = K.dequantize_acc(dev(acc.astype(np.int32))[None], sc, azp)
    np.testing.assert_array_equal(host(got)[0], kv["acc_dq_full"])
    for bits in range(2, 9):
        for asym in (False, True):
            tag = f"b{bits}_{'a' if asym else 's'}"
            got = K.requantize_acc(dev(acc.astype(np.int32))[None], sc, azp, None, bits, kv[f"rq_scale_{tag}"],
                                   _zp(kv[f"rq_zpf_{tag}"], kv[f"rq_zp_{tag}"]))
            np.testing.assert_array_equal(host(got)[0].astype(np.int64), kv[f"rq_out_{tag}"], err_msg=tag)


# ----------------------------------------------------------------------------- K4 / K5
GEMM_SHAPES = [  # (batch, M, N, K, shared_B)
    (1, 3, 1, 3, False), (1, 5, 2, 5, False), (1, 128, 256, 128, False), (1, 129, 257, 129, False),
    (1, 300, 768, 768, False), (1, 256, 1000, 768, False), (1, 1000, 200, 3072, False),
    (24, 197, 197, 64, False), (24, 197, 64, 197, False), (3, 70, 130, 96, True), (2, 1, 1, 1, False),
]


@pytest.mark.parametrize("batch,M,N,Kd,shared", GEMM_SHAPES)
def test_qgemm_raw_bit_exact(batch, M, N, Kd, shared):
    rng = np.random.default_rng(M * 7 + N * 3 + Kd)
    a = rng.integers(-128, 128, size=(batch, M, Kd)).astype(np.int8)
    b = rng.integers(-128, 128, size=(1 if shared else batch, Kd, N)).astype(np.int8)
    ref = np.matmul(a.astype(np.int64), b.astype(np.int64))
    oa = K.operand_from_codes(dev(a), "A", False)
    ob = K.operand_from_codes(dev(b), "B", False)
    got = host(K.qgemm(oa, ob)).astype(np.int64)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(host(K.qgemm(oa, ob, simt=True)).astype(np.int64), ref)


def test_qgemm_extreme_values_and_persistence():
    # all -128 x -128 over K=4096 stresses the accumulator; > 148*2 tiles exercises the
    # TMEM double buffer and the smem ring wrap-around of the persistent loop
    M, N, Kd = 128 * 40, 256 * 9, 4096
    a = torch.full((1, M, Kd), -128, dtype=torch.int8, device=DEV)
    b = torch.full((1, Kd, N), -128, dtype=torch.int8, device=DEV)
    a[0, 5, 7] = 127
    got = K.qgemm(K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", False))
    assert int(got[0, 0, 0]) == 128 * 128 * Kd
    assert int(got[0, 5, 3]) == 128 * 128 * (Kd - 1) - 127 * 128
    assert bool((got[0, 6:] == 128 * 128 * Kd).all())


def test_qgemm_4096_cubed_vs_cuda_core_gemm():
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randint(-128, 128, (1, 4096, 4096), generator=g, device=DEV, dtype=torch.int8)
    b = torch.randint(-128, 128, (1, 4096, 4096), generator=g, device=DEV, dtype=torch.int8)
    oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", False)
    assert torch.equal(K.qgemm(oa, ob), K.qgemm(oa, ob, simt=True))


@pytest.mark.parametrize("za,zb", [(None, None), (None, -3), (5, None), (5, -3)])
@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("shape", [(3, 150, 200, 320), (2, 300, 256, 768)])     # ragged N: general epilogues; N % 16 == 0: row epilogues
def test_qgemm_fused_epilogues(za, zb, bits, shape):
    rng = np.random.default_rng(11)
    lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
    batch, M, N, Kd = shape
    a = rng.integers(lo, hi + 1, size=(batch, M, Kd)).astype(np.int64)
    b = rng.integers(lo, hi + 1, size=(batch, Kd, N)).astype(np.int64)
    sa, sb = np.float32(0.021), np.float32(0.0043)
    acc, s, z = rq.q_matmul(a, sa, None if za is None else np.int64(za), b, sb, None if zb is None else np.int64(zb))
    oa = K.operand_from_codes(dev(a.astype(np.int8)), "A", zb is not None)
    ob = K.operand_from_codes(dev(b.astype(np.int8)), "B", za is not None)
    azp = K.AccZeroPoint(za, zb, Kd, oa.rowsum, ob.rowsum, False)
    # DEQUANT (+ float bias)
    bias = rng.normal(size=N).astype(np.float32)
    want = rq.dequantize(acc, s, z)
    got = K.qgemm(oa, ob, _lib.EPI_DEQUANT, s, azp)
    np.testing.assert_array_equal(host(got), want)
    got = K.qgemm(oa, ob, _lib.EPI_DEQUANT, s, azp, bias_f32=dev(bias))
    np.testing.assert_array_equal(host(got), bias + want)
    # the unfused route must agree too
    raw = K.qgemm(oa, ob)
    np.testing.assert_array_equal(host(K.dequantize_acc(raw, s, azp)), want)
    # REQUANT with an int64 bias (Gemm), asymmetric and symmetric outputs
    bq = rng.integers(-20000, 20000, size=N).astype(np.int64)
    d = rq.dequantize(acc + bq, s, z)
    for asym in (True, False):
        so, zo = rq.quant_parameters(d.min(), d.max(), bits, asym)
        want_q = rq.requantize(acc + bq, s, z, so, zo, bits)
        got_q = K.qgemm(oa, ob, _lib.EPI_REQUANT, s, azp, bias_q=dev(bq), out_bits=bits, out_scale=so,
                        out_zp=None if zo is None else int(zo))
        np.testing.assert_array_equal(host(got_q).astype(np.int64), want_q)
        np.testing.assert_array_equal(
            host(K.requantize_acc(raw, s, azp, dev(bq), bits, so, None if zo is None else int(zo))).astype(np.int64), want_q)
        # without a bias (the standalone kernel's lean route) and with bias values beyond 2^29 (general 64-bit route)
        want_nb = rq.requantize(acc, s, z, so, zo, bits)
        np.testing.assert_array_equal(host(K.requantize_acc(raw, s, azp, None, bits, so, None if zo is None else int(zo))).astype(np.int64), want_nb)
        np.testing.assert_array_equal(host(K.qgemm(oa, ob, _lib.EPI_REQUANT, s, azp, out_bits=bits, out_scale=so,
                                                   out_zp=None if zo is None else int(zo))).astype(np.int64), want_nb)
        bw = bq.copy()
        bw[::7] = (1 << 33) + 12345
        bw[3::11] = -(1 << 31) - 77
        want_w = rq.requantize(acc + bw, s, z, so, zo, bits)
        np.testing.assert_array_equal(host(K.qgemm(oa, ob, _lib.EPI_REQUANT, s, azp, bias_q=dev(bw), out_bits=bits, out_scale=so,
                                                   out_zp=None if zo is None else int(zo))).astype(np.int64), want_w)


def test_qgemm_residual_epilogue_and_ragged_output_rows():
    """(bias + dequant) + residual in the epilogue == the three separate float ops; ragged N (197) is
    written through a row-padded buffer and returned as a strided view."""
    rng = np.random.default_rng(23)
    for (batch, M, N, Kd) in ((1, 300, 768, 256), (6, 197, 197, 64), (2, 70, 50, 40)):
        a = rng.integers(-128, 128, size=(batch, M, Kd)).astype(np.int64)
        b = rng.integers(-128, 128, size=(1, Kd, N)).astype(np.int64)
        acc, s, z = rq.q_matmul(a, np.float32(0.02), np.int64(-9), b, np.float32(0.004), None)
        bias = rng.normal(size=N).astype(np.float32)
        resid = rng.normal(size=(batch, M, N)).astype(np.float32)
        oa = K.operand_from_codes(dev(a.astype(np.int8)), "A", False)
        ob = K.operand_from_codes(dev(b.astype(np.int8)), "B", True)
        azp = K.AccZeroPoint(-9, None, Kd, None, ob.rowsum, True)
        got = K.qgemm(oa, ob, _lib.EPI_DEQUANT, s, azp, bias_f32=dev(bias), residual=dev(resid))
        assert tuple(got.shape) == (batch, M, N)
        want = (bias + rq.dequantize(acc, s, z)) + resid
        np.testing.assert_array_equal(host(got), want)
        if N % 4:
            assert got.stride(1) == K.round_up(N, 4) and not got.is_contiguous()
            # the padded view feeds the row kernels directly
            e = np.exp(want / np.float32(8) - (want / np.float32(8)).max(-1, keepdims=True))
            np.testing.assert_allclose(host(K.softmax_div_lastdim(got, 8.0)), e / e.sum(-1, keepdims=True), rtol=1e-5)
            ref = K.quantize_operand(K.softmax_div_lastdim(got, 8.0), "A", 8, 1 / 255, -128, True)
            fq = K.softmax_quantize(got, 8.0, 8, 1 / 255, -128, True)
            assert torch.equal(fq.data, ref.data) and torch.equal(fq.rowsum, ref.rowsum)
        # residual given as a strided (row-padded) view
        rp = torch.zeros((batch, M, N + 4), device=DEV)
        rp[:, :, :N] = dev(resid)
        got2 = K.qgemm(oa, ob, _lib.EPI_DEQUANT, s, azp, bias_f32=dev(bias), residual=rp[:, :, :N])
        np.testing.assert_array_equal(host(got2), want)


@pytest.mark.parametrize("zp_out", [None, -11, 6])
def test_qgemm_quantize_into_next_operand(zp_out):
    """NQ_EPI_QUANT: bias + dequant -> quantize -> scatter into the next MatMul's K-major operand must equal
    the unfused route (GEMM -> float32 -> reshape/transpose -> quantize_operand), codes and row sums."""
    rng = np.random.default_rng(31)
    B, S, H, D, Kd = 3, 197, 4, 16, 96
    a = rng.integers(-128, 128, size=(1, B * S, Kd)).astype(np.int8)
    w = rng.integers(-128, 128, size=(1, Kd, H * D)).astype(np.int8)
    bias = dev(rng.normal(size=H * D).astype(np.float32))
    oa, ob = K.operand_from_codes(dev(a), "A", False), K.operand_from_codes(dev(w), "B", True)
    azp = K.AccZeroPoint(7, None, Kd, None, ob.rowsum, True)
    sc, so = 3.1e-4, 0.021
    f = K.qgemm(oa, ob, _lib.EPI_DEQUANT, sc, azp, bias_f32=bias).view(B, S, H, D)          # float reference route
    # Q as A operand / K^T as B operand: [B, H, S, D] rows = S, k = D
    ref = K.quantize_operand(f.permute(0, 2, 1, 3), "A", 8, so, zp_out, True)
    got = K.qgemm_to_operand(oa, ob, sc, azp, bias, 8, so, zp_out, "split_rows", H, S, True)
    assert got.data.shape == ref.data.shape and torch.equal(got.data, ref.data) and torch.equal(got.rowsum, ref.rowsum)
    refb = K.quantize_operand(f.permute(0, 2, 3, 1), "B", 8, so, zp_out, True)               # K^T: same layout
    assert torch.equal(got.data, refb.data) and torch.equal(got.rowsum, refb.rowsum)
    # V as B operand: logical [B, H, S(K), D(N)] -> rows = D, k = S (padded to 208)
    ref = K.quantize_operand(f.permute(0, 2, 1, 3), "B", 8, so, zp_out, True)
    got = K.qgemm_to_operand(oa, ob, sc, azp, bias, 8, so, zp_out, "split_cols", H, S, True)
    assert got.ld == 208 and torch.equal(got.data[:, :, :S], ref.data[:, :, :S]) and torch.equal(got.rowsum, ref.rowsum)
    # attention context: batched (B*H) x [S, S'] . [S', D] -> [B, S, H*D] operand of the output projection
    p8 = rng.integers(-128, 128, size=(B * H, S, 40)).astype(np.int8)
    v8 = rng.integers(-128, 128, size=(B * H, 40, D)).astype(np.int8)
    op_, ov_ = K.operand_from_codes(dev(p8), "A", True), K.operand_from_codes(dev(v8), "B", True)
    azp2 = K.AccZeroPoint(-128, 9, 40, op_.rowsum, ov_.rowsum, False)
    f2 = K.qgemm(op_, ov_, _lib.EPI_DEQUANT, sc, azp2).view(B, H, S, D).permute(0, 2, 1, 3).reshape(B, S, H * D)
    ref = K.quantize_operand(f2, "A", 8, so, zp_out, True)
    got = K.qgemm_to_operand(op_, ov_, sc, azp2, None, 8, so, zp_out, "merge_heads", H, S, True)
    assert torch.equal(got.data.view(-1), ref.data.view(-1)) and torch.equal(got.rowsum.view(-1), ref.rowsum.view(-1))


@pytest.mark.parametrize("N,div", [(197, 8.0), (64, None), (224, 2.5), (5, 8.0)])
def test_qgemm_softmax_epilogue(N, div):
    """NQ_EPI_SOFTMAX_QUANT vs the separate route (GEMM -> dequant -> Div -> Softmax -> quantize): the float
    probabilities agree to 1e-5 (row sums are added in a different order), so codes may differ by one step
    only where p / scale sits on a rounding boundary; row sums must match the emitted codes exactly."""
    rng = np.random.default_rng(N)
    bt, M, Kd = 7, 197, 64
    a = rng.integers(-128, 128, size=(bt, M, Kd)).astype(np.int8)
    b = rng.integers(-128, 128, size=(bt, Kd, N)).astype(np.int8)
    oa, ob = K.operand_from_codes(dev(a), "A", True), K.operand_from_codes(dev(b), "B", True)
    azp = K.AccZeroPoint(3, -4, Kd, oa.rowsum, ob.rowsum, False)
    sc = 2.0e-4
    f = K.qgemm(oa, ob, _lib.EPI_DEQUANT, sc, azp)
    pf = K.softmax_div_lastdim(f, div) if div else K.softmax_lastdim(f)
    for zp in (-128, None):
        ref = K.quantize_operand(pf, "A", 8, 1 / 255, zp, True)
        got = K.qgemm_softmax_to_operand(oa, ob, sc, azp, div, 8, 1 / 255, zp, True)
        assert got.ld == K.round_up(N, 16) and got.data.shape == ref.data.shape
        gc, rc = host(got.data)[:, :, :N].astype(np.int64), host(ref.data)[:, :, :N].astype(np.int64)
        assert np.abs(gc - rc).max() <= 1 and np.mean(gc != rc) < 2e-3, (np.abs(gc - rc).max(), np.mean(gc != rc))
        np.testing.assert_array_equal(host(got.rowsum).astype(np.int64), gc.sum(-1))
        # dequantized probabilities still sum to ~1
        if zp is not None:
            assert np.abs((gc - zp).sum(-1) / 255.0 - 1).max() < 0.6
        # and against the ORACLE chain on the same codes (numpy_quantization.py:44-61 -> :37-41 -> Div -> tensor.py:139-146
        # -> :24-34): float-glue contract vs the float64 evaluation, at most one step off the float32 route
        p32, p64 = _softmax_chain_oracle(a, b, np.float32(sc), 3, np.float32(1.0), -4, div)
        _assert_codes_within_contract(gc, p64, np.float32(1 / 255), zp, 8, f"softmax epilogue N={N}")
        pr = rq.quantize(p32, 8, np.float32(1 / 255), None if zp is None else np.int64(zp))
        assert np.abs(gc - pr).max() <= 1 and np.mean(gc != pr) < 3e-3, (int(np.abs(gc - pr).max()), float(np.mean(gc != pr)))


@pytest.mark.parametrize("zp", [None, -7, 12])
def test_qgemm_gelu_epilogue(zp):
    """NQ_EPI_GELU_QUANT vs the separate route (GEMM -> dequant + bias -> GELU chain -> quantize).  The fused
    GELU is float glue under the 1e-5 contract: every code must be the round-half-even of a value within
    1e-5 (relative, plus 1e-6 of the tensor's range) of the reference float GELU / scale + zp; row sums
    must match the emitted codes exactly."""
    rng = np.random.default_rng(11)
    M, N, Kd = 300, 192, 96
    a = rng.integers(-128, 128, size=(1, M, Kd)).astype(np.int8)
    w = rng.integers(-128, 128, size=(1, Kd, N)).astype(np.int8)
    bias = dev(rng.normal(size=N).astype(np.float32))
    oa, ob = K.operand_from_codes(dev(a), "A", False), K.operand_from_codes(dev(w), "B", True)
    azp = K.AccZeroPoint(5, None, Kd, None, ob.rowsum, True)
    sc, so = 2.0e-5, 0.03
    c = (1.4142135381698608, 1.0, 0.5)
    f = K.qgemm(oa, ob, _lib.EPI_DEQUANT, sc, azp, bias_f32=bias)
    ref = K.gelu_quantize(f, *c, 8, so, zp, True)
    got = K.qgemm_to_operand(oa, ob, sc, azp, bias, 8, so, zp, "rows", 1, M, True, gelu=c)
    gc, rc = host(got.data).astype(np.int64), host(ref.data).astype(np.int64)
    assert gc.shape == rc.shape == (1, M, N)
    assert np.abs(gc - rc).max() <= 1 and np.mean(gc != rc) < 5e-3, (np.abs(gc - rc).max(), np.mean(gc != rc))
    np.testing.assert_array_equal(host(got.rowsum).astype(np.int64).reshape(1, M), gc.sum(-1))
    # contract check against the float64 evaluation of the reference chain on the same float32 inputs
    x = host(f).astype(np.float64)
    g = rq.erf_poly(x / c[0])
    g = (g + c[1]) * x * c[2]
    t = g / so + (0 if zp is None else zp)
    tol = 0.5 + (1e-5 * np.abs(g) + 1e-6 * np.abs(g).max()) / so + 1e-9
    t = np.clip(t, -128, 127)
    assert (np.abs(gc - t) <= tol).all(), float((np.abs(gc - t) - tol).max())


@pytest.mark.parametrize("shape", [(5, 197, 64, 64, 208), (3, 130, 48, 64, 144), (2, 7, 20, 32, 16), (1, 300, 100, 112, 304)])
def test_transpose_s8(shape):
    """nq_transpose_s8: out[b][c][r] = in[b][r][c], padding bytes r in [R, ld_out) zero, ragged R / C / unaligned ld_in."""
    bt, R, Cc, ld_in, ld_out = shape
    rng = np.random.default_rng(R)
    a = rng.integers(-128, 128, size=(bt, R, ld_in)).astype(np.int8)
    src = dev(a)
    out = torch.full((bt, Cc, ld_out), 55, dtype=torch.int8, device=src.device)
    _lib.call("nq_transpose_s8", src.data_ptr(), bt, R, Cc, ld_in, R * ld_in, out.data_ptr(), ld_out, Cc * ld_out, 0)
    torch.cuda.synchronize()
    got = host(out)
    np.testing.assert_array_equal(got[:, :, :R], a[:, :, :Cc].transpose(0, 2, 1))
    assert (got[:, :, R:] == 0).all()


def test_split_cols_operand_via_transpose_equals_the_column_epilogue(monkeypatch):
    """The V operand [head][D][S] produced by the row-layout epilogue + nq_transpose_s8 carries exactly the codes of
    the column-layout (byte-scattering) epilogue."""
    rng = np.random.default_rng(8)
    B, S, H, D, Kd = 3, 197, 4, 64, 96
    a = rng.integers(-128, 128, size=(1, B * S, Kd)).astype(np.int8)
    w = rng.integers(-128, 128, size=(1, Kd, H * D)).astype(np.int8)
    bias = dev(rng.normal(size=H * D).astype(np.float32))
    oa, ob = K.operand_from_codes(dev(a), "A", False), K.operand_from_codes(dev(w), "B", True)
    azp = K.AccZeroPoint(-9, None, Kd, None, ob.rowsum, True)
    new = K.qgemm_to_operand(oa, ob, 3.0e-5, azp, bias, 8, 0.04, 6, "split_cols", H, S, False)
    monkeypatch.setattr(K, "SPLIT_COLS_VIA_TRANSPOSE", False)
    old = K.qgemm_to_operand(oa, ob, 3.0e-5, azp, bias, 8, 0.04, 6, "split_cols", H, S, True)
    assert new.rows == old.rows == D and new.k == old.k == S and new.ld == old.ld
    np.testing.assert_array_equal(host(new.data)[:, :, :S], host(old.data)[:, :, :S])
    np.testing.assert_array_equal(host(K.rowsum(new)), host(old.rowsum))


def test_qgemm_gelu_epilogue_rejects_chains_that_are_not_a_gelu():
    """The fused epilogue drops the sign copy and the lower clamp, which needs Add constant 1 and positive Div / Mul
    constants (a GELU is bounded below); other constants are refused so that the caller keeps the node-by-node route."""
    rng = np.random.default_rng(3)
    M, N, Kd = 128, 64, 64
    a = rng.integers(-128, 128, size=(1, M, Kd)).astype(np.int8)
    w = rng.integers(-128, 128, size=(1, Kd, N)).astype(np.int8)
    bias = dev(rng.normal(size=N).astype(np.float32))
    oa, ob = K.operand_from_codes(dev(a), "A", False), K.operand_from_codes(dev(w), "B", True)
    azp = K.AccZeroPoint(5, None, Kd, None, ob.rowsum, True)
    for c in ((1.4142135381698608, 0.5, 0.5), (1.4142135381698608, 1.0, -0.5), (-1.4142135381698608, 1.0, 0.5)):
        with pytest.raises(_lib.NqError):
            K.qgemm_to_operand(oa, ob, 2.0e-5, azp, bias, 8, 0.03, None, "rows", 1, M, False, gelu=c)
    with pytest.raises(_lib.NqError):                                     # c1 * c3 / out_scale outside the rounding window
        K.qgemm_to_operand(oa, ob, 2.0e-5, azp, bias, 8, 1.0e-8, None, "rows", 1, M, False, gelu=(1.4142135381698608, 1.0, 0.5))


def _softmax_chain_oracle(q8, kt8, sq, zq, sk, zk, div):
    """Reference chain up to the float32 probabilities (numpy_quantization.py:44-61, :37-41, Div, tensor.py:139-146) plus
    the float64 evaluation of the same chain on the exact integer scores (the contract's centre)."""
    acc, s1, z1 = rq.q_matmul(q8.astype(np.int64), sq, None if zq is None else np.int64(zq),
                              kt8.astype(np.int64), sk, None if zk is None else np.int64(zk))
    y = rq.dequantize(acc, s1, z1)
    if div is not None:
        y = y / np.float32(div)
    m = y + (-(y.max(axis=-1, keepdims=True)))
    e = np.exp(m)
    p32 = e / e.sum(axis=-1, keepdims=True)
    y64 = (acc - (0 if z1 is None else z1)).astype(np.float64) * np.float64(np.float32(s1)) / (1.0 if div is None else float(div))
    e64 = np.exp(y64 - y64.max(axis=-1, keepdims=True))
    p64 = e64 / e64.sum(axis=-1, keepdims=True)
    return p32, p64


def _assert_codes_within_contract(codes, value64, scale, zp, bits, what):
    """Float-glue contract: every emitted code is the round-half-even of a value within 1e-5 relative
    (+1e-6 of the tensor's range) of the float64 evaluation / scale + zp."""
    lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
    t = np.clip(value64 / float(scale) + (0 if zp is None else zp), lo, hi)
    tol = 0.5 + (1e-5 * np.abs(value64) + 1e-6 * np.abs(value64).max()) / float(scale) + 1e-9
    bad = np.abs(codes - t) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} codes outside the 1e-5 contract, worst excess {float((np.abs(codes - t) - tol).max())}"


ATTN_CASES = [  # S, D, zq, zk, zv, p_zp, bits
    (197, 64, 3, -4, 9, -128, 8),          # ViT-B/16 geometry, asymmetric everything
    (197, 64, -128, 127, -128, -137, 8),   # extreme zero-points: every constant pass split in two, P below its range
    (197, 64, None, None, None, None, 8),  # symmetric everything (P centred at 0 -> constant lo_p)
    (50, 32, -7, 2, -1, -128, 8),          # single query tile, short rows
    (208, 16, 1, None, 5, -130, 8),        # full 208-key tile
    (64, 16, 0, 0, 0, 100, 8),             # P zero-point far inside the range: lo_p - zp_p = -228 as two constant passes
    (197, 64, 2, -1, 3, -8, 4),            # 4-bit codes
    (130, 48, -3, 5, -2, -2, 2),           # 2-bit codes, two tiles with a 2-row tail
]


@pytest.mark.parametrize("S,D,zq,zk,zv,p_zp,bits", ATTN_CASES)
def test_fused_attention_vs_oracle_chain(S, D, zq, zk, zv, p_zp, bits):
    """nq_attention_s8 against the reference chain q_matmul -> dequantize -> Div -> softmax -> quantize -> q_matmul ->
    dequantize -> Transpose/Reshape -> quantize (numpy_quantization.py:24-61, tensor.py:139-146), evaluated by the
    oracle on identical codes and scales:
      * every emitted P code satisfies the 1e-5 float-glue contract against the float64 evaluation and differs from the
        reference's float32 route by at most one step on a small, measured fraction;
      * GIVEN the emitted P codes, the P.V accumulator route is exact: the merged-heads output codes equal the oracle's
        q_matmul -> dequantize -> quantize of those codes bit for bit (and so do the row sums)."""
    rng = np.random.default_rng(S * 7 + D + bits)
    B, H = 2, 3
    lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
    q8 = rng.integers(lo, hi + 1, size=(B * H, S, D)).astype(np.int8)
    kt8 = rng.integers(lo, hi + 1, size=(B * H, D, S)).astype(np.int8)            # logical K^T [D, S]
    v8 = rng.integers(lo, hi + 1, size=(B * H, S, D)).astype(np.int8)
    oq, ok, ov = (K.operand_from_codes(dev(q8), "A", False), K.operand_from_codes(dev(kt8), "B", False),
                  K.operand_from_codes(dev(v8), "B", False))
    for o in (oq, ok, ov):
        o.batch_shape = (B, H)
    amp = float(hi - lo)
    sq = sk = np.float32(np.sqrt(8.0 * 12.0 / (D * (amp / 3.5) ** 2)))           # scores / 8 of a few units: peaky but not one-hot
    s_v, s_o = np.float32(0.02 * 255 / amp), np.float32(0.011 * 255 / amp)
    p32, p64 = _softmax_chain_oracle(q8, kt8, sq, zq, sk, zk, 8.0)
    s_p = np.float32(p32.max() / amp) if p_zp is not None else np.float32(2 * p32.max() / amp)
    o_zp = None if p_zp is None else -6 if bits == 8 else -1
    s1 = float(np.float32(sq) * np.float32(sk))
    s2 = float(np.float32(s_p) * np.float32(s_v))
    got, P = K.attention(oq, ok, ov, s1, zq, zk, 8.0, bits, s_p, p_zp, s2, zv, bits, s_o, o_zp, True, dump_p=True)
    Pc = host(P).astype(np.int64)
    # (1) the probabilities' codes
    _assert_codes_within_contract(Pc, p64, s_p, p_zp, bits, "P")
    P_ref = rq.quantize(p32, bits, s_p, None if p_zp is None else np.int64(p_zp))
    d = np.abs(Pc - P_ref)
    record("attention_P_vs_float32_route", S=S, D=D, bits=bits, p_zp=str(p_zp), flip_fraction=np.mean(d != 0), max_step=d.max())
    assert d.max() <= 1 and np.mean(d != 0) < 3e-3, (int(d.max()), float(np.mean(d != 0)))
    # (2) given the emitted codes: exact integer route to the output operand
    acc2, sc2, z2 = rq.q_matmul(Pc, s_p, None if p_zp is None else np.int64(p_zp), v8.astype(np.int64), s_v,
                                None if zv is None else np.int64(zv))
    ctx = rq.dequantize(acc2, sc2, z2).reshape(B, H, S, D).transpose(0, 2, 1, 3).reshape(B, S, H * D)
    want = rq.quantize(ctx, bits, s_o, None if o_zp is None else np.int64(o_zp))
    gc = host(got.data).astype(np.int64).reshape(B, S, H * D)
    np.testing.assert_array_equal(gc, want)
    np.testing.assert_array_equal(host(got.rowsum).astype(np.int64).reshape(B, S), want.sum(-1))
    # (3) end to end against the reference's own P codes: only the flipped P codes can move an output code
    acc3, _, z3 = rq.q_matmul(P_ref, s_p, None if p_zp is None else np.int64(p_zp), v8.astype(np.int64), s_v,
                              None if zv is None else np.int64(zv))
    ref_out = rq.quantize(rq.dequantize(acc3, sc2, z3).reshape(B, H, S, D).transpose(0, 2, 1, 3).reshape(B, S, H * D), bits, s_o,
                          None if o_zp is None else np.int64(o_zp))
    record("attention_out_vs_float32_route", S=S, D=D, bits=bits, p_zp=str(p_zp), flip_fraction=np.mean(gc != ref_out),
           max_step=np.abs(gc - ref_out).max())
    assert np.abs(gc - ref_out).max() <= 1 and np.mean(gc != ref_out) < 0.1, (int(np.abs(gc - ref_out).max()), float(np.mean(gc != ref_out)))


def test_fused_attention_rejects_out_of_window_parameters():
    """Parameters outside the host-checked windows are an error (the executor then takes the two-GEMM route), never a
    silently wrong result: P zero-point far below its range (narrow softmax range calibrated away from zero)."""
    rng = np.random.default_rng(3)
    q8 = rng.integers(-128, 128, size=(2, 60, 32)).astype(np.int8)
    oq, ok, ov = (K.operand_from_codes(dev(q8), "A", False), K.operand_from_codes(dev(q8.transpose(0, 2, 1).copy()), "B", False),
                  K.operand_from_codes(dev(q8), "B", False))
    for o in (oq, ok, ov):
        o.batch_shape = (1, 2)
    with pytest.raises(_lib.NqError, match="zp_p"):
        K.attention(oq, ok, ov, 2e-4, 1, 2, 8.0, 8, 7.8e-6, -638, 1e-6, 3, 8, 0.01, -5, False)
    with pytest.raises(_lib.NqError, match="zq"):
        K.attention(oq, ok, ov, 2e-4, 300, 2, 8.0, 8, 1 / 255, -128, 1e-4, 3, 8, 0.01, -5, False)


def test_qgemm_rejects_bad_arguments():
    a = torch.zeros((1, 4, 8), dtype=torch.int8, device=DEV)
    ep = _lib.Epilogue()
    with pytest.raises(_lib.NqError, match="multiples of 16"):
        _lib.call("nq_qgemm_s8", a.data_ptr(), a.data_ptr(), a.data_ptr(), 4, 4, 8, 1, 8, 8, 4, 0, 0, 0,
                  _lib.C.byref(ep), None)


# ----------------------------------------------------------------------------- K7-K9
def test_float_glue_vs_numpy(kv):
    rng = np.random.default_rng(3)
    np.testing.assert_allclose(host(K.unary("erf", dev(kv["erf_in"]))), kv["erf_out"], rtol=1e-5, atol=1e-7)
    x = (rng.normal(size=(7, 33, 768)) * 2).astype(np.float32)
    xd = dev(x)
    for name, fn in (("exp", np.exp), ("tanh", np.tanh), ("sqrt", lambda v: np.sqrt(np.abs(v))),
                     ("inv", lambda v: 1 / v), ("neg", lambda v: -v), ("relu", lambda v: (v > 0) * v),
                     ("sigmoid", lambda v: 1 / (1.0 + np.exp(-v)))):
        arg = dev(np.abs(x)) if name == "sqrt" else xd
        np.testing.assert_allclose(host(K.unary(name, arg)), fn(x), rtol=1e-5, atol=1e-30, err_msg=name)
    assert np.signbit(host(K.unary("relu", dev(np.array([-1.0, 2.0], np.float32))))[0])        # -0.0 quirk
    # GELU chain == five separate float32 ops
    c1, c2, c3 = np.float32(1.4142135381698608), np.float32(1.0), np.float32(0.5)
    want = (x * (rq.erf_poly(x / c1) + c2)) * c3
    np.testing.assert_allclose(host(K.gelu_erf(xd, c1, c2, c3)), want, rtol=1e-5, atol=1e-6)
    # broadcasting binaries are bit-exact (single IEEE op each)
    y = rng.normal(size=(7, 33, 768)).astype(np.float32)
    bias = rng.normal(size=768).astype(np.float32)
    np.testing.assert_array_equal(host(K.binary("add", xd, dev(y))), x + y)
    np.testing.assert_array_equal(host(K.binary("add", dev(bias), xd)), bias + x)
    np.testing.assert_array_equal(host(K.binary("mul", xd, dev(bias))), x * bias)
    np.testing.assert_array_equal(host(K.binary("div", xd, dev(np.array(8.0, np.float32)))), x / np.float32(8))
    col = rng.normal(size=(7, 33, 1)).astype(np.float32)
    np.testing.assert_array_equal(host(K.binary("mul", xd, dev(col))), x * col)
    np.testing.assert_array_equal(host(K.binary("add", xd.transpose(0, 1), dev(y).transpose(0, 1))),
                                  (x + y).transpose(1, 0, 2))
    # LayerNorm (model.py:134-152) and Softmax (tensor.py:139-146): 1e-5 relative
    g, b = (1 + rng.normal(size=768) * 0.02).astype(np.float32), (rng.normal(size=768) * 0.02).astype(np.float32)
    mean = x.mean(-1, keepdims=True)
    d = x - mean
    want = d * (1 / np.sqrt((d * d).mean(-1, keepdims=True) + np.float32(1e-12))) * g + b
    np.testing.assert_allclose(host(K.layernorm(xd, dev(g), dev(b), 1e-12)), want, rtol=1e-5, atol=2e-6)
    x2 = rng.normal(size=(5, 12, 197, 197)).astype(np.float32) * 3
    e = np.exp(x2 - x2.max(-1, keepdims=True))
    np.testing.assert_allclose(host(K.softmax_lastdim(dev(x2))), e / e.sum(-1, keepdims=True), rtol=1e-5, atol=1e-10)
    for cols in (5, 130, 300, 1500):
        x3 = rng.normal(size=(9, cols)).astype(np.float32)
        e = np.exp(x3 - x3.max(-1, keepdims=True))
        np.testing.assert_allclose(host(K.softmax_lastdim(dev(x3))), e / e.sum(-1, keepdims=True), rtol=1e-5)
        g3, b3 = rng.normal(size=cols).astype(np.float32), rng.normal(size=cols).astype(np.float32)
        d = x3 - x3.mean(-1, keepdims=True)
        want = d * (1 / np.sqrt((d * d).mean(-1, keepdims=True) + np.float32(1e-5))) * g3 + b3
        np.testing.assert_allclose(host(K.layernorm(dev(x3), dev(g3), dev(b3), 1e-5)), want, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(host(K.reduce_lastdim("mean", dev(x3), True)), x3.mean(-1, keepdims=True),
                                   rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(host(K.reduce_lastdim("max", dev(x3), False)), x3.max(-1))


@pytest.mark.parametrize("zp", [None, -3, 6])
def test_fused_producer_quantize_kernels_equal_the_two_step_route(zp):
    """LayerNorm->quantize, [Div->]Softmax->quantize, GELU->quantize emit exactly the codes (and row
    sums) that the float kernel followed by quantize_operand produces."""
    rng = np.random.default_rng(17)
    x = (rng.normal(size=(3, 50, 768)) * 2).astype(np.float32)
    g, b = (1 + rng.normal(size=768) * 0.1).astype(np.float32), (rng.normal(size=768) * 0.1).astype(np.float32)
    xd, gd, bd = dev(x), dev(g), dev(b)
    ref = K.quantize_operand(K.layernorm(xd, gd, bd, 1e-12), "A", 8, 0.03, zp, True)
    got = K.layernorm_quantize(xd, gd, bd, 1e-12, 8, 0.03, zp, True)
    assert torch.equal(got.data, ref.data) and torch.equal(got.rowsum, ref.rowsum)
    # float-glue variant (one FMA for gamma/beta, reciprocal of the scale): codes within one step on a small
    # fraction of rounding-boundary cases, row sums exact for the emitted codes
    glue = K.layernorm_quantize(xd, gd, bd, 1e-12, 8, 0.03, zp, True, float_glue=True)
    gc, rc = host(glue.data).astype(np.int64), host(ref.data).astype(np.int64)
    assert np.abs(gc - rc).max() <= 1 and np.mean(gc != rc) < 5e-3, (np.abs(gc - rc).max(), np.mean(gc != rc))
    np.testing.assert_array_equal(host(glue.rowsum).astype(np.int64).reshape(gc.shape[:-1]), gc.sum(-1))
    s = (rng.normal(size=(2, 4, 197, 197)) * 5).astype(np.float32)
    sd = dev(s)
    for div in (None, 8.0):
        f = K.softmax_lastdim(K.binary("div", sd, dev(np.array(div, np.float32)))) if div else K.softmax_lastdim(sd)
        if div:
            assert torch.equal(K.softmax_div_lastdim(sd, div), f)
        ref = K.quantize_operand(f, "A", 8, 1.0 / 255, zp, True)
        got = K.softmax_quantize(sd, div, 8, 1.0 / 255, zp, True)
        assert got.ld == 208 and torch.equal(got.data, ref.data) and torch.equal(got.rowsum, ref.rowsum)
    for shape in ((3, 50, 3072), (5, 7, 30)):
        h = (rng.normal(size=shape) * 2).astype(np.float32)
        hd = dev(h)
        c = (1.4142135381698608, 1.0, 0.5)
        ref = K.quantize_operand(K.gelu_erf(hd, *c), "A", 4, 0.2, zp, True)
        got = K.gelu_quantize(hd, *c, 4, 0.2, zp, True)
        assert torch.equal(got.data, ref.data) and torch.equal(got.rowsum, ref.rowsum)


def test_copy_and_im2col(kv):
    rng = np.random.default_rng(9)
    x = rng.normal(size=(2, 5, 3, 8)).astype(np.float32)
    np.testing.assert_array_equal(host(K.materialize(dev(x).permute(0, 2, 3, 1))), x.transpose(0, 2, 3, 1))
    np.testing.assert_array_equal(host(K.materialize(dev(x)[:, 1:4, :, ::2])), x[:, 1:4, :, ::2])
    cx, cw, cb = kv["conv_x"], kv["conv_w"], kv["conv_b"]
    cols, oh, ow = K.im2col(dev(cx), 3, 2, (0, 2, 2, 1), (2, 1))
    wmat = cw.transpose(2, 3, 1, 0).reshape(-1, 2)
    y = (host(cols) @ wmat).reshape(2, oh, ow, 2).transpose(0, 3, 1, 2) + cb[None, :, None, None]
    np.testing.assert_allclose(y, kv["conv_y"], rtol=1e-5, atol=1e-5)
    q = rng.integers(-128, 128, size=(2, 3, 9, 10)).astype(np.int8)
    qc, oh, ow = K.im2col(dev(q), 3, 2, (0, 2, 2, 1), (2, 1), pad_value=-7)
    qp = np.pad(q.transpose(0, 2, 3, 1), ((0, 0), (0, 2), (2, 1), (0, 0)), constant_values=-7)
    ref = np.stack([qp[:, i * 2:i * 2 + 3, j:j + 2, :].reshape(2, -1) for i in range(oh) for j in range(ow)], 1)
    np.testing.assert_array_equal(host(qc)[:, :18].reshape(2, oh * ow, 18), ref)


@pytest.mark.parametrize("geom", [
    # (n, C, H, W, O, kh, kw, pads, strides)
    (3, 64, 9, 10, 32, 3, 2, (0, 2, 2, 1), (2, 1)),          # test_conv2d geometry, one K slice per tap, 3 stages
    (2, 64, 7, 6, 48, 3, 1, (1, 0, 1, 0), (1, 1)),           # K = 192: last stage holds a single 64-byte slice
    (5, 128, 12, 11, 130, 2, 2, (1, 1, 0, 0), (3, 2)),       # two slices per tap, N tail, stride 3 x 2
    (2, 192, 5, 5, 16, 1, 1, (0, 0, 0, 0), (1, 1)),          # pointwise conv
    (37, 64, 8, 9, 256, 3, 3, (1, 1, 1, 1), (1, 1)),         # many M tiles with a ragged last one
])
def test_implicit_gemm_conv_equals_patch_matrix_route(geom):
    """nq_qconv2d_s8 (im2col-mode TMA, no patch matrix) gives the integer accumulators of the reference's
    extract_sliding_windows + matmul (numpy_helper.py:18-92) bit for bit, and the same float32 result as
    nq_im2col + nq_qgemm_s8 with the dequantize + bias epilogue."""
    n, Cc, H, W, O, kh, kw, pads, strides = geom
    rng = np.random.default_rng(n * 100 + Cc)
    zx = -7
    xq = rng.integers(-128, 128, size=(n, Cc, H, W)).astype(np.int8)
    wq = rng.integers(-127, 128, size=(O, Cc, kh, kw)).astype(np.int8)
    bias = rng.normal(size=O).astype(np.float32)
    ph0, pw0, ph1, pw1 = pads
    sh, sw = strides
    xp = np.pad(xq.transpose(0, 2, 3, 1), ((0, 0), (ph0, ph1), (pw0, pw1), (0, 0)), constant_values=zx).astype(np.int64)
    oh, ow = (xp.shape[1] - kh) // sh + 1, (xp.shape[2] - kw) // sw + 1
    patches = np.stack([xp[:, i * sh:i * sh + kh, j * sw:j * sw + kw, :].reshape(n, -1)
                        for i in range(oh) for j in range(ow)], 1).reshape(n * oh * ow, -1)
    wmat = wq.transpose(0, 2, 3, 1).reshape(O, -1).astype(np.int64)
    acc = patches @ wmat.T

    xd = dev(xq)
    opb = K.operand_from_codes(dev(wq.transpose(0, 2, 3, 1).reshape(O, -1).copy()), "A", True)
    nhwc = K.nhwc_pad(xd, pads, zx)
    hp, wp = xp.shape[1:3]
    np.testing.assert_array_equal(host(nhwc), xp)
    # the same image straight from float32 (quantize fused into the relayout): codes of nq_quantize_f32
    xf = (rng.normal(size=xq.shape) * 3).astype(np.float32)
    for bits, sc, zp in ((8, 0.05, zx), (8, 0.04, None), (4, 0.7, 2)):
        want_codes = host(K.quantize(dev(xf), bits, sc, zp)).transpose(0, 2, 3, 1)
        pc = 0 if zp is None else zp
        got = host(K.nhwc_pad(dev(xf), pads, pc, quant=(bits, sc, zp)))
        np.testing.assert_array_equal(got, np.pad(want_codes, ((0, 0), (ph0, ph1), (pw0, pw1), (0, 0)), constant_values=pc))
    raw, oh2, ow2 = K.qconv2d(nhwc, opb, kh, kw, strides)
    assert (oh2, ow2) == (oh, ow)
    np.testing.assert_array_equal(host(raw), acc)

    k = kh * kw * Cc
    azp = K.AccZeroPoint(zx, None, k, None, opb.rowsum, True)
    y, _, _ = K.qconv2d(nhwc, opb, kh, kw, strides, _lib.EPI_DEQUANT, 3e-4, azp, bias_f32=dev(bias))
    cols, _, _ = K.im2col(xd, kh, kw, pads, strides, zx)
    opa = K.Operand(cols.view(1, cols.shape[0], cols.shape[1]), (), cols.shape[0], k, cols.shape[1], None)
    y2 = K.qgemm(opa, opb, _lib.EPI_DEQUANT, 3e-4, azp, bias_f32=dev(bias))
    assert torch.equal(y, y2.view(y.shape))
    want = ((acc - zx * wmat.sum(1)[None, :]).astype(np.float64) * np.float64(np.float32(3e-4))).astype(np.float32) + bias
    np.testing.assert_array_equal(host(y), want)


# ----------------------------------------------------------------------------- K10 / K11
def test_minmax_and_pack_roundtrip():
    rng = np.random.default_rng(2)
    mm = K.minmax_slots(3, DEV)
    xs = [rng.normal(size=n).astype(np.float32) for n in (1, 1000, (1 << 20) + 5)]
    xs[1] = -np.abs(xs[1]) - 1                                  # all-negative tensor
    for i, x in enumerate(xs):
        K.minmax_into(dev(x), mm, i)
    got = host(mm)
    for i, x in enumerate(xs):
        assert got[i, 0] == x.min() and got[i, 1] == x.max()
    for bits in range(2, 9):
        lo, hi = -2 ** (bits - 1), 2 ** (bits - 1) - 1
        for n in (1, 7, 8, 9, 4096, 100003):
            q = rng.integers(lo, hi + 1, size=n).astype(np.int8)
            p = K.pack(dev(q), bits)
            assert p.numel() == (n * bits + 7) // 8
            np.testing.assert_array_equal(host(K.unpack(p, n, bits)), q)
            # layout contract: little-endian bitstream of two's-complement fields
            bitsream = np.unpackbits(host(p), bitorder="little")[: n * bits].reshape(n, bits)
            val = (bitsream * (1 << np.arange(bits))).sum(1)
            val = np.where(val >= 2 ** (bits - 1), val - 2 ** bits, val)
            np.testing.assert_array_equal(val, q)
