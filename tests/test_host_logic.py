"""CPU-side tests of the host logic: ONNX-lite reader/writer, graph zoo, quant_parameters,
batch sharding and the world_size-2 (gloo) path of the one collective (calibration min/max)."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numpy_quant_b200 import _lib, distributed as nqd, onnx_lite as ol, zoo
from numpy_quant_b200.numpy_quantization import quant_parameters
from oracle import ref_quant as rq

warnings.simplefilter("ignore")
G = os.path.join(os.path.dirname(__file__), "golden")


def test_onnx_lite_reads_reference_mlp_and_roundtrips():
    m = ol.load(os.path.join(G, "mlp.onnx"))
    assert [n.op_type for n in m.graph.node] == ["Gemm", "Relu", "Gemm", "Sigmoid"]
    assert m.opset == 10 and m.graph.input[0].shape == ["batch_size", 2]
    w = ol.to_array(m.graph.initializer[0])
    assert w.shape == (5, 2) and w.dtype == np.float32 and w[0, 0] == np.float32(4.403712272644043)
    attrs = {a.name: ol.get_attribute_value(a) for a in m.graph.node[0].attribute}
    assert attrs == {"alpha": 1.0, "beta": 1.0, "transB": 1}
    m2 = ol.load_model_from_string(ol.serialize(m))
    for a, b in zip(m.graph.initializer, m2.graph.initializer):
        np.testing.assert_array_equal(ol.to_array(a), ol.to_array(b))
    assert [(n.name, n.op_type, n.input, n.output) for n in m.graph.node] == \
           [(n.name, n.op_type, n.input, n.output) for n in m2.graph.node]


def test_onnx_lite_attribute_kinds_roundtrip():
    n = ol.make_node("Foo", ["a"], ["b"], name="n", axis=-1, eps=1e-12, perm=[0, 2, 1], value=np.array([1, -1, -1], np.int64),
                     s="abc", scales=[0.5, 2.0])
    g = ol.GraphProto(node=[n], input=[ol.ValueInfoProto("a", ol.FLOAT, [1, "N"])], output=[ol.ValueInfoProto("b")])
    m = ol.load_model_from_string(ol.serialize(ol.ModelProto(graph=g)))
    got = {a.name: ol.get_attribute_value(a) for a in m.graph.node[0].attribute}
    assert got["axis"] == -1 and got["perm"] == [0, 2, 1] and got["s"] == b"abc"
    assert got["eps"] == pytest.approx(1e-12) and got["scales"] == pytest.approx([0.5, 2.0])
    np.testing.assert_array_equal(ol.to_array(got["value"]), [1, -1, -1])
    assert m.graph.input[0].shape == [1, "N"]
    with pytest.raises(ValueError, match="externally"):
        ol.to_array(ol.TensorProto(name="w", dims=[2], data_location=1, external_data={"location": "w.data"}))


def test_onnx_lite_external_data_roundtrip(tmp_path):
    """The reference's exporter writes ViT weights with save_as_external_data=True (models/vit.py:71-86): initializers
    whose payload sits in a side file load transparently, and save() can produce the same layout."""
    m = zoo.vit_graph(batch=1, layers=1, hidden=32, heads=4, intermediate=64, image_size=32, classes=3)
    path = tmp_path / "vit.onnx"
    ol.save(m, path, save_as_external_data=True, size_threshold=256)
    side = tmp_path / "vit.onnx.data"
    assert side.exists() and side.stat().st_size > 0
    bare = ol.load(path, load_external=False)
    ext = [t for t in bare.graph.initializer if t.data_location == 1]
    small = [t for t in bare.graph.initializer if t.data_location != 1]
    assert ext and small and all(not t.raw_data for t in ext)
    assert all(int(t.external_data["offset"]) % 64 == 0 and t.external_data["location"] == "vit.onnx.data" for t in ext)
    with pytest.raises(ValueError, match="externally"):
        ol.to_array(ext[0])
    full = ol.load(path)
    assert [t.name for t in full.graph.initializer] == [t.name for t in m.graph.initializer]
    for a, b in zip(m.graph.initializer, full.graph.initializer):
        assert b.data_location == 0
        np.testing.assert_array_equal(ol.to_array(a), ol.to_array(b))
    # a truncated side file is an error, not silent garbage
    side.write_bytes(side.read_bytes()[:100])
    with pytest.raises(ValueError, match="outside"):
        ol.load(path)


def test_onnx_lite_external_data_stays_inside_the_model_directory(tmp_path):
    """`location` comes from the model file: absolute paths, `..` escapes and symlinks out of the model's directory
    are refused instead of being read into tensors."""
    secret = tmp_path / "secret.bin"
    secret.write_bytes(np.arange(4, dtype=np.float32).tobytes())
    mdir = tmp_path / "model"
    mdir.mkdir()
    m = zoo.vit_graph(batch=1, layers=1, hidden=32, heads=4, intermediate=64, image_size=32, classes=3)
    path = mdir / "vit.onnx"
    ol.save(m, path, save_as_external_data=True, size_threshold=256)
    for loc in ("../secret.bin", str(secret), "sub/../../secret.bin"):
        bare = ol.load(path, load_external=False)
        t = next(t for t in bare.graph.initializer if t.data_location == 1)
        t.external_data = {"location": loc, "offset": "0", "length": "16"}
        with pytest.raises(ValueError, match="refused|escapes"):
            ol.load_external_data(bare, mdir)
    (mdir / "link.data").symlink_to(secret)
    bare = ol.load(path, load_external=False)
    t = next(t for t in bare.graph.initializer if t.data_location == 1)
    t.external_data = {"location": "link.data", "offset": "0", "length": "16"}
    with pytest.raises(ValueError, match="escapes"):
        ol.load_external_data(bare, mdir)


def test_vit_zoo_census_matches_committed_topology():
    """SURVEY.md §3.5 census of models/vit/vit_image_classifier_no_weights.onnx."""
    from collections import Counter
    m = zoo.vit_graph()
    c = Counter(n.op_type for n in m.graph.node)
    assert len(m.graph.node) == 516 and len(m.graph.initializer) == 200
    assert c == {"Add": 109, "Constant": 104, "MatMul": 96, "Reshape": 49, "Transpose": 49, "Mul": 25,
                 "LayerNormalization": 25, "Div": 24, "Softmax": 12, "Erf": 12, "Concat": 2, "Conv": 1, "Shape": 1,
                 "Slice": 1, "ConstantOfShape": 1, "Equal": 1, "Where": 1, "Expand": 1, "Gather": 1, "Gemm": 1}
    b = zoo.vit_graph(batch=7, layers=1, hidden=32, heads=4, intermediate=64, image_size=32, classes=3)
    shapes = [ol.to_array(a.t).tolist() for n in b.graph.node if n.op_type == "Constant" for a in n.attribute
              if a.t is not None and len(a.t.dims) == 1 and a.t.dims[0] in (3, 4)]
    assert all(s[0] == 7 for s in shapes if len(s) in (3, 4) and s != [3])
    for t in b.graph.initializer:
        assert ol.to_array(t).max() > 0                  # symmetric scale 2*max/(2^b-1) must be positive


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_quant_parameters_matches_oracle_bitwise(bits):
    rng = np.random.default_rng(bits)
    for _ in range(500):
        a, b = (np.float32(v) for v in rng.normal(size=2) * rng.choice([1e-4, 1.0, 1e4]))
        lo, hi = min(a, b), max(a, b)
        for asym in (False, True):
            s1, z1 = quant_parameters(lo, hi, bits, asym)
            s2, z2 = rq.quant_parameters(lo, hi, bits, asym)
            assert s1.dtype == np.float32 and s1.tobytes() == s2.tobytes()
            assert (z1 is None) == (z2 is None) and (z1 is None or int(z1) == int(z2))
    s, z = quant_parameters(np.float32(-1.0), np.float32(1.0), 2, True)        # KA: unclamped, scalar-zero idiom
    assert z == 0 and isinstance(z, np.int64)
    s, z = quant_parameters(np.float32(0.5), np.float32(1.0), 8, True)
    assert int(z) == -383                                                       # not clamped to [-128, 127]


def test_shard_bounds_cover_batch_exactly():
    for n in (1, 7, 256, 4096):
        for ws in (1, 2, 3, 8):
            spans = [nqd.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    x = np.arange(10).reshape(10, 1)
    assert np.concatenate([nqd.shard_batch([x], r, 3)[0] for r in range(3)]).tolist() == x.tolist()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, ws, port, full, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = nqd.shard_bounds(full.shape[1], rank, ws)
        shard = full[:, lo:hi]                                   # [n_values, my images, features]
        mm = torch.stack([shard.reshape(shard.shape[0], -1).min(1).values,
                          shard.reshape(shard.shape[0], -1).max(1).values], dim=1)
        red = nqd.allreduce_minmax(mm).numpy()
        gathered = nqd.gather_outputs(shard.numpy().transpose(1, 0, 2))
        # group=False: rank-local statistics, no collective -- only rank 0 calls it here, as bench.py's single-rank
        # configurations do; a collective issued by one rank would dead-lock against the barrier that follows
        local = nqd.allreduce_minmax(mm, group=False) if rank == 0 else mm
        assert local is mm
        dist.barrier()
        out.put((rank, red, None if gathered is None else gathered.shape))
    finally:
        dist.destroy_process_group()


def test_world_size_2_calibration_stats_are_bit_identical_to_single_process():
    """The only collective of the path: all-reduce(max) of [max, -min]."""
    rng = np.random.default_rng(0)
    full = torch.from_numpy(rng.normal(size=(37, 9, 11)).astype(np.float32))       # 37 values, 9 images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, full, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([full.reshape(37, -1).min(1).values.numpy(), full.reshape(37, -1).max(1).values.numpy()], 1)
    for rank, red, gshape in res:
        assert red.tobytes() == want.tobytes()                   # exact, on every rank
        for i in range(37):                                      # -> identical quantization parameters
            s, z = quant_parameters(np.float32(red[i, 0]), np.float32(red[i, 1]), 8, True)
            s0, z0 = quant_parameters(np.float32(want[i, 0]), np.float32(want[i, 1]), 8, True)
            assert s.tobytes() == s0.tobytes() and int(z) == int(z0)
    assert res[0][2] == (9, 37, 11) and res[1][2] is None


def test_allreduce_is_identity_without_process_group():
    mm = torch.tensor([[0.0, 1.0], [-2.0, 3.0]])
    assert nqd.allreduce_minmax(mm) is mm and nqd.world() == (0, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_a_gpu():
    from numpy_quant_b200.model import Model
    from numpy_quant_b200.tensor import FTensor
    with pytest.raises(_lib.NqError, match="no CPU fallback"):
        FTensor(np.zeros(3, np.float32))
    with pytest.raises(_lib.NqError):
        Model.from_onnx(zoo.gemm_graph(3, 4, 2))
