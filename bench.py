#!/usr/bin/env python
"""Headline benchmark: quantized ViT-B/16 (int8, batch 256 per GPU, synthetic 224x224)
images/s through the drop-in API, per the contract in the task statement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

Own arm   : numpy_quant_b200 (sm_100a kernels).  `value` = device-resident throughput,
            `e2e` = same call with pinned-host inputs (H2D inside the timed region) and
            host logits (D2H), `roofline` = the tensor-core GEMM launches timed live with CUDA
            events (against three denominators), `sustained` = a >= 3 s loop with its own clock
            samples, `cpu_baseline` = the oracle (NumPy restatement of the reference) on a
            bounded sample, `extra_configs` = BASELINE configs 3 / 4 / 5 measured in the same
            process with their own clock samples and the reference's NumPy routines beside them.
Reference : `--impl reference` times the reference's CPU algorithm (oracle port: int64
            np.matmul etc.) on the host cores on a bounded sample of the same workload.
Under torchrun each rank processes its own 256-image shard (weak scaling, no collective in
the forward); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "images/s"
BATCH = int(os.environ.get("NQ_BENCH_BATCH", "256"))
BITS = int(os.environ.get("NQ_BENCH_BITS", "8"))   # 4 / 2: BASELINE config 4 (packed sub-byte weights), diagnostic runs only
METRIC = f"quantized ViT-B/16 int{BITS} inference throughput"
VIT = dict(image_size=224, patch_size=16, hidden=768, heads=12, intermediate=3072, layers=12, classes=1000)
GOP_PER_IMAGE = 34.90                      # integer MatMul+Gemm ops per image, 2*MACs
NOMINAL_INT8_TOPS = 4500.0                 # B200 dense int8 (BASELINE north_star names it)


def workload_config(n_gpus: int) -> dict:
    name = ("configs[1]: ViT-B/16 image classifier, int8 QModel" if BITS == 8 else
            f"configs[3]: ViT-B/16 image classifier, int{BITS} QModel (codes packed to {BITS} bits in HBM)")
    return {"workload": f"{name}, batch {BATCH} per GPU, synthetic 224x224",
            "bit_width": BITS, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "image": "3x224x224",
            "graph": "zoo.vit_graph (= models/vit/vit_image_classifier_no_weights.onnx topology, 516 nodes)",
            "execution": "qmodel(inputs, retain=False, graph=True): fused interpreter captured into one CUDA graph; e2e: "
                         "qmodel.submit(host inputs).result(), two submissions in flight (copies overlap kernels)",
            "weights": "synthetic N(0,0.02), default_rng(0)", "parallelism": f"dp{n_gpus} (batch shards, no collective)",
            "l2_policy": "inputs+activations per step (>= 154 MB in, ~4 GB touched) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded sample (1 image; 0-layer and 1-layer ViT-B/16 graphs)
# ------------------------------------------------------------------------------------------
class CpuSample:
    """Times the reference algorithm on 1 image of ViT-B/16 geometry through a graph with 0 and
    with 1 encoder layer; a full 12-layer image costs t0 + 12 * (t1 - t0)."""
    sample = ("1 image x (stem+head graph, and stem+1 encoder layer+head graph) of ViT-B/16 int8 via the oracle port "
              "(int64 np.matmul, single-threaded); images/s = 1 / (t_stem + 12 * t_layer)")

    def __init__(self, full_depth: bool = False):
        from numpy_quant_b200 import onnx_lite as ol, zoo
        from oracle import ref_graph as rg
        self.rg = rg
        rng = np.random.default_rng(1)
        self.x = rng.normal(size=(1, 3, 224, 224)).astype(np.float32)
        self.plans = []
        for layers in ((0, 1, 12) if full_depth else (0, 1)):
            cfg = dict(VIT, layers=layers)
            g = rg.import_graph(zoo.vit_graph(batch=1, seed=0, **cfg), ol)
            self.plans.append(rg.calibrate(g, [self.x], BITS))

    def step(self) -> float:
        """One timed sample -> seconds per image, extrapolated to the full model."""
        ts = []
        for plan in self.plans[:2]:
            t0 = time.perf_counter()
            self.rg.run_quant(plan, [self.x])
            ts.append(time.perf_counter() - t0)
        t_stem, t_layer = ts[0], max(ts[1] - ts[0], 1e-9)
        return t_stem + 12 * t_layer               # seconds per image

    def full(self) -> float:
        """Seconds for ONE full 12-layer image (validates the extrapolation of `step`)."""
        t0 = time.perf_counter()
        self.rg.run_quant(self.plans[2], [self.x])
        return time.perf_counter() - t0


_WORKER_SAMPLE = None


def _ref_worker_init():
    global _WORKER_SAMPLE
    import warnings
    warnings.simplefilter("ignore")
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    _WORKER_SAMPLE = CpuSample()


def _ref_worker_step(_):
    return _WORKER_SAMPLE.step()


def _ref_full_image(_):
    import warnings
    warnings.simplefilter("ignore")
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    s = CpuSample(full_depth=True)
    return s.step(), s.full()


def run_reference_arm(args) -> None:
    """The reference's CPU algorithm (oracle port) on all host cores the process may use: the int64 np.matmul at its
    heart is single-threaded, so the cores are filled with independent images (one worker process per core, one
    image each per step); throughput = sum over workers of 1 / (seconds per image).  The per-image time is an
    extrapolation (stem + 12 x one layer); it is validated once per run against one full 12-layer image."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import warnings
    warnings.simplefilter("ignore")
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, int(os.environ.get("NQ_REF_MAX_WORKERS", "64"))))
    ctx = mp.get_context("fork")
    validation = None
    if not os.environ.get("NQ_REF_SKIP_VALIDATION"):
        with ctx.Pool(1) as one:                       # alone on the machine: one full image vs its own extrapolation
            est, full = one.map(_ref_full_image, [0])[0]
        validation = {"full_12_layer_image_s": full, "extrapolated_s": est, "ratio_full_over_extrapolated": full / est,
                      "note": "one worker, one image, all 12 layers through the oracle vs t_stem + 12 * t_layer of the same worker"}
    with ctx.Pool(cores, initializer=_ref_worker_init) as pool:
        for _ in range(max(0, min(args.warmup, 1))):     # one warm-up sample is enough for NumPy
            pool.map(_ref_worker_step, range(cores))
        t0 = time.perf_counter()
        per_step = [pool.map(_ref_worker_step, range(cores)) for _ in range(args.steps)]
        wall = time.perf_counter() - t0
    value = float(np.mean([sum(1.0 / s for s in secs) for secs in per_step]))
    sample = CpuSample.sample + f"; {cores} worker processes (one image each per step, single-threaded BLAS)"
    cfg = workload_config(args.gpus)
    cfg["workload"] = (f"ViT-B/16 int{BITS} through the reference algorithm (oracle port) on the host: 1 image per worker, stem + 1 "
                       f"encoder layer measured, x12 extrapolated, {cores} workers -- the same graph, weights and image "
                       f"shape as the own arm's `{workload_config(args.gpus)['workload']}`")
    cfg["execution"] = "oracle/ref_graph.run_quant (NumPy int64 matmul), one forked worker per host core"
    cfg["parallelism"] = f"{cores} host processes"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64 (NumPy) / f32", "data": "synthetic",
            "impl": "reference", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extrapolation_check": validation,
            "gpu_launches": 0}
    emit_json(line)


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows),
                "power_w": float(np.median(pw)) if pw else None}


def load_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def tops_fractions(tops: float, peaks: dict, lib_tops) -> dict:
    """One achieved int8 TOPS figure against every denominator in use (none of them is hidden)."""
    sus, burst = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
    return {"of_2x_measured_bf16_sustained": tops / (2 * sus) if sus else None,
            "of_2x_measured_bf16_burst": tops / (2 * burst) if burst else None,
            "of_library_int8_gemm_same_run": tops / lib_tops if lib_tops else None,
            "of_nominal_dense_int8_4500": tops / NOMINAL_INT8_TOPS}


def library_int8_tops(torch, dev):
    """cuBLASLt int8 GEMM through torch._int_mm, 8192^3, on the same tensor pipe in the same run (a measured int8 peak)."""
    try:
        a8 = torch.randint(-128, 127, (8192, 8192), device=dev, dtype=torch.int8)
        for _ in range(3):
            torch._int_mm(a8, a8)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(10):
            torch._int_mm(a8, a8)
        l1.record()
        torch.cuda.synchronize()
        return 10 * 2.0 * 8192 ** 3 / (l0.elapsed_time(l1) * 1e-3) / 1e12
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# extra configs (BASELINE configs 3 / 4 / 5), measured in the same process
# ------------------------------------------------------------------------------------------
def _timed_flush(torch, fn, flush, iters=8, warmup=3):
    """Median ms of `fn` alone: CUDA events on the launching stream, L2 flushed (256 MB write) before every call."""
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def extra_config5(torch, dev, local, peaks, lib_tops) -> dict:
    """configs[4]: standalone qGEMM 4096^3 at bit_width 2..8 (A asymmetric, B symmetric; RAW / DEQUANT / REQUANT
    epilogues) and quantize / dequantize / requantize / pack on float32 [4096, 4096], with the reference's NumPy
    routines (oracle port of numpy_quantization.py:24-72) timed on the host on identical arrays."""
    from numpy_quant_b200 import _lib, kernels as K
    from oracle import ref_quant as rq
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    n = 4096
    gemm = []
    for bits in range(2, 9):
        lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
        rng = np.random.default_rng(0)                                       # default_rng(0).integers(lo, hi + 1)
        a = torch.from_numpy(rng.integers(lo, hi + 1, size=(1, n, n)).astype(np.int8)).to(dev)
        b = torch.from_numpy(rng.integers(lo, hi + 1, size=(1, n, n)).astype(np.int8)).to(dev)
        oa, ob = K.operand_from_codes(a, "A", False), K.operand_from_codes(b, "B", True)
        azp = K.AccZeroPoint(3 if bits > 2 else 1, None, n, None, ob.rowsum, True)
        row = {"bit_width": bits}
        for name, kw in (("raw_s32", dict(mode=_lib.EPI_RAW)),
                         ("dequant_f32", dict(mode=_lib.EPI_DEQUANT, scale=1e-4, azp=azp)),
                         ("requant", dict(mode=_lib.EPI_REQUANT, scale=1e-4, azp=azp, out_bits=bits, out_scale=0.05 * 255 / (hi - lo), out_zp=-1))):
            ms = _timed_flush(torch, lambda: K.qgemm(oa, ob, **kw), flush, iters=5, warmup=2)
            tops = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
            row[name] = {"ms": ms, "tops": tops, "frac": tops_fractions(tops, peaks, lib_tops)}
        gemm.append(row)
        del a, b, oa, ob
    # quantization kernels on [4096, 4096] (algorithmic bytes of SURVEY 8d)
    x_host = np.random.default_rng(0).normal(size=(n, n)).astype(np.float32)
    x = torch.from_numpy(x_host).to(dev)
    quant = {}

    def put(name, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        quant[name] = {"ms": ms, "gb_s": gbs, "frac_of_measured_hbm": gbs / hbm}

    ne = n * n
    put("quantize_f32_to_s8_asym", _timed_flush(torch, lambda: K.quantize(x, 8, 0.03, -5), flush), 5 * ne)
    put("quantize_f32_to_s8_sym", _timed_flush(torch, lambda: K.quantize(x, 8, 0.03, None), flush), 5 * ne)
    put("quantize_to_gemm_operand_with_rowsum", _timed_flush(torch, lambda: K.quantize_operand(x, "A", 8, 0.03, -5, True), flush), 5 * ne)
    q8 = K.quantize(x, 8, 0.03, -5)
    put("dequantize_s8_to_f32", _timed_flush(torch, lambda: K.dequantize(q8, 0.03, -5), flush), 5 * ne)
    acc = torch.randint(-(1 << 20), 1 << 20, (1, n, n), device=dev, dtype=torch.int32)
    colsum = torch.randint(-5000, 5000, (1, n), device=dev, dtype=torch.int32)
    azp = K.AccZeroPoint(3, None, n, None, colsum, True)
    put("dequantize_s32_acc_to_f32", _timed_flush(torch, lambda: K.dequantize_acc(acc, 1e-4, azp), flush), 8 * ne)
    put("requantize_s32_acc_to_s8", _timed_flush(torch, lambda: K.requantize_acc(acc, 1e-4, azp, None, 8, 0.05, -3), flush), 5 * ne)
    for bits in (4, 2):
        qb = K.quantize(x, bits, 0.5, -1).reshape(-1)
        put(f"pack_s8_to_{bits}bit", _timed_flush(torch, lambda: K.pack(qb, bits), flush), (1 + bits / 8) * ne)
        pk = K.pack(qb, bits)
        put(f"unpack_{bits}bit_to_s8", _timed_flush(torch, lambda: K.unpack(pk, ne, bits), flush), (1 + bits / 8) * ne)
    # the same kernels at the ViT-B batch-256 activation size [50432, 3072] (618 MB in): where a launch is long enough for
    # the fixed ~8 us of launch + ramp-up + tail of a [4096, 4096] call (13 us of traffic at peak) not to dominate
    big = {}
    xb = torch.randn(50432, 3072, device=dev)
    nb_ = xb.numel()
    for name, fn, byts in (("quantize_f32_to_s8_asym", lambda: K.quantize(xb, 8, 0.03, -5), 5 * nb_),
                           ("quantize_to_gemm_operand_with_rowsum", lambda: K.quantize_operand(xb, "A", 8, 0.03, -5, True), 5 * nb_)):
        msb = _timed_flush(torch, fn, flush, iters=5)
        big[name] = {"ms": msb, "gb_s": byts / (msb * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (msb * 1e-3) / 1e9 / hbm}
    accb = torch.randint(-(1 << 20), 1 << 20, (1, 50432, 3072), device=dev, dtype=torch.int32)
    csb = torch.randint(-5000, 5000, (1, 3072), device=dev, dtype=torch.int32)
    azb = K.AccZeroPoint(3, None, 768, None, csb, True)
    for name, fn, byts in (("dequantize_s32_acc_to_f32", lambda: K.dequantize_acc(accb, 1e-4, azb), 8 * nb_),
                           ("requantize_s32_acc_to_s8", lambda: K.requantize_acc(accb, 1e-4, azb, None, 8, 0.05, -3), 5 * nb_)):
        msb = _timed_flush(torch, fn, flush, iters=5)
        big[name] = {"ms": msb, "gb_s": byts / (msb * 1e-3) / 1e9, "frac_of_measured_hbm": byts / (msb * 1e-3) / 1e9 / hbm}
    q4 = K.quantize(xb, 4, 0.5, -1).reshape(-1)
    msb = _timed_flush(torch, lambda: K.pack(q4, 4), flush, iters=5)
    big["pack_s8_to_4bit"] = {"ms": msb, "gb_s": 1.5 * nb_ / (msb * 1e-3) / 1e9, "frac_of_measured_hbm": 1.5 * nb_ / (msb * 1e-3) / 1e9 / hbm}
    del xb, accb, q4
    torch.cuda.empty_cache()
    clocks = sampler.stop()
    # ---- the reference's NumPy routines on the host, identical arrays (single thread: NumPy ufuncs / int64 matmul)
    cpu = {}
    s8, zp = np.float32(0.03), np.int64(-5)
    t0 = time.perf_counter(); qh = rq.quantize(x_host, 8, s8, zp); t1 = time.perf_counter()
    cpu["quantize_4096x4096_s"] = t1 - t0
    t0 = time.perf_counter(); rq.dequantize(qh, s8, zp); t1 = time.perf_counter()
    cpu["dequantize_4096x4096_s"] = t1 - t0
    acc_h = acc[0].cpu().numpy().astype(np.int64)
    zp_h = colsum.cpu().numpy().astype(np.int64) * 3
    t0 = time.perf_counter(); rq.requantize(acc_h, np.float32(1e-4), zp_h, np.float32(0.05), np.int64(-3), 8); t1 = time.perf_counter()
    cpu["requantize_4096x4096_s"] = t1 - t0
    m = 1024
    rng = np.random.default_rng(0)
    ah, bh = rng.integers(-128, 128, size=(m, m)).astype(np.int64), rng.integers(-128, 128, size=(m, m)).astype(np.int64)
    t0 = time.perf_counter(); rq.q_matmul(ah, np.float32(0.02), np.int64(3), bh, np.float32(0.01), None); t1 = time.perf_counter()
    cpu["q_matmul_1024_cubed_s"] = t1 - t0
    cpu["q_matmul_4096_cubed_s_extrapolated"] = (t1 - t0) * 64
    cpu["q_matmul_gops"] = 2.0 * m ** 3 / (t1 - t0) / 1e9
    cpu["note"] = ("oracle port of numpy_quantization.py:24-72 (NumPy, one thread); q_matmul measured at 1024^3 (int64 np.matmul), "
                   "4096^3 = 64 x that (same cubic kernel, no blocking in NumPy's integer matmul)")
    best8 = next(r for r in gemm if r["bit_width"] == 8)
    speed = {"qgemm_4096_vs_numpy_extrapolated": cpu["q_matmul_4096_cubed_s_extrapolated"] / (best8["raw_s32"]["ms"] * 1e-3),
             "quantize_vs_numpy": cpu["quantize_4096x4096_s"] / (quant["quantize_f32_to_s8_asym"]["ms"] * 1e-3),
             "requantize_vs_numpy": cpu["requantize_4096x4096_s"] / (quant["requantize_s32_acc_to_s8"]["ms"] * 1e-3)}
    return {"workload": "configs[4]: qGEMM 4096^3 at bit_width 2..8 + quantize / dequantize / requantize / pack on f32 [4096, 4096]",
            "timing": "each kernel alone, CUDA events, L2 flushed (256 MB write) before every call, median",
            "qgemm_4096_cubed": gemm, "quant_kernels_4096x4096": quant, "quant_kernels_50432x3072": big, "numpy_host_baseline": cpu, "gpu_over_numpy": speed,
            "clocks": clocks}


def extra_config3(torch, dev, local, peaks, lib_tops, steps: int) -> dict:
    """configs[2]: Conv2d block (test_conv2d geometry scaled up: 64 -> 128 channels, 57x58, kernel (3,2), pads (0,2,2,1),
    strides (2,1)), int8, batch 1024, through Model.from_onnx / quantize / QModel: quantize -> padded NHWC, implicit-GEMM
    conv (im2col-mode TMA) with dequantize + bias."""
    from numpy_quant_b200 import kernels as K, onnx_lite as ol, zoo
    from numpy_quant_b200.model import Model
    from oracle import ref_graph as rg
    nb = 1024
    proto = zoo.conv_graph(nb, 64, (57, 58), 128, (3, 2), (0, 2, 2, 1), (2, 1), seed=0)
    x = torch.from_numpy(np.random.default_rng(0).normal(size=(nb, 64, 57, 58)).astype(np.float32)).to(dev)
    model = Model.from_onnx(proto)
    q = model.quantize([x[:8]], bit_width=8, group=False)    # rank 0 only: no statistics exchange
    model.release()
    q.release()
    for _ in range(3):
        q([x], retain=False, device_outputs=True, graph=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        q([x], retain=False, device_outputs=True, graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    K.GEMM_TIMER = []
    q([x], retain=False, device_outputs=True)
    torch.cuda.synchronize()
    timer, K.GEMM_TIMER = K.GEMM_TIMER, None
    clocks = sampler.stop()
    gemm_ms = sum(e[1].elapsed_time(e[2]) for e in timer)
    gop = 2.0 * (nb * 29 * 60) * 128 * 384 / 1e9
    tops = gop / gemm_ms if gemm_ms > 0 else None            # GOP / ms == TOPS
    # NumPy reference (fake-quant float conv, model.py:95-100 through the oracle) on 4 images
    xs = x[:4].cpu().numpy()
    plan = rg.calibrate(rg.import_graph(zoo.conv_graph(4, 64, (57, 58), 128, (3, 2), (0, 2, 2, 1), (2, 1), seed=0), ol), [xs], 8)
    t0 = time.perf_counter(); rg.run_quant(plan, [xs]); t1 = time.perf_counter()
    cpu_ips = 4 / (t1 - t0)
    del q, x
    torch.cuda.empty_cache()
    return {"workload": "configs[2]: Conv2d block int8, batch 1024, x[1024,64,57,58] w[128,64,3,2] pads (0,2,2,1) strides (2,1) -> [1024,128,29,60]",
            "images_s": nb / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
            "implicit_gemm": {"ms": gemm_ms, "gop": gop, "tops": tops, "frac": tops_fractions(tops, peaks, lib_tops) if tops else None,
                              "note": "nq_qconv2d_s8 alone (CUDA events, eager pass); the float32 output write (876 MB) bounds it"},
            "numpy_host_baseline": {"images_s": cpu_ips, "sample": "4 images through the oracle (quantize, float im2col conv on dequantized codes), one thread"},
            "clocks": clocks}


def extra_config4(torch, dev, local, rank, bits: int, steps: int) -> dict:
    """configs[3]: ViT-B/16 with bit_width 4 / 2 (activations and weights, reference semantics), weights stored packed in
    HBM, 512 images per GPU (batch 4096 over 8 GPUs), sharded calibration with the min/max all-reduce."""
    from numpy_quant_b200 import zoo
    from numpy_quant_b200.model import Model
    nb = 512
    model = Model.from_onnx(zoo.vit_graph(batch=nb, seed=0, **VIT))
    x = torch.from_numpy(np.random.default_rng(100 + rank).normal(size=(nb, 3, 224, 224)).astype(np.float32)).to(dev)
    q = model.quantize([x], bit_width=bits, keep_values=False)
    model.release()
    q.release()
    info = q.pack_weights()
    torch.cuda.empty_cache()
    for _ in range(3):
        q([x], retain=False, device_outputs=True, graph=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = q([x], retain=False, device_outputs=True, graph=True)[0]
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    finite = bool(torch.isfinite(out).all())
    del q, x, out
    torch.cuda.empty_cache()
    return {"bit_width": bits, "images_per_gpu": nb, "ms_per_step": ms, "images_s_this_gpu": nb / (ms * 1e-3), "steps": steps,
            "packed_weight_bytes": info["resident_bytes"], "int8_weight_bytes": info["int8_bytes"], "finite_outputs": finite, "clocks": clocks}


def run_own_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from numpy_quant_b200 import distributed as nqd, kernels as K, zoo
    from numpy_quant_b200.model import Model

    rank, local, ws = nqd.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    peaks = load_peaks()

    # ---- model: synthetic ViT-B/16, calibrated on this rank's shard, stats all-reduced ----------
    proto = zoo.vit_graph(batch=BATCH, seed=0, **VIT)
    model = Model.from_onnx(proto)
    rng = np.random.default_rng(1 + rank)
    x_host = torch.from_numpy(rng.normal(size=(BATCH, 3, 224, 224)).astype(np.float32)).pin_memory()
    x_dev = x_host.to(dev)
    # NCCL all-reduce(min/max) inside when ws > 1; activations of the calibration pass are freed as it goes
    qmodel = model.quantize([x_dev], bit_width=BITS, keep_values=False)
    model.release()
    qmodel.release()
    if BITS < 8:
        qmodel.pack_weights()
    torch.cuda.empty_cache()

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph

    def step_device():
        return qmodel([x_dev], retain=False, device_outputs=True, graph=use_graph)[0]

    def step_e2e():
        # pinned host inputs -> H2D inside the call, logits back on the host (D2H + sync)
        return qmodel([x_host], retain=False, graph=use_graph)[0]

    def step_eager_instrumented():
        return qmodel([x_dev], retain=False, device_outputs=True)[0]

    # ---- device-resident throughput ---------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = K.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = K.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    ms_rank = ms
    # ---- roofline pass: the same K steps through the eager interpreter with every tensor-core GEMM
    #      launch bracketed by CUDA events on the launching stream (events cannot sit inside a graph)
    step_eager_instrumented()
    barrier()
    K.GEMM_TIMER = []
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(args.steps):
        step_eager_instrumented()
    r1.record()
    barrier()
    timer, K.GEMM_TIMER = K.GEMM_TIMER, None
    eager_ms = r0.elapsed_time(r1)
    # ---- end to end (host buffers) ---------------------------------------------------------
    # serving loop: two submissions in flight, so the H2D copy of step k+1 and the D2H copy of step k-1 overlap
    # the kernels of step k; every step's copies and its host-visible logits are inside the timed region
    for _ in range(min(args.warmup, 3)):
        step_e2e()
    pending = None
    for _ in range(2):
        nxt = qmodel.submit([x_host])
        if pending is not None:
            pending.result()
        pending = nxt
    pending.result()
    barrier()
    t0 = time.perf_counter()
    pending = None
    for _ in range(args.steps):
        nxt = qmodel.submit([x_host])
        if pending is not None:
            logits = pending.result()[0]
        pending = nxt
    logits = pending.result()[0]
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- sustained: the same device-resident step for >= 3 s, with its own clock record ----------
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(np.ceil(3000.0 / max(ms / args.steps, 1e-3))))
        barrier()
        s_sampler = ClockSampler(local)
        s_sampler.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_sus):
            step_device()
        s1.record()
        barrier()
        sus_ms = s0.elapsed_time(s1)
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_rank": sus_ms, "clocks": s_sampler.stop()}

    # ---- the one collective of the path, timed on its own: all-reduce(MAX) of the calibration statistics ----
    allreduce_us = None
    n_stats = 2 * len(qmodel.quant_params)
    if ws > 1:
        mm = torch.randn(len(qmodel.quant_params), 2, device=dev)
        for _ in range(5):
            nqd.allreduce_minmax(mm)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            nqd.allreduce_minmax(mm)
        a1.record()
        torch.cuda.synchronize()
        allreduce_us = a0.elapsed_time(a1) * 1e3 / 20

    # ---- strong scaling datum: the SAME global batch of 256 images split over the ranks ----------
    strong = None
    if ws > 1 and BATCH % ws == 0 and not args.no_extra:
        sb = BATCH // ws
        m2 = Model.from_onnx(zoo.vit_graph(batch=sb, seed=0, **VIT))
        xs = x_dev[:sb].contiguous()
        q2 = m2.quantize([xs], bit_width=BITS, keep_values=False)
        m2.release()
        q2.release()
        for _ in range(3):
            q2([xs], retain=False, device_outputs=True, graph=True)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            q2([xs], retain=False, device_outputs=True, graph=True)
        g1.record()
        barrier()
        strong = g0.elapsed_time(g1)
        del q2, xs
        torch.cuda.empty_cache()

    # ---- extra configs on every rank (config 4 is a multi-GPU configuration), before the reductions ----
    extra = {}
    lib_tops = library_int8_tops(torch, dev) if rank == 0 else None
    if not args.no_extra:
        del qmodel
        torch.cuda.empty_cache()
        c4 = [extra_config4(torch, dev, local, rank, b, max(3, min(args.steps, 6))) for b in (4, 2)]
        if ws > 1:
            for c in c4:
                t = torch.tensor([c["ms_per_step"]], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                c["ms_per_step_max_over_ranks"] = float(t[0])
        for c in c4:
            worst = c.get("ms_per_step_max_over_ranks", c["ms_per_step"])
            c["images_s_all_gpus"] = c["images_per_gpu"] * ws / (worst * 1e-3)
        extra["config4_vit_int4_int2_packed"] = {
            "workload": f"configs[3]: ViT-B/16 int4 / int2, packed weights, 512 images per GPU x {ws} GPU(s) = global batch {512 * ws} "
                        "(4096 at 8 GPUs), calibration sharded with the min/max all-reduce", "runs": c4}

    per_rank = [ms_rank]
    if ws > 1:
        t = torch.tensor([ms, e2e_s * 1e3, sustained["ms_rank"] if sustained else 0.0, strong or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        sus_ms_max, strong_ms = float(t[2]), float(t[3])
        allr = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(ws)]
        dist.all_gather(allr, torch.tensor([ms_rank], dtype=torch.float64, device=dev))
        per_rank = [float(v[0]) for v in allr]
    else:
        e2e_ms = e2e_s * 1e3
        sus_ms_max, strong_ms = (sustained["ms_rank"] if sustained else 0.0), 0.0

    # single-GPU configurations: rank 0 alone, after the job's last collective (nothing below may touch the process group)
    if rank == 0 and not args.no_extra:
        extra["config3_conv_block_b1024"] = extra_config3(torch, dev, local, peaks, lib_tops, max(3, min(args.steps, 10)))
        extra["config5_microbench_4096"] = extra_config5(torch, dev, local, peaks, lib_tops)

    # timer entries: (ops, start, end[, tag]); the fused attention kernel is reported next to the GEMM family
    gemm = [e for e in timer if len(e) == 3]
    attn = [e for e in timer if len(e) == 4]
    gemm_ops = sum(e[0] for e in gemm)
    gemm_ms = sum(e[1].elapsed_time(e[2]) for e in gemm)
    attn_ops = sum(e[0] for e in attn)
    attn_ms = sum(e[1].elapsed_time(e[2]) for e in attn)
    if rank != 0:
        return
    bf16 = peaks.get("bf16_tflops_sustained")
    peak_tops = 2.0 * bf16 if bf16 else 2.0 * 1400.0
    # DRAM traffic per GEMM launch from the committed ncu --set full captures (launch-weighted mean over the
    # captured shapes of the step); None if the summary file is missing
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_dram_traffic.json")))
        n = sum(tr["launches_per_step"].values())
        traffic = sum(tr["per_launch_bytes"][k] * c for k, c in tr["launches_per_step"].items()) / n
        traffic_note = (f"mean DRAM bytes per launch over the {n} of {len(gemm) // max(args.steps, 1)} GEMM launches per step "
                        "whose shape has an ncu --set full capture (profiles/r02_gemm_dram_traffic.json)")
    except Exception:
        pass
    achieved = gemm_ops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    images = BATCH * ws * args.steps
    value = images / (ms * 1e-3)
    whole_tops = GOP_PER_IMAGE * 1e9 * images / (ms * 1e-3) / 1e12
    attn_tops = (attn_ops / (attn_ms * 1e-3) / 1e12) if attn_ms > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8 x int8 -> int32 (tcgen05 kind::i8), f32 glue", "data": "synthetic",
        "config": workload_config(ws),
        "e2e": {"value": images / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4) * ws,
                "d2h_bytes_per_step": int(np.asarray(logits).nbytes) * ws, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TFLOP/s",
                     "frac": (achieved / peak_tops) if achieved else None, "traffic": traffic, "traffic_note": traffic_note,
                     "frac_all_denominators": tops_fractions(achieved, peaks, lib_tops) if achieved else None,
                     "library_int8_tops_8192": lib_tops,
                     "bound_note": ("the int8 GEMMs of this graph carry fused epilogues (dequantize / GELU / quantize for the next "
                                    "MatMul) and are epilogue-bound: ablation and per-tile clock traces (profiles/r02_gemm_ablation.md, "
                                    "r02_gemm_tile_trace.md) put the MLP-1 launch at 175 us with the main loop switched off against "
                                    "189 us complete; the epilogue runs at ~0.7 warp instructions per cycle per scheduler, the rate "
                                    "its FFMA2 / ALU / MUFU mix issues at (r02_pipe_rates.md), so instructions per element are the "
                                    "lever (r02_fc1_gelu_ncu_full.md); the float32 + residual epilogues (output projection, MLP-2) "
                                    "are bound by the residual stream"),
                     "kernel": "nq::qgemm_kernel<BN, epilogue> (all instantiations launched in the step: the int8 GEMMs)",
                     "launches_per_step": len(gemm) // max(args.steps, 1),
                     "share_of_step": gemm_ms / eager_ms if eager_ms else None,
                     "attention_kernel": {"kernel": "nq::attn::attn_kernel (QK^T + softmax + P.V + merge heads)",
                                          "launches_per_step": len(attn) // max(args.steps, 1),
                                          "share_of_step": attn_ms / eager_ms if eager_ms else None,
                                          "us_per_launch": 1e3 * attn_ms / max(len(attn), 1),
                                          "achieved_tops": attn_tops,
                                          "note": "softmax on the CUDA cores; zero-point terms and both MatMuls on the tensor pipe"},
                     "whole_step": {"int_tops": whole_tops, "frac": tops_fractions(whole_tops, peaks, lib_tops)},
                     "measured_in": "eager pass of the same steps (events cannot be recorded inside a CUDA graph)",
                     "eager_ms_per_step": eager_ms / args.steps,
                     "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops_sustained (int8 = 2x bf16 on the tensor pipe; "
                                     "the file has no int8 entry); every other denominator is in frac_all_denominators") if bf16 else
                                    "fallback 2 x 1400 (MEASURED_PEAKS.json absent)",
                     "int_ops_per_image": GOP_PER_IMAGE * 1e9},
        "int_tops_whole_step": whole_tops,
    }
    if sustained:
        line["sustained"] = {"images_s": BATCH * ws * sustained["steps"] / (sus_ms_max * 1e-3), "seconds": sus_ms_max * 1e-3,
                             "steps": sustained["steps"], "clocks": sustained["clocks"]}
    if ws > 1:
        line["multi_gpu"] = {"per_rank_ms_per_step": {"min": min(per_rank) / args.steps, "max": max(per_rank) / args.steps},
                             "note": "no collective in the timed region: the spread is per-GPU clocks under sw_power_cap, not communication",
                             "calibration_allreduce_us": allreduce_us, "calibration_allreduce_floats": n_stats,
                             "strong_scaling": {"global_batch": BATCH, "batch_per_gpu": BATCH // ws, "ms_per_step": strong_ms / args.steps,
                                                "images_s": (BATCH * args.steps / (strong_ms * 1e-3)) if strong_ms else None}}
    if extra:
        line["extra_configs"] = extra
    if ws == 1 and not args.no_cpu_baseline:
        import warnings
        warnings.simplefilter("ignore")
        cs = CpuSample()
        sec = cs.step()
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": 1, "kind": "port", "sample": cs.sample}
    emit_json(line)


_JSON_FD = None


def emit_json(line: dict) -> None:
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main() -> None:
    # stdout carries exactly one JSON line: everything else that writes to file descriptor 1 (NCCL prints its
    # "NCCL version ..." banner there) is sent to stderr for the duration of the run
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager interpreter instead of CUDA-graph replay")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs block (configs 3 / 4 / 5) and the strong-scaling datum")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 3 s sustained loop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
