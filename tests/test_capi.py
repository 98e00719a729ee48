"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a
GPU or libcuda, and exports every symbol include/nq_b200.h declares."""
import os
import re

import pytest

from numpy_quant_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "nq_b200.h")).read()
    declared = set(re.findall(r"\b(nq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in nq_b200.h but not exported by libnq_b200.so"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_version_and_error_string(lib):
    assert lib.nq_version() >= 100
    assert isinstance(lib.nq_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    # rejected before any CUDA call is made
    with pytest.raises(_lib.NqError, match="bit_width"):
        _lib.call("nq_quantize_f32", None, 16, 9, 1.0, 0, 0, None, None)
    with pytest.raises(_lib.NqError, match="epilogue descriptor"):
        _lib.call("nq_qgemm_s8", None, None, None, 1, 1, 1, 1, 16, 16, 1, 0, 0, 0, None, None)


def test_conv_entry_points_validate_arguments(lib):
    ep = _lib.Epilogue()
    with pytest.raises(_lib.NqError, match="multiple of 64"):          # channels: one filter tap = whole 64-byte K slices
        _lib.call("nq_qconv2d_s8", None, None, None, 1, 8, 8, 3, 3, 3, 1, 1, 16, 32, 16, _lib.C.byref(ep), None)
    with pytest.raises(_lib.NqError, match="does not fit"):
        _lib.call("nq_qconv2d_s8", None, None, None, 1, 2, 8, 64, 3, 3, 1, 1, 16, 576, 16, _lib.C.byref(ep), None)
    with pytest.raises(_lib.NqError, match="strides"):
        _lib.call("nq_qconv2d_s8", None, None, None, 1, 8, 8, 64, 3, 3, 9, 1, 16, 576, 16, _lib.C.byref(ep), None)
    ep.mode = _lib.EPI_QUANT
    with pytest.raises(_lib.NqError, match="RAW, DEQUANT or REQUANT"):
        _lib.call("nq_qconv2d_s8", None, None, None, 1, 8, 8, 64, 3, 3, 1, 1, 16, 576, 16, _lib.C.byref(ep), None)
    with pytest.raises(_lib.NqError, match="multiple of 4"):
        _lib.call("nq_nhwc_pad", None, 1, 1, 3, 8, 8, 0, 0, 0, 0, 0, 8, 1.0, 0, 0, None, None)
    with pytest.raises(_lib.NqError, match="elem_bytes"):
        _lib.call("nq_nhwc_pad", None, 2, 1, 64, 8, 8, 0, 0, 0, 0, 0, 8, 1.0, 0, 0, None, None)


def test_no_libcuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out, out
