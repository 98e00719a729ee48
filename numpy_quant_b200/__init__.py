"""numpy_quant_b200 -- B200-native (sm_100a) implementation of numpy-quant's
quantized-inference hot path: `Model.from_onnx -> model.quantize(calib, bit_width) ->
qmodel(inputs)` and the FTensor / QTensor operator surface, executed by hand-written CUDA
kernels (libnq_b200.so) behind a C ABI.  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"


def install_as_numpy_quant() -> None:
    """Make `import numpy_quant.model / .tensor / .numpy_quantization` resolve to this
    package, so scripts written against the reference run unchanged."""
    import importlib
    import sys
    pkg = sys.modules[__name__]
    sys.modules.setdefault("numpy_quant", pkg)
    for sub in ("model", "tensor", "numpy_quantization"):
        sys.modules.setdefault(f"numpy_quant.{sub}", importlib.import_module(f"{__name__}.{sub}"))
