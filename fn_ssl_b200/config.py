"""Engine selection for the LSTM layers.

    "tcgen05" : tensor-core engine (fp16 operands, fp32 accumulate / cell state, fp16 grids)   -- the product path
    "simt"    : fp32 CUDA-core engine (fp32 grids)                                               -- exact reference engine
    "auto"    : tcgen05 whenever the layer shape is supported by it, simt otherwise (default)

Override per model (`model.engine = "simt"`) or globally with FNSSL_ENGINE.
"""
import os

import torch

DEFAULT_ENGINE = os.environ.get("FNSSL_ENGINE", "auto")
TC_AVAILABLE = True               # the tcgen05 engine is compiled into libfnssl_b200.so


def resolve(engine: str, hidden_sizes) -> str:
    engine = engine or DEFAULT_ENGINE
    if engine not in ("auto", "tcgen05", "simt"):
        raise RuntimeError(f"unknown engine {engine!r} (auto / tcgen05 / simt)")
    if engine == "auto":
        # fp16 grids + tensor cores; layers the tcgen05 kernel is not built for run the CUDA-core kernel on the
        # same grids (packing.run_lstm).  Hidden sizes the CUDA-core kernel does not know either -> error there.
        return "tcgen05" if TC_AVAILABLE else "simt"
    if engine == "tcgen05" and not TC_AVAILABLE:
        raise RuntimeError("the tcgen05 engine is not compiled into libfnssl_b200.so")
    return engine


def grid_dtype(engine: str) -> torch.dtype:
    return torch.float16 if engine == "tcgen05" else torch.float32


def engine_code(engine: str) -> int:
    return 1 if engine == "tcgen05" else 0
