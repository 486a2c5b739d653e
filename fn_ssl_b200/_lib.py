"""ctypes binding of libfnssl_b200.so (include/fnssl_b200.h).  Fails loudly: there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfnssl_b200.so")
_lock = threading.Lock()
_lib = None

ABI_VERSION = 5
F32, F16 = 0, 1
ALONG_FREQ, ALONG_TIME = 0, 1
ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1
PAIRS_M, PAIRS_MM, PAIRS_ALL = 0, 1, 2
NORM_NONE, NORM_FORGETTING, NORM_GLOBAL = 0, 1, 2


class LstmArgs(C.Structure):
    """struct fnssl_lstm_args (include/fnssl_b200.h)."""
    _fields_ = [
        ("engine", C.c_int32), ("axis", C.c_int32),
        ("nb", C.c_int32), ("nt", C.c_int32), ("nf", C.c_int32),
        ("hidden", C.c_int32), ("num_dirs", C.c_int32), ("dtype", C.c_int32),
        ("src0", C.c_void_p), ("c0", C.c_int32), ("ld0", C.c_int32),
        ("src1", C.c_void_p), ("c1", C.c_int32), ("ld1", C.c_int32),
        ("weights", C.c_void_p), ("weights_bytes", C.c_int64),
        ("out0", C.c_void_p), ("out0_ld", C.c_int32), ("out0_off", C.c_int32),
        ("addend", C.c_void_p), ("addend_ld", C.c_int32),
        ("out1", C.c_void_p), ("out1_ld", C.c_int32),
        ("h_state", C.c_void_p), ("c_state", C.c_void_p), ("state_flags", C.c_int32),
    ]


_fp = C.c_void_p


class SnFconvWeights(C.Structure):
    """struct fnssl_sn_fconv_weights."""
    _fields_ = [("ln_w", _fp), ("ln_b", _fp), ("conv_wp", _fp), ("conv_b", _fp), ("prelu", _fp)]


class SnFreqArgs(C.Structure):
    """struct fnssl_sn_freq_args."""
    _fields_ = [
        ("nb", C.c_int32), ("nt", C.c_int32), ("nf", C.c_int32),
        ("hidden", C.c_int32), ("squeeze", C.c_int32), ("groups", C.c_int32), ("fkernel", C.c_int32),
        ("is_first", C.c_int32),
        ("x", _fp), ("cin", C.c_int32), ("x_ld", C.c_int32),
        ("enc_wp", _fp), ("enc_b", _fp), ("enc_kernel", C.c_int32),
        ("fconv1", SnFconvWeights),
        ("lnf_w", _fp), ("lnf_b", _fp), ("sq_wt", _fp), ("sq_b", _fp), ("full_wt", _fp), ("full_b", _fp),
        ("usq_w", _fp), ("usq_b", _fp),
        ("fconv2", SnFconvWeights),
        ("out", _fp),
        ("t_begin", C.c_int32),
    ]


class MambaWeights(C.Structure):
    """struct fnssl_mamba_weights."""
    _fields_ = [(n, _fp) for n in ("ln_w", "ln_b", "in_proj_wt", "conv_w", "conv_b", "x_proj_wt", "dt_proj_w",
                                   "dt_proj_b", "A_log", "D", "out_proj_wt")]


class SnTimeArgs(C.Structure):
    """struct fnssl_sn_time_args."""
    _fields_ = [
        ("nb", C.c_int32), ("nt", C.c_int32), ("nf", C.c_int32), ("hidden", C.c_int32),
        ("d_inner", C.c_int32), ("d_state", C.c_int32), ("dt_rank", C.c_int32), ("d_conv", C.c_int32),
        ("pool", C.c_int32),
        ("x", _fp), ("work", _fp), ("out", _fp),
        ("m", MambaWeights * 2),
        ("state", _fp * 2), ("state_flags", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/fnssl_b200.h declares
_vp, _i, _f, _i64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
SIGNATURES = {
    "fnssl_abi_version": (_i, []),
    "fnssl_last_error": (C.c_char_p, []),
    "fnssl_stft_num_frames": (_i, [_i, _i, _i]),
    "fnssl_stft_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fnssl_norm_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "fnssl_norm_stream_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.c_longlong, _vp, _vp, _vp]),
    "fnssl_feature_rows": (_i, [_i, _i, _i]),
    "fnssl_feature_channels": (_i, [_i, _i]),
    "fnssl_features_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _i, _i, _vp, _vp]),
    "fnssl_stft_features_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "fnssl_cfirst_to_grid": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "fnssl_grid_to_cfirst": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "fnssl_grid_copy": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i64, _i, _vp]),
    "fnssl_grid_add": (_i, [_vp, _vp, _vp, _i, _i64, _vp]),
    "fnssl_lstm_forward": (_i, [C.POINTER(LstmArgs), _vp]),
    "fnssl_lstm_train_saved_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "fnssl_lstm_forward_train": (_i, [C.POINTER(LstmArgs), _vp, _i64, _vp]),
    "fnssl_lstm_backward": (_i, [C.POINTER(LstmArgs), _vp, _i64, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "fnssl_ipd_head_backward": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "fnssl_lstm_tc_supported": (_i, [_i, _i, _i]),
    "fnssl_lstm_tc_kernel_for": (_i, [C.POINTER(LstmArgs)]),
    "fnssl_lstm_tc_error_site": (_i, []),
    "fnssl_lstm_tc4_trace": (_i, [C.POINTER(C.c_longlong)]),
    "fnssl_ipd_head_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fnssl_linear_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "fnssl_linear_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fnssl_doa_decode_idl": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fnssl_reflect_pad": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "fnssl_sn_freq_forward": (_i, [C.POINTER(SnFreqArgs), _vp]),
    "fnssl_sn_time_forward": (_i, [C.POINTER(SnTimeArgs), _vp]),
    "fnssl_sn_head_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "fnssl_dpipd_targets": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _i, _f, _i, _vp, _vp, _vp]),
    "fnssl_ipd_mse_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fnssl_ipd_pit_mse_loss": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fnssl_conv3x3_train_workspace_bytes": (_sz, [_i, _i]),
    "fnssl_conv3x3_forward": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "fnssl_conv3x3_backward_data": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "fnssl_conv3x3_backward_weight": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fnssl_causcnn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "fnssl_causcnn_forward": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
}


def load(build_if_missing: bool = True):
    """Load the C-ABI library (building it with nvcc if it is absent and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        # The library is git-ignored but lives in the working tree: after an edit of csrc/*.cu a plain import must not run a
        # stale binary.  build() is stamp-checked (sha256 of every source + header + flags), so when nvcc is present it is
        # called on every first load and only recompiles what changed; without nvcc the existing library must match HEAD.
        from . import build as _build
        path = os.environ.get("FNSSL_B200_LIB")            # development / profiling only: an experimental variant library
        if path:                                            # (tools/build_variant.py); never set in production
            pass
        elif build_if_missing and _build.have_nvcc():
            _build.build()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m fn_ssl_b200.build`")
        elif _build.stale_sources():
            raise RuntimeError(f"{LIB_PATH} is older than its sources ({', '.join(_build.stale_sources())}) and nvcc is not "
                               "available to rebuild it: run `python -m fn_ssl_b200.build` where nvcc exists")
        lib = C.CDLL(path or LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.fnssl_abi_version() != ABI_VERSION:
            raise RuntimeError("libfnssl_b200.so: ABI version mismatch, rebuild it")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fnssl_last_error()
        raise RuntimeError("fnssl_b200: " + (msg.decode() if msg else f"error {rc}"))
