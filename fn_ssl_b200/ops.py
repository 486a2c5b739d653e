"""Tensor-level wrappers over the C ABI (include/fnssl_b200.h).

PyTorch supplies device memory and the current CUDA stream; every computation below happens inside
libfnssl_b200.so.  Inputs that are not CUDA tensors raise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ALONG_FREQ, ALONG_TIME, ENGINE_SIMT, ENGINE_TCGEN05, F16, F32, NORM_FORGETTING, NORM_GLOBAL,
                   NORM_NONE, PAIRS_ALL, PAIRS_M, PAIRS_MM)

Tensor = torch.Tensor

_DTYPES = {torch.float32: F32, torch.float16: F16}
PAIRING = {"M": PAIRS_M, "MM": PAIRS_MM, "ALL": PAIRS_ALL}
NORM_GIVEN = 3   # FNSSL_NORM_GIVEN: mu is an input of features() / stft_features() (streaming)


# launch accounting / per-layer timing (used by bench.py; off by default)
LAUNCHES = 0                 # kernels of libfnssl_b200.so enqueued so far
TC_LSTM_LAUNCHES = 0         # ... of which LSTM layers on the tcgen05 engine (tests assert the product engine really ran)
_PROFILE = None              # list of (label, flops, bytes, start_event, end_event) when enabled


def profile_start() -> None:
    global _PROFILE
    _PROFILE = []


def profile_stop():
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    return rec or []


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def profiled(label: str, flops: float, nbytes: float, launch) -> None:
    """Run `launch()` (one kernel launch); when profiling is on, bracket it with CUDA events on the current stream."""
    if _PROFILE is None:
        launch()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    _PROFILE.append((label, flops, nbytes, e0, e1))


def _stream() -> int:
    """Current stream of the CURRENT device -- every entry point runs under `on_tensor_device`, which makes the device of its
    tensors current first, so this is the current stream of the tensors' device."""
    return torch.cuda.current_stream().cuda_stream


def on_tensor_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument as the current device.  The C ABI takes raw pointers and a
    stream: cudaFuncSetAttribute, tensor-map encodes, the launch itself and `_stream()` all act on the *current* device, so a
    model living on cuda:1 while cuda:0 is current would otherwise launch into the wrong context."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None:
            for a in kwargs.values():
                if isinstance(a, torch.Tensor) and a.is_cuda:
                    dev = a.device
                    break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _need_cuda(*tensors: Tensor) -> None:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only; got a %s tensor -- no CPU fallback exists"
                               % t.device.type)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"fn_ssl_b200: tensors of one call live on different devices ({dev} and {t.device})")


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def code_of(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise RuntimeError(f"fn_ssl_b200 grids are float32 or float16, not {dtype}")


def pad_channels(c: int, dtype: torch.dtype) -> int:
    """Channel stride of a grid holding c channels: fp16 grids are padded to a multiple of 16 channels
    (one tcgen05 K step / TMA's 16-byte rule), fp32 grids to a multiple of 4."""
    m = 16 if dtype == torch.float16 else 4
    return (c + m - 1) // m * m


# ---------------------------------------------------------------------------------------------
# front end
# ---------------------------------------------------------------------------------------------

def stft_num_frames(nsample: int, win_len: int = 512, hop: int = 256) -> int:
    return _lib.load().fnssl_stft_num_frames(nsample, win_len, hop)


@on_tensor_device
def stft(signal: Tensor, win_len: int = 512, hop: int = 256, nfft: int = 512,
         want_magsum: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """(nb, nsample, nch) f32 -> ((nb, nfft/2+1, nt, nch) complex64, magsum (nb, nch, nt) | None)."""
    _need_cuda(signal)
    lib = _lib.load()
    if signal.dim() != 3:
        raise RuntimeError("stft: expected (nbatch, nsample, nch)")
    x = signal.contiguous().float()
    nb, nsample, nch = x.shape
    nt = lib.fnssl_stft_num_frames(nsample, win_len, hop)
    if nt <= 0:
        raise RuntimeError(f"stft: signal of {nsample} samples is shorter than one {win_len}-sample window")
    spec = torch.empty((nb, nfft // 2 + 1, nt, nch, 2), dtype=torch.float32, device=x.device)
    magsum = torch.empty((nb, nch, nt), dtype=torch.float32, device=x.device) if want_magsum else None
    _count(1)
    _lib.check(lib.fnssl_stft_forward(x.data_ptr(), nb, nsample, nch, win_len, hop, nfft, spec.data_ptr(),
                                      _ptr(magsum), _stream()))
    return torch.view_as_complex(spec), magsum




@on_tensor_device
def norm_stream(magsum: Tensor, pairing: str, sample_length: int, t0: int, mu_state: Tensor) -> Tensor:
    """forgetting_norm for frames [t0, t0+nt) of a stream; mu_state (R,) f32 is read (t0 > 0) and updated in place."""
    _need_cuda(magsum, mu_state)
    lib = _lib.load()
    nb, nch, nt = magsum.shape
    pm = PAIRING[pairing]
    R = lib.fnssl_feature_rows(nb, nch, pm)
    if mu_state.shape != (R,) or mu_state.dtype != torch.float32 or not mu_state.is_contiguous():
        raise RuntimeError(f"norm_stream: mu_state must be a contiguous float32 tensor of shape ({R},)")
    mu = torch.empty((R, nt), dtype=torch.float32, device=magsum.device)
    _count(1)
    _lib.check(lib.fnssl_norm_stream_forward(magsum.contiguous().data_ptr(), nb, nch, nt, 257, pm, sample_length, int(t0),
                                             mu_state.data_ptr(), mu.data_ptr(), _stream()))
    return mu


@on_tensor_device
def features(spec: Tensor, magsum: Optional[Tensor], pairing: str, norm: int, sample_length: int, eps: float,
             dtype: torch.dtype, want_cfirst: bool = False, mu: Optional[Tensor] = None
             ) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """spec (nb,257,nt,nch) complex64 -> grid (R, nt, 256, ld) of dtype, mu (R, nt), [(R, C, 256, nt) f32].
    norm = NORM_GIVEN takes the normaliser `mu` (R, nt) from the caller (norm_stream)."""
    _need_cuda(spec)
    lib = _lib.load()
    nb, nbins, nt, nch = spec.shape
    if nbins != 257:
        raise RuntimeError("features: the path is built for 257-bin spectra (nfft = 512)")
    pm = PAIRING[pairing]
    R = lib.fnssl_feature_rows(nb, nch, pm)
    Cc = lib.fnssl_feature_channels(nch, pm)
    ld = pad_channels(Cc, dtype)
    sp = torch.view_as_real(spec.contiguous())
    feat = torch.empty((R, nt, 256, ld), dtype=dtype, device=spec.device)
    if norm == NORM_GIVEN:
        if mu is None or mu.shape != (R, nt) or mu.dtype != torch.float32 or not mu.is_contiguous():
            raise RuntimeError("features: NORM_GIVEN needs a contiguous float32 mu of shape (R, nt)")
    else:
        mu = torch.empty((R, nt), dtype=torch.float32, device=spec.device)
    cf = torch.empty((R, Cc, 256, nt), dtype=torch.float32, device=spec.device) if want_cfirst else None
    _count(2 if norm in (NORM_FORGETTING, NORM_GLOBAL) else 1)
    _lib.check(lib.fnssl_features_forward(sp.data_ptr(), _ptr(magsum), nb, nt, nch, pm, norm, sample_length, float(eps),
                                          mu.data_ptr(), feat.data_ptr(), code_of(dtype), ld, _ptr(cf), _stream()))
    return feat, mu, cf


@on_tensor_device
def stft_features(signal: Tensor, pairing: str, norm: int, sample_length: int, eps: float, dtype: torch.dtype,
                  win_len: int = 512, hop: int = 256, nfft: int = 512, mu: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Fused front end: (nb, nsample, nch) f32 -> (grid (R, nt, 256, ld) of dtype, mu (R, nt) | None); the complex spectrum is
    never written to HBM (two FFT passes over the L2-resident signal; see fnssl_stft_features_forward)."""
    _need_cuda(signal, mu)
    lib = _lib.load()
    if signal.dim() != 3:
        raise RuntimeError("stft_features: expected (nbatch, nsample, nch)")
    x = signal.contiguous().float()
    nb, nsample, nch = x.shape
    nt = lib.fnssl_stft_num_frames(nsample, win_len, hop)
    if nt <= 0:
        raise RuntimeError(f"stft_features: signal of {nsample} samples is shorter than one {win_len}-sample window")
    pm = PAIRING[pairing]
    R = lib.fnssl_feature_rows(nb, nch, pm)
    ld = pad_channels(lib.fnssl_feature_channels(nch, pm), dtype)
    feat = torch.empty((R, nt, 256, ld), dtype=dtype, device=x.device)
    magsum = None
    if norm in (NORM_FORGETTING, NORM_GLOBAL):
        magsum = torch.empty((nb, nch, nt), dtype=torch.float32, device=x.device)
        mu = torch.empty((R, nt), dtype=torch.float32, device=x.device)
    elif norm == NORM_GIVEN:
        if mu is None or mu.shape != (R, nt) or mu.dtype != torch.float32 or not mu.is_contiguous():
            raise RuntimeError("stft_features: NORM_GIVEN needs a contiguous float32 mu of shape (R, nt)")
    else:
        mu = None
    _count(3 if magsum is not None else 1)
    _lib.check(lib.fnssl_stft_features_forward(x.data_ptr(), nb, nsample, nch, win_len, hop, nfft, pm, norm, sample_length,
                                               float(eps), _ptr(magsum), _ptr(mu), feat.data_ptr(), code_of(dtype), ld, _stream()))
    return feat, mu


@on_tensor_device
def cfirst_to_grid(x: Tensor, dtype: torch.dtype, ld: Optional[int] = None, nt_alloc: Optional[int] = None) -> Tensor:
    """(nb, C, nf, nt) f32 -> grid (nb, nt_alloc or nt, nf, ld) of dtype; padding channels / frames are zero."""
    _need_cuda(x)
    lib = _lib.load()
    x = x.contiguous().float()
    nb, Cc, nf, nt = x.shape
    ld = ld or pad_channels(Cc, dtype)
    nta = nt_alloc or nt
    if ld != Cc or nta != nt:
        g = torch.zeros((nb, nta, nf, ld), dtype=dtype, device=x.device)
    else:
        g = torch.empty((nb, nta, nf, ld), dtype=dtype, device=x.device)
    _count(1 if nta == nt else nb)
    if nta == nt:
        _lib.check(lib.fnssl_cfirst_to_grid(x.data_ptr(), nb, Cc, nf, nt, g.data_ptr(), code_of(dtype), ld, 0, _stream()))
    else:  # per-utterance rows are nta frames apart
        for b in range(nb):
            _lib.check(lib.fnssl_cfirst_to_grid(x[b].data_ptr(), 1, Cc, nf, nt, g[b].data_ptr(), code_of(dtype), ld, 0,
                                                _stream()))
    return g


@on_tensor_device
def grid_to_cfirst(g: Tensor, Cc: int, ch_off: int = 0) -> Tensor:
    _need_cuda(g)
    lib = _lib.load()
    nb, nt, nf, ld = g.shape
    out = torch.empty((nb, Cc, nf, nt), dtype=torch.float32, device=g.device)
    _count(1)
    _lib.check(lib.fnssl_grid_to_cfirst(g.data_ptr(), code_of(g.dtype), ld, ch_off, nb, Cc, nf, nt, out.data_ptr(), _stream()))
    return out


@on_tensor_device
def grid_copy(src: Tensor, Cc: int, dst_dtype: torch.dtype, dst_ld: Optional[int] = None, src_off: int = 0) -> Tensor:
    """Copy / convert / re-pad the first Cc channels (from src_off) of a grid into a new grid."""
    _need_cuda(src)
    lib = _lib.load()
    src = src.contiguous()
    ld = dst_ld or pad_channels(Cc, dst_dtype)
    shape = tuple(src.shape[:-1]) + (ld,)
    dst = torch.zeros(shape, dtype=dst_dtype, device=src.device) if ld != Cc else torch.empty(shape, dtype=dst_dtype, device=src.device)
    npos = src.numel() // src.shape[-1]
    _count(1)
    _lib.check(lib.fnssl_grid_copy(src.data_ptr(), code_of(src.dtype), src.shape[-1], src_off, dst.data_ptr(),
                                   code_of(dst_dtype), ld, 0, npos, Cc, _stream()))
    return dst


@on_tensor_device
def grid_add(a: Tensor, b: Tensor) -> Tensor:
    _need_cuda(a, b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise RuntimeError("grid_add: shape/dtype mismatch")
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    _count(1)
    _lib.check(_lib.load().fnssl_grid_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), code_of(a.dtype), a.numel(), _stream()))
    return out


# ---------------------------------------------------------------------------------------------
# LSTM
# ---------------------------------------------------------------------------------------------

@on_tensor_device
def lstm(engine: int, axis: int, src0: Tensor, c0: int, src1: Optional[Tensor], c1: int, weights: Tensor, hidden: int,
         num_dirs: int, addend: Optional[Tensor] = None, want_h: bool = True,
         out0: Optional[Tensor] = None, out0_off: int = 0,
         state: Optional[Tuple[Tensor, Tensor]] = None, inplace_addend: bool = False, duplicate: bool = False
         ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """One LSTM layer over a grid (see fnssl_lstm_forward).  Returns (h grid, h + addend grid).
    inplace_addend: the sum is accumulated into `addend` itself (out1 aliases addend; tensor-core engine only) -- the
    caller must not need the residual operand afterwards and it must not be one of the layer's inputs.
    duplicate (no addend): the second result is a second copy of h, written by the kernel itself (for a consumer that reads
    h as an input AND accumulates onto a copy of it in place -- block 1's narrow-band layer).
    state = (h, c): float32 (rows, hidden) tensors the layer starts from and overwrites with its final state
    (nn.LSTM's (h_0, c_0) -> (h_n, c_n); uni-directional layers only)."""
    _need_cuda(src0, src1, weights, addend)
    lib = _lib.load()
    nb, nt, nf, ld0 = src0.shape
    oc = hidden * num_dirs
    dt = src0.dtype
    dev = src0.device
    if not src0.is_contiguous() or (src1 is not None and not src1.is_contiguous()):
        raise RuntimeError("lstm: grids must be contiguous")
    if out0 is None and want_h:
        out0 = torch.empty((nb, nt, nf, oc), dtype=dt, device=dev)
    if addend is not None and inplace_addend and engine == ENGINE_TCGEN05:
        if addend.data_ptr() in (src0.data_ptr(), src1.data_ptr() if src1 is not None else 0):
            raise RuntimeError("lstm: an in-place residual operand must not be an input of the same layer")
        if not addend.is_contiguous() or addend.shape != (nb, nt, nf, oc):
            raise RuntimeError("lstm: in-place residual operand must be a contiguous (nb, nt, nf, dirs*hidden) grid")
        out1 = addend
    else:
        if duplicate and addend is not None:
            raise RuntimeError("lstm: duplicate=True is the addend-less form of the second output")
        out1 = torch.empty((nb, nt, nf, oc), dtype=dt, device=dev) if (addend is not None or duplicate) else None
    a = _lib.LstmArgs()
    a.engine, a.axis = engine, axis
    a.nb, a.nt, a.nf = nb, nt, nf
    a.hidden, a.num_dirs, a.dtype = hidden, num_dirs, code_of(dt)
    a.src0, a.c0, a.ld0 = src0.data_ptr(), c0, ld0
    a.src1, a.c1, a.ld1 = (_ptr(src1), c1, src1.shape[-1]) if src1 is not None else (None, 0, 0)
    a.weights, a.weights_bytes = weights.data_ptr(), weights.numel() * weights.element_size()
    a.out0 = _ptr(out0)
    a.out0_ld = out0.shape[-1] if out0 is not None else 0
    a.out0_off = out0_off
    a.addend = _ptr(addend)
    a.addend_ld = addend.shape[-1] if addend is not None else 0
    a.out1 = _ptr(out1)
    a.out1_ld = oc if out1 is not None else 0
    if state is not None:
        rows = nb * nt if axis == ALONG_FREQ else nb * nf
        for s_ in state:
            if s_.shape != (rows, hidden) or s_.dtype != torch.float32 or not s_.is_contiguous() or s_.device != dev:
                raise RuntimeError(f"lstm: state tensors must be contiguous float32 ({rows}, {hidden}) on {dev}")
        a.h_state, a.c_state, a.state_flags = state[0].data_ptr(), state[1].data_ptr(), 3
    _count(1)
    if engine == ENGINE_TCGEN05:
        global TC_LSTM_LAUNCHES
        TC_LSTM_LAUNCHES += 1
    if _PROFILE is not None:
        rows, steps = (nb * nt, nf) if axis == ALONG_FREQ else (nb * nf, nt)
        flops = 2.0 * rows * steps * num_dirs * 4 * hidden * (c0 + c1 + hidden)
        esz = src0.element_size()
        nbytes = float(rows) * steps * esz * ((c0 + c1) + oc * ((1 if out0 is not None else 0) + (2 if addend is not None else 0)
                                                                + (1 if duplicate else 0)))
        label = "lstm_%s_%s_H%d_x%d_in%d" % ("tc" if engine == ENGINE_TCGEN05 else "simt",
                                             "full" if axis == ALONG_FREQ else "narrow", hidden, num_dirs, c0 + c1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.fnssl_lstm_forward(C.byref(a), _stream()))
        e1.record()
        _PROFILE.append((label, flops, nbytes, e0, e1))
    else:
        _lib.check(lib.fnssl_lstm_forward(C.byref(a), _stream()))
    return out0, out1


# ---------------------------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------------------------

@on_tensor_device
def ipd_head(x: Tensor, Cc: int, weight: Tensor, bias: Tensor) -> Tensor:
    _need_cuda(x, weight, bias)
    nb, nt, nf, ld = x.shape
    out = torch.empty((nb, nt // 12, 2 * nf), dtype=torch.float32, device=x.device)
    w = weight.detach().contiguous().float()
    b = bias.detach().contiguous().float()
    _count(1)
    _lib.check(_lib.load().fnssl_ipd_head_forward(x.data_ptr(), code_of(x.dtype), ld, nb, nt, nf, Cc, w.data_ptr(),
                                                  b.data_ptr(), out.data_ptr(), _stream()))
    return out


@on_tensor_device
def linear(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    _need_cuda(x, weight, bias)
    shp = x.shape
    x2 = x.contiguous().float().reshape(-1, shp[-1])
    w = weight.detach().contiguous().float()
    b = bias.detach().contiguous().float()
    y = torch.empty((x2.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    _count(1)
    _lib.check(_lib.load().fnssl_linear_forward(x2.data_ptr(), w.data_ptr(), b.data_ptr(), x2.shape[0], x2.shape[1],
                                                w.shape[0], y.data_ptr(), _stream()))
    return y.reshape(*shp[:-1], w.shape[0])


@on_tensor_device
def causcnn(src0: Tensor, c0: int, src1: Optional[Tensor], c1: int, w1: Tensor, w2: Tensor, w3: Tensor) -> Tensor:
    """CausCnnBlock over grid inputs -> (nb, cout, nf, nt//12) f32."""
    _need_cuda(src0, src1, w1, w2, w3)
    lib = _lib.load()
    nb, nt, nf, ld0 = src0.shape
    hid, cin = w1.shape[0], w1.shape[1]
    cout = w3.shape[0]
    if cin != c0 + c1 or tuple(w1.shape[2:]) != (3, 3):
        raise RuntimeError("causcnn: conv1 weight shape does not match the input channels / 3x3 kernel")
    ws = torch.empty(int(lib.fnssl_causcnn_workspace_bytes(nb, nt, nf, cin, hid, cout)), dtype=torch.uint8, device=src0.device)
    out = torch.empty((nb, cout, nf, (nt // 3) // 4), dtype=torch.float32, device=src0.device)
    ws_f = [w.detach().contiguous().float() for w in (w1, w2, w3)]
    _count(6)
    _lib.check(lib.fnssl_causcnn_forward(src0.data_ptr(), c0, ld0, _ptr(src1), c1, src1.shape[-1] if src1 is not None else 0,
                                         code_of(src0.dtype), nb, nt, nf, ws_f[0].data_ptr(), ws_f[1].data_ptr(),
                                         ws_f[2].data_ptr(), hid, cout, ws.data_ptr(), out.data_ptr(), _stream()))
    return out
