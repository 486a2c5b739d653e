"""The training step on the GPU: DP-IPD regression targets, the losses, and the layers with their
backward passes.

    dpipd_targets(...)        DPIPD.forward(source_doa) + the ground-truth branch of data_preprocess
                              (FN-SSL/Lightning/Module.py:464-497, main.py:227-265; IPDnet/runIPDnetOn.py:256-290)
    ipd_mse_loss(...)         cal_loss, FN-SSL/Lightning/main.py:191-198
    ipd_pit_mse_loss(...)     frame-level PIT loss, IPDnet/runIPDnetOn.py:188-206

    lstm_layer(...)           one LSTM layer over a grid WITH its backward pass (torch.autograd.Function over
                              fnssl_lstm_forward_train / fnssl_lstm_backward): what autograd does behind nn.LSTM in the
                              reference's training_step (FN-SSL/Lightning/main.py:95-109 through Model.py:38,46)
    ipd_head_train(...)       AvgPool(12) -> Linear(256,2) -> tanh -> [cos | sin] with its backward (Model.py:79-87)
    conv3x3_causal(...), causcnn_train(...)   IPDnet's causal conv block with its backward (IPDnet/FixedAarryIPDnet.py:42-73)

Scope: targets are built per batch on the host in the reference (float64 numpy loops + a host->device copy); here they and the
losses are CUDA kernels.  The backward pass exists for the fp32 engine (lstm_train.cu, head.cu): `FN_SSL` / `FNblock` in train
mode run it (fn_ssl_b200/Model.py: residual adds and dropout are torch elementwise ops between the layer kernels), and so does
`IPDnet` (conv_train.cu; ReLU / pooling / tanh are torch ops).  The tensor-core kernels are inference kernels; IPDnet2 has no
backward kernels and raises in train mode.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

import ctypes as C

from . import _lib, ops
from .packing import pack_lstm_simt, pack_lstm_whh_t, unpack_lstm_simt_grad

Tensor = torch.Tensor


def mic_pairs(nmic: int, ch_mode: str) -> np.ndarray:
    """(P, 2) int32 microphone index pairs in the order of DPIPD.data_adjust (Module.py:500-514)."""
    if ch_mode == 'M':
        pairs = [(0, m) for m in range(1, nmic)]
    elif ch_mode == 'MM':
        pairs = [(i, j) for i in range(nmic - 1) for j in range(i + 1, nmic)]
    else:
        raise Exception('Microphone channel mode unrecognised')
    return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)


def non_source_target(mic_pos: np.ndarray, fre_use: Sequence[int] = range(1, 257), order: int = 0) -> np.ndarray:
    """IPDnet's target for a silent source (runIPDnetOn.py:209-222): [J0(2 pi f d_m / 340) | zeros] per non-reference mic,
    (2*len(fre_use), nmic-1) float64.  Host-side constant (computed once per array geometry)."""
    from scipy.special import jn
    d = np.sqrt(np.sum((mic_pos[1:] - mic_pos[0, :]) ** 2, axis=1))
    fr = (2 * np.pi * np.linspace(0, 8000, 257) / 340)[list(fre_use)]
    cols = [np.concatenate((jn(order, fr * dist), np.zeros(256))) for dist in d]
    return np.array(cols).T


@ops.on_tensor_device
def dpipd_targets(source_doa: Tensor, mic_location, vad: Optional[Tensor] = None, ch_mode: str = 'MM', nf: int = 257,
                  fre_max: float = 8000.0, speed: float = 340.0, fre_range_used: range = range(1, 257),
                  vad_threshold: float = 0.0, per_source: bool = False, non_source: Optional[Tensor] = None) -> Tensor:
    """source_doa (nb, nt, 2, ns) [elevation, azimuth] radians on the device -> DP-IPD targets
        per_source=False (FN-SSL, main.py:259): (nb, nt, 2*nbins, P)     = sum_s gate_s * [cos | sin](2 pi f ITD_s)
        per_source=True  (IPDnet):              (nb, nt, 2*nbins, P, ns) with silent sources replaced by `non_source` (2*nbins, P)
    gate_s = 1 if vad[b,t,s] > vad_threshold else 0 (no gating when vad is None: tar_useVAD = False)."""
    ops._need_cuda(source_doa, vad, non_source)
    if source_doa.dim() != 4 or source_doa.shape[2] != 2:
        raise RuntimeError("dpipd_targets: source_doa must be (nb, nt, 2, nsource)")
    lib = _lib.load()
    dev = source_doa.device
    nb, nt, _, ns = source_doa.shape
    mic = np.asarray(mic_location, dtype=np.float64).reshape(-1, 3)
    pairs = mic_pairs(mic.shape[0], ch_mode)
    bins = list(fre_range_used)
    if bins != list(range(bins[0], bins[0] + len(bins))):
        raise RuntimeError("dpipd_targets: fre_range_used must be a contiguous range")
    P, nbins = pairs.shape[0], len(bins)
    mic_d = torch.as_tensor(mic, dtype=torch.float32, device=dev).contiguous()
    pairs_d = torch.as_tensor(pairs, device=dev).contiguous()
    doa = source_doa.detach().float().contiguous()
    vad_d = vad.detach().float().contiguous() if vad is not None else None
    if vad_d is not None and tuple(vad_d.shape) != (nb, nt, ns):
        raise RuntimeError(f"dpipd_targets: vad must be (nb, nt, nsource) = {(nb, nt, ns)}")
    ns_d = None
    if non_source is not None:
        ns_d = non_source.detach().float().contiguous()
        if tuple(ns_d.shape) != (2 * nbins, P):
            raise RuntimeError(f"dpipd_targets: non_source must be (2*nbins, P) = {(2 * nbins, P)}")
    shape = (nb, nt, 2 * nbins, P, ns) if per_source else (nb, nt, 2 * nbins, P)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    ops._count(1)
    _lib.check(lib.fnssl_dpipd_targets(doa.data_ptr(), ops._ptr(vad_d), mic_d.data_ptr(), pairs_d.data_ptr(), nb, nt, ns, mic.shape[0],
                                       P, nf, float(fre_max), float(speed), bins[0], nbins, float(vad_threshold), int(per_source),
                                       ops._ptr(ns_d), out.data_ptr(), ops._stream()))
    return out


class _MseLoss(torch.autograd.Function):
    """Scalar loss from the CUDA kernel; gradient w.r.t. the prediction is analytic (2 (pred - gt) / N), so a head trained on
    top of frozen hot-path features can back-propagate through it."""

    @staticmethod
    def forward(ctx, pred, gt, nb):
        lib = _lib.load()
        P = pred.shape[0] // nb
        _, nt, nf2 = pred.shape
        p, g = pred.detach().float().contiguous(), gt.detach().float().contiguous()
        ws = torch.empty(nb * nt, dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        ops._count(2)
        _lib.check(lib.fnssl_ipd_mse_loss(p.data_ptr(), g.data_ptr(), nb, P, nt, nf2, ws.data_ptr(), loss.data_ptr(), ops._stream()))
        ctx.save_for_backward(p, g)
        ctx.nb = nb
        return loss

    @staticmethod
    def backward(ctx, grad):
        p, g = ctx.saved_tensors
        nb = ctx.nb
        P = p.shape[0] // nb
        gt_as_pred = g.permute(0, 3, 1, 2).reshape(nb * P, p.shape[1], p.shape[2])      # inverse of removebatch + permute
        return grad * 2.0 * (p - gt_as_pred) / p.numel(), None, None


@ops.on_tensor_device
def ipd_mse_loss(pred_ipd: Tensor, gt_ipd: Tensor) -> Tensor:
    """cal_loss of FN-SSL (main.py:191-198): pred (nb*P, nt, 2nf) is re-batched to (nb, nt, 2nf, P) and compared with the
    target (nb, nt, 2nf, P); mean squared error, a 0-d tensor."""
    ops._need_cuda(pred_ipd, gt_ipd)
    nb, nt, nf2, P = gt_ipd.shape
    if pred_ipd.shape != (nb * P, nt, nf2):
        raise RuntimeError(f"ipd_mse_loss: pred must be (nb*P, nt, 2nf) = {(nb * P, nt, nf2)}, got {tuple(pred_ipd.shape)}")
    return _MseLoss.apply(pred_ipd, gt_ipd, nb)


class _PitLoss(torch.autograd.Function):
    """Loss and permutation from the CUDA kernel; the gradient w.r.t. the prediction is analytic: with the per-frame permutation
    fixed (as torchmetrics' pit_permutate does, runIPDnetOn.py:203-205) it is 2 (pred_perm - gt) / N routed back through it."""

    @staticmethod
    def forward(ctx, pred, gt):
        nb, nt, _, _, ns = pred.shape
        rows = nb * nt
        p = pred.detach().float().reshape(rows, -1, ns).contiguous()
        g = gt.detach().float().reshape(rows, -1, ns).contiguous()
        if p.shape != g.shape:
            raise RuntimeError(f"ipd_pit_mse_loss: prediction {tuple(p.shape)} and target {tuple(g.shape)} differ")
        lib = _lib.load()
        ws = torch.empty(rows, dtype=torch.float32, device=p.device)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        perm = torch.empty((rows, ns), dtype=torch.int32, device=p.device)
        ops._count(2)
        _lib.check(lib.fnssl_ipd_pit_mse_loss(p.data_ptr(), g.data_ptr(), rows, p.shape[1], ns, ws.data_ptr(), loss.data_ptr(),
                                              perm.data_ptr(), ops._stream()))
        ctx.save_for_backward(p, g, perm)
        ctx.shape = tuple(pred.shape)
        ctx.mark_non_differentiable(perm)
        return loss, perm

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad, _grad_perm):
        p, g, perm = ctx.saved_tensors
        idx = perm.long().unsqueeze(1).expand_as(p)                     # pred_perm[r, :, s] = p[r, :, perm[r, s]]
        d = (torch.gather(p, 2, idx) - g) * (2.0 / p.numel())
        gp = torch.zeros_like(p).scatter_(2, idx, d)
        return (grad * gp).reshape(ctx.shape), None


@ops.on_tensor_device
def ipd_pit_mse_loss(pred_batch: Tensor, ipd_gt_batch: Tensor) -> Tuple[Tensor, Tensor]:
    """Frame-level PIT loss of IPDnet (runIPDnetOn.py:196-206): pred (nb, nt, 2nf, nmic-1, ns), target (nb*nt, 2nf, nmic-1, ns)
    [or the same 5-D shape]; per frame the source permutation with the smallest MSE is applied to the prediction, the loss is
    the MSE after permuting.  Returns (loss 0-d, best_perm (nb*nt, ns) int32 with pred index per target source); the loss
    back-propagates to the prediction."""
    ops._need_cuda(pred_batch, ipd_gt_batch)
    return _PitLoss.apply(pred_batch, ipd_gt_batch)


# ---------------------------------------------------------------------------------------------
# LSTM layer and DP-IPD head with their backward passes (fp32 engine)
# ---------------------------------------------------------------------------------------------

def _lstm_args(axis, src0, c0, src1, c1, w, hidden, num_dirs, out0):
    nb, nt, nf, ld0 = src0.shape
    a = _lib.LstmArgs()
    a.engine, a.axis = ops.ENGINE_SIMT, axis
    a.nb, a.nt, a.nf = nb, nt, nf
    a.hidden, a.num_dirs, a.dtype = hidden, num_dirs, ops.F32
    a.src0, a.c0, a.ld0 = src0.data_ptr(), c0, ld0
    a.src1, a.c1, a.ld1 = (src1.data_ptr(), c1, src1.shape[-1]) if src1 is not None else (None, 0, 0)
    a.weights, a.weights_bytes = w.data_ptr(), w.numel() * w.element_size()
    a.out0, a.out0_ld, a.out0_off = out0.data_ptr(), out0.shape[-1], 0
    return a


class _LstmLayer(torch.autograd.Function):
    """h = LSTM(concat(src0[..., :c0], src1[..., :c1])) over a grid; params = nn.LSTM's (weight_ih, weight_hh, bias_ih, bias_hh)
    per direction.  Backward = fnssl_lstm_backward (BPTT kernel + two reduction products), once per forward."""

    @staticmethod
    def forward(ctx, src0, src1, axis, c0, c1, hidden, num_dirs, *params):
        lib = _lib.load()
        dirs = [tuple(p.detach().float() for p in params[4 * d:4 * d + 4]) for d in range(num_dirs)]
        w = pack_lstm_simt(dirs)
        nb, nt, nf, _ = src0.shape
        out = torch.empty((nb, nt, nf, hidden * num_dirs), dtype=torch.float32, device=src0.device)
        nbytes = int(lib.fnssl_lstm_train_saved_bytes(nb, nt, nf, hidden, num_dirs))
        saved = torch.empty(nbytes // 4, dtype=torch.float32, device=src0.device)
        a = _lstm_args(axis, src0, c0, src1, c1, w, hidden, num_dirs, out)
        ops._count(1)
        _lib.check(lib.fnssl_lstm_forward_train(C.byref(a), saved.data_ptr(), nbytes, ops._stream()))
        ctx.save_for_backward(src0, src1, w, out, saved, pack_lstm_whh_t(dirs))
        ctx.cfg = (axis, c0, c1, hidden, num_dirs)
        ctx.consumed = False
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        if ctx.consumed:
            raise RuntimeError("lstm_layer: the saved activations are consumed by the first backward pass (no retain_graph)")
        ctx.consumed = True
        src0, src1, w, out, saved, whh_t = ctx.saved_tensors
        axis, c0, c1, hidden, num_dirs = ctx.cfg
        lib = _lib.load()
        with torch.cuda.device(src0.device):
            dout = dout.contiguous().float()
            d0 = d1 = None
            if ctx.needs_input_grad[0]:
                d0 = torch.empty_like(src0) if src0.shape[-1] == c0 else torch.zeros_like(src0)
            if src1 is not None and ctx.needs_input_grad[1]:
                d1 = torch.empty_like(src1) if src1.shape[-1] == c1 else torch.zeros_like(src1)
            dw = torch.empty_like(w)
            a = _lstm_args(axis, src0, c0, src1, c1, w, hidden, num_dirs, out)
            ops._count(3)
            _lib.check(lib.fnssl_lstm_backward(C.byref(a), saved.data_ptr(), saved.numel() * 4, whh_t.data_ptr(), dout.data_ptr(),
                                               dout.shape[-1], ops._ptr(d0), d0.shape[-1] if d0 is not None else 0, ops._ptr(d1),
                                               d1.shape[-1] if d1 is not None else 0, dw.data_ptr(), ops._stream()))
        grads = []
        for d, g4 in enumerate(unpack_lstm_simt_grad(dw, num_dirs, c0 + c1, hidden)):
            grads += [g if ctx.needs_input_grad[7 + 4 * d + i] else None for i, g in enumerate(g4)]
        return (d0, d1, None, None, None, None, None) + tuple(grads)


@ops.on_tensor_device
def lstm_layer(src0: Tensor, c0: int, src1: Optional[Tensor], c1: int, params, axis: int) -> Tensor:
    """One differentiable LSTM layer over fp32 grids: src0 (nb, nt, nf, ld0 >= c0) [+ src1 (.., ld1 >= c1) concatenated],
    params = fn_ssl_b200.packing.LSTMParams (nn.LSTM's parameter set) -> h (nb, nt, nf, dirs*hidden).
    axis = ops.ALONG_FREQ (full-band: sequences over f) | ops.ALONG_TIME (narrow-band: sequences over t)."""
    ops._need_cuda(src0, src1)
    if src0.dtype != torch.float32 or (src1 is not None and src1.dtype != torch.float32):
        raise RuntimeError("lstm_layer: the training path runs on float32 grids")
    if params.hidden_size not in (32, 64, 128, 256):
        raise RuntimeError(f"lstm_layer: hidden size {params.hidden_size} not supported (32, 64, 128, 256)")
    if c0 + (c1 if src1 is not None else 0) != params.input_size:
        raise RuntimeError(f"lstm_layer: {c0} + {c1} input channels, the layer takes {params.input_size}")
    flat = [t for d in params.directions() for t in d]
    return _LstmLayer.apply(src0.contiguous(), src1.contiguous() if src1 is not None else None, axis, c0, c1 if src1 is not None else 0,
                            params.hidden_size, params.num_dirs, *flat)


class _IpdHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        y = ops.ipd_head(x, x.shape[-1], weight, bias)
        ctx.save_for_backward(x, weight.detach().float().contiguous(), y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        nb, nt, nf, Cc = x.shape
        with torch.cuda.device(x.device):
            dy = dy.contiguous().float()
            dx = torch.empty_like(x) if nt % 12 == 0 else torch.zeros_like(x)
            dw = torch.empty_like(w)
            db = torch.empty(2, dtype=torch.float32, device=x.device)
            ops._count(1)
            _lib.check(_lib.load().fnssl_ipd_head_backward(x.data_ptr(), Cc, nb, nt, nf, Cc, w.data_ptr(), y.data_ptr(), dy.data_ptr(),
                                                           dx.data_ptr(), Cc, dw.data_ptr(), db.data_ptr(), ops._stream()))
        return dx, dw, db


@ops.on_tensor_device
def ipd_head_train(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Differentiable FN-SSL head on an fp32 grid x (nb, nt, nf, C): AvgPool over 12 frames -> Linear(C, 2) -> tanh ->
    (nb, nt//12, 2*nf) = [channel 0 over f | channel 1 over f] (Model.py:79-87)."""
    ops._need_cuda(x, weight, bias)
    if x.dtype != torch.float32 or x.shape[-1] != weight.shape[1] or x.shape[-1] > 512:
        raise RuntimeError("ipd_head_train: x must be a float32 grid with C = weight.shape[1] <= 512 channels")
    return _IpdHead.apply(x.contiguous(), weight, bias)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        y = ops.linear(x, weight, bias)
        ctx.save_for_backward(x.detach().float().reshape(-1, x.shape[-1]).contiguous(), weight.detach().float().contiguous())
        ctx.xshape = tuple(x.shape)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        rows, inf = x2.shape
        outf = w.shape[0]
        with torch.cuda.device(x2.device):
            dy2 = dy.contiguous().float().reshape(rows, outf)
            dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
            dw, db = torch.empty_like(w), torch.empty(outf, dtype=torch.float32, device=w.device)
            ops._count(2)
            _lib.check(_lib.load().fnssl_linear_backward(x2.data_ptr(), w.data_ptr(), dy2.data_ptr(), rows, inf, outf, ops._ptr(dx),
                                                         dw.data_ptr(), db.data_ptr(), ops._stream()))
        return (dx.reshape(ctx.xshape) if dx is not None else None), dw, db


@ops.on_tensor_device
def linear_train(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Differentiable y = x @ weight^T + bias on the CUDA kernels of head.cu (the DOA classifier Linear(512,180), Model.py:71,88-89)."""
    ops._need_cuda(x, weight, bias)
    return _Linear.apply(x, weight, bias)


# ---------------------------------------------------------------------------------------------
# IPDnet's causal conv block with its backward pass
# ---------------------------------------------------------------------------------------------

class _Conv3x3(torch.autograd.Function):
    """y = causal conv3x3(concat(in0[..., :c0], in1[..., :c1])) on fp32 grids (nb, nt, nf, C): nn.Conv2d(kernel 3x3, padding (1,2),
    bias=False) + crop of the last two frames (FixedAarryIPDnet.py:50-52,62-64); weight (O, C, 3, 3)."""

    @staticmethod
    def forward(ctx, in0, in1, c0, c1, weight):
        lib = _lib.load()
        nb, nt, nf, ld0 = in0.shape
        w = weight.detach().float().contiguous()
        O = w.shape[0]
        out = torch.empty((nb, nt, nf, O), dtype=torch.float32, device=in0.device)
        work = torch.empty(9 * (c0 + c1) * O, dtype=torch.float32, device=in0.device)
        ops._count(2)
        _lib.check(lib.fnssl_conv3x3_forward(in0.data_ptr(), c0, ld0, ops._ptr(in1), c1, in1.shape[-1] if in1 is not None else 0,
                                             nb, nt, nf, w.data_ptr(), O, work.data_ptr(), out.data_ptr(), O, ops._stream()))
        ctx.save_for_backward(in0, in1, w)
        ctx.cfg = (c0, c1)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        in0, in1, w = ctx.saved_tensors
        c0, c1 = ctx.cfg
        lib = _lib.load()
        nb, nt, nf, ld0 = in0.shape
        O = w.shape[0]
        with torch.cuda.device(in0.device):
            dy = dy.contiguous().float()
            work = torch.empty(9 * (c0 + c1) * O, dtype=torch.float32, device=in0.device)
            d0 = d1 = dw = None
            if ctx.needs_input_grad[0]:
                d0 = torch.empty_like(in0) if ld0 == c0 else torch.zeros_like(in0)
            if in1 is not None and ctx.needs_input_grad[1]:
                d1 = torch.empty_like(in1) if in1.shape[-1] == c1 else torch.zeros_like(in1)
            if d0 is not None or d1 is not None:
                ops._count(2)
                _lib.check(lib.fnssl_conv3x3_backward_data(dy.data_ptr(), O, O, nb, nt, nf, w.data_ptr(), c0, c1, work.data_ptr(),
                                                           ops._ptr(d0), d0.shape[-1] if d0 is not None else 0, ops._ptr(d1),
                                                           d1.shape[-1] if d1 is not None else 0, ops._stream()))
            if ctx.needs_input_grad[4]:
                dw = torch.empty_like(w)
                ops._count(2)
                _lib.check(lib.fnssl_conv3x3_backward_weight(in0.data_ptr(), c0, ld0, ops._ptr(in1), c1,
                                                             in1.shape[-1] if in1 is not None else 0, dy.data_ptr(), O, O, nb, nt, nf,
                                                             work.data_ptr(), dw.data_ptr(), ops._stream()))
        return d0, d1, None, None, dw


@ops.on_tensor_device
def conv3x3_causal(in0: Tensor, c0: int, in1: Optional[Tensor], c1: int, weight: Tensor) -> Tensor:
    """Differentiable causal 3x3 conv over fp32 grids: (nb, nt, nf, ld0 >= c0) [+ (.., ld1 >= c1) concatenated] -> (nb, nt, nf, O)."""
    ops._need_cuda(in0, in1, weight)
    if in0.dtype != torch.float32 or (in1 is not None and in1.dtype != torch.float32):
        raise RuntimeError("conv3x3_causal: the training path runs on float32 grids")
    if tuple(weight.shape[1:]) != (c0 + (c1 if in1 is not None else 0), 3, 3):
        raise RuntimeError("conv3x3_causal: weight shape does not match the input channels / 3x3 kernel")
    return _Conv3x3.apply(in0.contiguous(), in1.contiguous() if in1 is not None else None, c0, c1 if in1 is not None else 0, weight)


def causcnn_train(src0: Tensor, c0: int, src1: Optional[Tensor], c1: int, w1: Tensor, w2: Tensor, w3: Tensor) -> Tensor:
    """CausCnnBlock.forward (FixedAarryIPDnet.py:61-73) with a backward pass: grids in, (nb, cout, nf, nt//12) out (the reference's
    layout).  The three convs are the CUDA products of conv_train.cu; ReLU, the two average poolings over 3 and 4 frames and tanh
    are torch elementwise / view ops on the grids."""
    nb, nt, nf, _ = src0.shape
    a = torch.relu(conv3x3_causal(src0, c0, src1, c1, w1))
    nt1 = nt // 3
    a = a[:, :nt1 * 3].reshape(nb, nt1, 3, nf, a.shape[-1]).mean(dim=2)          # AvgPool2d((1, 3)) over time
    a = torch.relu(conv3x3_causal(a, a.shape[-1], None, 0, w2))
    nt2 = nt1 // 4
    a = a[:, :nt2 * 4].reshape(nb, nt2, 4, nf, a.shape[-1]).mean(dim=2)          # AvgPool2d((1, 4))
    y = torch.tanh(conv3x3_causal(a, a.shape[-1], None, 0, w3))
    return y.permute(0, 3, 2, 1)


# ---------------------------------------------------------------------------------------------
# The reference's LightningModule surface for training, without Lightning
# ---------------------------------------------------------------------------------------------

class FNSSLTrainModule(torch.nn.Module):
    """Plain-PyTorch counterpart of the training-relevant surface of the reference's LightningModule `MyModel`
    (FN-SSL/Lightning/main.py:80-279; pytorch_lightning is not a dependency of this package): `forward`, `data_preprocess`
    (features AND the DP-IPD targets, both on the device), `cal_loss`, `training_step`, `validation_step`, `predict_step`,
    `configure_optimizers` -- same names, argument meaning and return values, so a Lightning subclass only has to forward to them.

        batch = (mic_sig_batch (nb, nsample, nch) f32,
                 {'doa': (nb, nseg, 2, nsource) [elevation, azimuth] rad, 'vad_sources': (nb, nseg, nvad, nsource)})
        step  = module.training_step(batch, 0)["loss"];  step.backward();  distributed.all_reduce_gradients(module.arch);  opt.step()
    """

    def __init__(self, tar_useVAD: bool = True, ch_mode: str = 'MM', fs: int = 16000, win_len: int = 512, nfft: int = 512,
                 win_shift_ratio: float = 0.5, mic_location=((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0)), speed: float = 340.0,
                 arch: Optional[torch.nn.Module] = None):
        super().__init__()
        if (win_len, nfft, win_shift_ratio) != (512, 512, 0.5):
            raise RuntimeError("FNSSLTrainModule: the front end is built for win_len = nfft = 512, hop 256")
        if arch is None:
            from .Model import FN_SSL
            arch = FN_SSL()                                                # main.py:100
        self.arch = arch
        self.tar_useVAD, self.ch_mode, self.nfft, self.fre_max, self.speed = tar_useVAD, ch_mode, nfft, fs / 2, speed
        self.mic_location = np.asarray(mic_location, dtype=np.float64)
        self.fre_range_used = range(1, nfft // 2 + 1, 1)                   # main.py:128

    def _device(self) -> torch.device:
        return next(self.arch.parameters()).device

    def forward(self, x: Tensor) -> Tensor:
        return self.arch(x)

    def data_preprocess(self, mic_sig_batch: Optional[Tensor] = None, gt_batch: Optional[dict] = None, eps: float = 1e-6,
                        nor_flag: bool = True) -> list:
        """main.py:200-266: [network input (nb*P, 4, 256, nt)] + [gt dict with 'ipd' (nb, nseg, 2*256, P) added] -- STFT, pair
        re-batching, forgetting_norm and the feature assembly are the CUDA front end; the targets come from `dpipd_targets`."""
        from .pipeline import data_preprocess_fnssl
        dev = self._device()
        data = []
        if mic_sig_batch is not None:
            data += data_preprocess_fnssl(mic_sig_batch.to(dev), ch_mode=self.ch_mode, eps=eps, nor_flag=nor_flag)
        if gt_batch is not None:
            doa = gt_batch['doa'].to(dev).float()
            vad = gt_batch['vad_sources'].to(dev).float().mean(dim=2)      # (nb, nseg, nsource), main.py:241
            ipd = dpipd_targets(doa, self.mic_location, vad=vad if self.tar_useVAD else None, ch_mode=self.ch_mode,
                                nf=self.nfft // 2 + 1, fre_max=self.fre_max, speed=self.speed, fre_range_used=self.fre_range_used)
            gt_batch = dict(gt_batch)
            gt_batch['doa'], gt_batch['ipd'], gt_batch['vad_sources'] = doa, ipd, vad
            data += [gt_batch]
        return data

    def cal_loss(self, pred_batch: Tensor, gt_batch: dict) -> Tensor:
        return ipd_mse_loss(pred_batch, gt_batch['ipd'])                   # main.py:191-198

    def _step(self, batch) -> Tensor:
        in_batch, gt_batch = self.data_preprocess(batch[0], batch[1])
        return self.cal_loss(self(in_batch), gt_batch)

    def training_step(self, batch, batch_idx: int = 0) -> dict:
        return {"loss": self._step(batch)}                                 # main.py:95-103

    def validation_step(self, batch, batch_idx: int = 0) -> Tensor:
        with torch.no_grad():
            return self._step(batch)                                       # main.py:105-115 without the DOA metrics

    def predict_step(self, batch: Tensor, batch_idx: int = 0) -> Tensor:
        return self(self.data_preprocess(mic_sig_batch=batch.permute(0, 2, 1))[0])     # main.py:183-189

    def configure_optimizers(self) -> dict:
        optimizer = torch.optim.Adam(self.arch.parameters(), lr=0.001)     # main.py:268-279
        lr_scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.8988, last_epoch=-1)
        return {'optimizer': optimizer, 'lr_scheduler': {'scheduler': lr_scheduler, 'monitor': 'valid/loss'}}


class IPDnetTrainModule(torch.nn.Module):
    """The same for IPDnet: the training-relevant surface of `MyModel` in IPDnet/runIPDnetOn.py:80-304 -- `data_preprocess` returns
    [network input (nb, 2M, 256, nt), doa, per-source DP-IPD targets (nb*nt2, 512, M-1, nsrc) with silent sources replaced by the
    non-source (Bessel) target, dp_vad], `cal_loss` is the frame-level PIT loss, `training_step(batch)` -> {"loss": ...}.

        batch = (mic_sig_batch (nb, nsample, M) f32,
                 {'doa': (nb, nt2, 2, nsrc) rad, 'dp_signal': (nb, nsample, M, nsrc) direct-path signals for the VAD})"""

    def __init__(self, tar_useVAD: bool = True, ch_mode: str = 'M', fs: int = 16000, win_len: int = 512, nfft: int = 512,
                 win_shift_ratio: float = 0.5, max_source: int = 2, mic_pos=((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0)),
                 arch: Optional[torch.nn.Module] = None):
        super().__init__()
        if (win_len, nfft, win_shift_ratio) != (512, 512, 0.5):
            raise RuntimeError("IPDnetTrainModule: the front end is built for win_len = nfft = 512, hop 256")
        if arch is None:
            from .FixedAarryIPDnet import IPDnet
            arch = IPDnet()                                                # runIPDnetOn.py:101 (2-mic IPDnet)
        self.arch = arch
        self.tar_useVAD, self.ch_mode, self.nfft, self.fre_max, self.max_source = tar_useVAD, ch_mode, nfft, fs / 2, max_source
        self.mic_pos = np.asarray(mic_pos.cpu().numpy() if torch.is_tensor(mic_pos) else mic_pos, dtype=np.float64)
        self.fre_range_used = range(1, nfft // 2 + 1, 1)
        self.register_buffer("non_source_tar", torch.from_numpy(non_source_target(self.mic_pos, self.fre_range_used)).float(),
                             persistent=False)                             # euclidean_distances_to_bessel, :209-222

    def _device(self) -> torch.device:
        return next(self.arch.parameters()).device

    def forward(self, x: Tensor) -> Tensor:
        return self.arch(x)

    def cal_vad(self, dp_mic_sig_batch: Tensor, spec: Tensor) -> Tensor:
        """runIPDnetOn.py:224-235: per source, mean over the 257 bins of |STFT(direct path)| / |STFT(mixture)| on microphone 0,
        averaged over the 12 frames of an output frame -> (nb, nt // 12, max_source).  The STFTs are the CUDA kernel."""
        nb, nf, nt, _ = spec.shape
        mix = spec[:, :, :, 0].abs()
        vad = torch.stack([(ops.stft(dp_mic_sig_batch[:, :, :, s].contiguous())[0][:, :, :, 0].abs() / mix).mean(dim=1)
                           for s in range(self.max_source)], dim=2)         # (nb, nt, nsrc)
        nt2 = nt // 12
        return vad[:, :nt2 * 12].reshape(nb, nt2, 12, self.max_source).mean(dim=2)

    def data_preprocess(self, mic_sig_batch: Tensor, acoustic_scene_batch: Optional[dict] = None, eps: float = 1e-6) -> list:
        dev = self._device()
        sig = mic_sig_batch.to(dev)
        spec, magsum = ops.stft(sig, want_magsum=True)
        _, _, cf = ops.features(spec, magsum, 'ALL', ops.NORM_FORGETTING, 280, eps, torch.float32, want_cfirst=True)     # :240-254
        data = [cf]
        if acoustic_scene_batch is None:
            return data
        dp_vad = self.cal_vad(acoustic_scene_batch['dp_signal'].to(dev).float(), spec)
        doa = acoustic_scene_batch['doa'].to(dev).float()
        ipd = dpipd_targets(doa, self.mic_pos, vad=dp_vad, ch_mode=self.ch_mode, nf=self.nfft // 2 + 1, fre_max=self.fre_max,
                            speed=340.0, fre_range_used=self.fre_range_used, vad_threshold=0.001, per_source=True,
                            non_source=self.non_source_tar.to(dev))        # :256-283
        nb, nt2 = ipd.shape[0], ipd.shape[1]
        data += [doa, ipd.reshape(nb * nt2, ipd.shape[2], ipd.shape[3], ipd.shape[4])]
        if self.tar_useVAD:
            data += [dp_vad]
        return data

    def cal_loss(self, pred_batch: Tensor, gt_batch: list) -> Tensor:
        return ipd_pit_mse_loss(pred_batch, gt_batch[1])[0]                # :196-206

    def _step(self, batch) -> Tensor:
        data = self.data_preprocess(batch[0], batch[1])
        return self.cal_loss(self(data[0]), data[1:])

    def training_step(self, batch, batch_idx: int = 0) -> dict:
        return {"loss": self._step(batch)}                                 # :144-154

    def validation_step(self, batch, batch_idx: int = 0) -> Tensor:
        with torch.no_grad():
            return self._step(batch)

    def predict_step(self, batch: Tensor, batch_idx: int = 0) -> Tensor:
        return self(self.data_preprocess(mic_sig_batch=batch.permute(0, 2, 1))[0])[0]      # :182-186 (first utterance, as there)

    def configure_optimizers(self) -> dict:
        optimizer = torch.optim.Adam(self.arch.parameters(), lr=0.0005)    # :293-304
        lr_scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=0.975, last_epoch=-1)
        return {'optimizer': optimizer, 'lr_scheduler': {'scheduler': lr_scheduler, 'monitor': 'valid/loss'}}
