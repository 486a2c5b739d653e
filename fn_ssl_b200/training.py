"""Training-side forward pieces on the GPU: DP-IPD regression targets and the losses.

    dpipd_targets(...)        DPIPD.forward(source_doa) + the ground-truth branch of data_preprocess
                              (FN-SSL/Lightning/Module.py:464-497, main.py:227-265; IPDnet/runIPDnetOn.py:256-290)
    ipd_mse_loss(...)         cal_loss, FN-SSL/Lightning/main.py:191-198
    ipd_pit_mse_loss(...)     frame-level PIT loss, IPDnet/runIPDnetOn.py:188-206

Scope: these are the FORWARD computations of the training step that sit next to the hot path (targets are built per batch on
the host in the reference: float64 numpy loops + a host->device copy).  The backward pass of the fused LSTM kernels is not
implemented -- `FN_SSL` / `IPDnet` still raise in train mode -- so the losses are exposed as plain functions (with an analytic
gradient w.r.t. the prediction for callers that train a head on frozen features), not as a Lightning training_step.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

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


@ops.on_tensor_device
def ipd_pit_mse_loss(pred_batch: Tensor, ipd_gt_batch: Tensor) -> Tuple[Tensor, Tensor]:
    """Frame-level PIT loss of IPDnet (runIPDnetOn.py:196-206): pred (nb, nt, 2nf, nmic-1, ns), target (nb*nt, 2nf, nmic-1, ns)
    [or the same 5-D shape]; per frame the source permutation with the smallest MSE is applied to the prediction, the loss is
    the MSE after permuting.  Returns (loss 0-d, best_perm (nb*nt, ns) int32 with pred index per target source)."""
    ops._need_cuda(pred_batch, ipd_gt_batch)
    nb, nt, _, _, ns = pred_batch.shape
    rows = nb * nt
    p = pred_batch.detach().float().reshape(rows, -1, ns).contiguous()
    g = ipd_gt_batch.detach().float().reshape(rows, -1, ns).contiguous()
    if p.shape != g.shape:
        raise RuntimeError(f"ipd_pit_mse_loss: prediction {tuple(p.shape)} and target {tuple(g.shape)} differ")
    lib = _lib.load()
    ws = torch.empty(rows, dtype=torch.float32, device=p.device)
    loss = torch.empty((), dtype=torch.float32, device=p.device)
    perm = torch.empty((rows, ns), dtype=torch.int32, device=p.device)
    ops._count(2)
    _lib.check(lib.fnssl_ipd_pit_mse_loss(p.data_ptr(), g.data_ptr(), rows, p.shape[1], ns, ws.data_ptr(), loss.data_ptr(),
                                          perm.data_ptr(), ops._stream()))
    return loss, perm
