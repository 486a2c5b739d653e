"""CPU oracle for the training-side forward pieces.  TEST INFRASTRUCTURE (see fnssl_oracle.py).

Parity status: PINNED for the targets and the FN-SSL loss -- tests/golden/make_golden.py runs the reference's own ``DPIPD`` /
``RemoveChFromBatch`` modules plus the restated lines of ``data_preprocess`` / ``cal_loss`` (the Lightning wrappers cannot be
imported: pytorch_lightning is absent) and stores the results in tests/golden/train_golden.npz.  The frame-level PIT search
comes from ``torchmetrics.functional.audio.permutation_invariant_training`` (third party, not vendored, not installed;
IPDnet/runIPDnetOn.py:19,203): restated here as the exhaustive search over ``itertools.permutations`` it performs for <= 3
speakers -- the loss value does not depend on tie-breaking, the returned permutation does ("parity unpinned" for ties only).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from . import fnssl_oracle as orc


def dpipd_of_doa(source_doa: np.ndarray, mic_location: np.ndarray, nf: int = 257, fre_max: float = 8000.0, ch_mode: str = "MM",
                 speed: float = 340.0) -> np.ndarray:
    """``DPIPD.forward(source_doa)`` FN-SSL/Lightning/Module.py:464-497: (nb, nt, 2, ns) -> complex (nb, nt, nf, P, ns),
    phase = +2 pi f r.(mic_m1 - mic_m2)/c  (the reference's "-2 pi f ITD * (-1)")."""
    sd = source_doa.transpose(0, 1, 3, 2)                                               # (nb, nt, ns, 2)
    nmic = mic_location.shape[-2]
    r = np.stack([np.sin(sd[..., 0]) * np.cos(sd[..., 1]), np.sin(sd[..., 0]) * np.sin(sd[..., 1]), np.cos(sd[..., 0])], axis=3)
    fre = np.linspace(0.0, fre_max, nf)
    ipd = np.empty(sd.shape[:3] + (nf, nmic, nmic))
    for m1 in range(nmic):
        for m2 in range(nmic):
            itd = np.dot(r, mic_location[m1, :] - mic_location[m2, :]) / speed          # :485
            ipd[..., m1, m2] = 2 * np.pi * fre[None, None, None, :] * itd[..., None]    # :486-487
    return orc._pair_select(np.exp(1j * ipd), ch_mode).transpose(0, 1, 3, 4, 2)         # :490-492


def fnssl_targets(source_doa: np.ndarray, vad: np.ndarray, mic_location: np.ndarray, ch_mode: str = "MM", tar_use_vad: bool = True,
                  fre_range_used=range(1, 257), **kw) -> torch.Tensor:
    """Ground-truth branch of ``data_preprocess`` FN-SSL/Lightning/main.py:227-265 -> (nb, nt, 2*nbins, P) float32.
    ``vad`` is the already time-averaged (nb, nt, ns) activity (:244)."""
    d = dpipd_of_doa(source_doa, mic_location, ch_mode=ch_mode, **kw)
    bins = list(fre_range_used)
    ipd = torch.from_numpy(np.concatenate((d.real[:, :, bins], d.imag[:, :, bins]), axis=2).astype(np.float32))   # :240-242
    if tar_use_vad:
        gate = torch.from_numpy(np.asarray(vad, dtype=np.float32)).clone()
        gate[gate <= 0] = 0                                                                                          # :253-255
        gate[gate > 0] = 1
        ipd = ipd * gate[:, :, None, None, :]
    return ipd.sum(dim=-1)                                                                                           # :259


def ipdnet_targets(source_doa: np.ndarray, vad: np.ndarray, mic_location: np.ndarray, non_source: np.ndarray, ch_mode: str = "M",
                   fre_range_used=range(1, 257), th: float = 0.001, **kw) -> torch.Tensor:
    """IPDnet/runIPDnetOn.py:256-283 -> (nb, nt, 2*nbins, P, ns): VAD-gated per-source targets, silent sources replaced by the
    non-source target."""
    d = dpipd_of_doa(source_doa, mic_location, ch_mode=ch_mode, **kw)
    bins = list(fre_range_used)
    ipd = torch.from_numpy(np.concatenate((d.real[:, :, bins], d.imag[:, :, bins]), axis=2).astype(np.float32))
    gate = torch.from_numpy(np.asarray(vad, dtype=np.float32)).clone()
    gate[gate <= th] = 0
    gate[gate > th] = 1
    ipd = ipd * gate[:, :, None, None, :]
    ns_t = torch.from_numpy(np.asarray(non_source)).to(ipd)
    nb, nt, _, _, ns = ipd.shape
    for i in range(nb):
        for j in range(nt):
            for k in range(ns):
                if (ipd[i, j, :, :, k] == 0).all():
                    ipd[i, j, :, :, k] = ns_t
    return ipd


def fnssl_loss(pred: torch.Tensor, gt_ipd: torch.Tensor) -> torch.Tensor:
    """``cal_loss`` FN-SSL/Lightning/main.py:191-198; RemoveChFromBatch = (nb*P, nt, 2nf) -> (nb, P, nt, 2nf) (Module.py:407-423)."""
    nb = gt_ipd.shape[0]
    P = pred.shape[0] // nb
    reb = pred.reshape((nb, P) + tuple(pred.shape[1:])).permute(0, 2, 3, 1)
    return torch.nn.functional.mse_loss(reb.contiguous(), gt_ipd.contiguous())


def ipdnet_pit_loss(pred: torch.Tensor, gt: torch.Tensor):
    """``cal_loss`` IPDnet/runIPDnetOn.py:196-206 with torchmetrics' exhaustive PIT restated.  pred (nb, nt, 2nf, P, ns),
    gt reshapeable to (nb*nt, 2nf*P, ns).  Returns (loss, best_perm (rows, ns): prediction index per target source)."""
    nb, nt, _, _, ns = pred.shape
    p = pred.reshape(nb * nt, -1, ns).permute(0, 2, 1)                                   # (rows, ns, K)
    g = gt.reshape(nb * nt, -1, ns).permute(0, 2, 1)
    # metric matrix [row, target j, prediction i] = mean_k (p[i] - g[j])^2 (MSE_loss, :188-192)
    mtx = ((p[:, None, :, :] - g[:, :, None, :]) ** 2).mean(-1)
    perms = list(itertools.permutations(range(ns)))
    cost = torch.stack([sum(mtx[:, j, pm[j]] for j in range(ns)) / ns for pm in perms], dim=1)    # (rows, nperm)
    best = cost.argmin(dim=1)
    best_perm = torch.tensor(perms)[best]                                                # (rows, ns)
    pp = torch.stack([torch.index_select(pr, 0, bp) for pr, bp in zip(p, best_perm)])   # pit_permutate
    return torch.nn.functional.mse_loss(pp.contiguous(), g.contiguous()), best_perm.to(torch.int32)
