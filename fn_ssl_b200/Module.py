"""Front-end layers: drop-in for the hot-path classes of FN-SSL/Module.py, FN-SSL/Lightning/Module.py and
IPDnet/Module.py (same names / constructor arguments / tensor layouts), executed by libfnssl_b200.so.

    STFT                (reference FN-SSL/Lightning/Module.py:28-68; IPDnet/Module.py:25-63)
    AddChToBatch        (:376-405)      RemoveChFromBatch (:407-421)
    forgetting_norm     (FN-SSL/Lightning/utils_.py:9-55)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops

Tensor = torch.Tensor


class STFT(nn.Module):
    """signal (nbatch, nsample, nch) -> STFT coefficients (nbatch, nf, nt, nch) complex64.
    Periodic Hann window, center=False, un-normalised, one-sided; nt = floor((nsample-win_len)/shift + 1)."""

    def __init__(self, win_len, win_shift_ratio, nfft, win='hann'):
        super().__init__()
        self.win_len = win_len
        self.win_shift_ratio = win_shift_ratio
        self.nfft = nfft
        self.win = win

    def forward(self, signal: Tensor) -> Tensor:
        if self.win != 'hann':
            raise Exception("fn_ssl_b200.STFT: only win='hann' is implemented (the only window any caller uses)")
        spec, _ = ops.stft(signal, self.win_len, int(self.win_len * self.win_shift_ratio), self.nfft)
        return spec


def _pairs(nch: int, ch_mode: str):
    if ch_mode == 'M':
        return [(0, m) for m in range(1, nch)]
    if ch_mode == 'MM':
        return [(i, j) for i in range(nch - 1) for j in range(i + 1, nch)]
    raise Exception('ch_mode unrecognised')


class AddChToBatch(nn.Module):
    """(nb, nch, ...) -> (nb*P, 2, ...): 'M' pairs (0, m); 'MM' all i<j pairs.  Pure indexing (one gather);
    the fused pipeline never materialises this tensor -- the pairing happens inside the feature kernel."""

    def __init__(self, ch_mode):
        super().__init__()
        self.ch_mode = ch_mode

    def forward(self, data: Tensor) -> Tensor:
        nb, nch = data.shape[:2]
        idx = torch.tensor(_pairs(nch, self.ch_mode), device=data.device)          # (P, 2)
        out = data[:, idx]                                                         # (nb, P, 2, ...)
        return out.reshape((nb * idx.shape[0], 2) + tuple(data.shape[2:])).contiguous()


class RemoveChFromBatch(nn.Module):
    """(nb*nmic, nt, nf) -> (nb, nmic, nt, nf)."""

    def __init__(self, ch_mode):
        super().__init__()
        self.ch_mode = ch_mode

    def forward(self, data: Tensor, nb: int) -> Tensor:
        nmic = int(data.shape[0] / nb)
        return data.reshape((nb, nmic) + tuple(data.shape[1:])).float().contiguous()


def forgetting_norm(input: Tensor, sample_length: int = 298) -> Tensor:
    """input [B, C, F, T] magnitudes -> [B, 1, 1, T] recursive frame-mean normaliser (utils_.py:9-55).
    The per-frame sums over F are a torch reduction (plumbing); the T-sequential recursion runs in
    fnssl_norm_forward.  The fused pipeline gets the sums from the STFT kernel instead."""
    assert input.ndim == 4
    ops._need_cuda(input)
    B, Cc, Fq, T = input.shape
    magsum = input.float().sum(dim=2).contiguous()                                  # (B, C, T)
    mu = torch.empty((B, T), dtype=torch.float32, device=input.device)
    _lib.check(_lib.load().fnssl_norm_forward(magsum.data_ptr(), B, Cc, T, Fq, _lib.PAIRS_ALL, _lib.NORM_FORGETTING,
                                              sample_length, mu.data_ptr(), ops._stream()))
    return mu.reshape(B, 1, 1, T)
