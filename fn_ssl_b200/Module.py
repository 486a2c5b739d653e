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


@ops.on_tensor_device
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


# ---------------------------------------------------------------------------------------------
# IPD -> DOA decoding ("next" row of the scope contract): DPIPD templates, SourceDetectLocalize, PredDOA.pred2DOA
# ---------------------------------------------------------------------------------------------

import numpy as np  # noqa: E402


class DPIPD(nn.Module):
    """Complex direct-path IPD templates of a far-field source for every (elevation, azimuth) candidate and the DP-IPD
    of given source DOAs -- host-side geometry in numpy, as in the reference (FN-SSL/Lightning/Module.py:424-514).
    forward(source_doa=None) -> (dpipd_template (nele,nazi,nf,npairs), dpipd | None, [ele_candidate, azi_candidate])."""

    def __init__(self, ndoa_candidate, mic_location, nf=257, fre_max=8000, ch_mode='M', speed=343.0):
        super().__init__()
        self.ndoa_candidate = ndoa_candidate
        self.mic_location = np.asarray(mic_location, dtype=np.float64)
        self.nf, self.fre_max, self.speed, self.ch_mode = nf, fre_max, speed, ch_mode
        nele, nazi = ndoa_candidate
        ele = np.linspace(0, np.pi, nele)
        azi = np.linspace(-np.pi, np.pi, nazi)
        unit = np.stack([np.outer(np.sin(ele), np.cos(azi)), np.outer(np.sin(ele), np.sin(azi)),
                         np.repeat(np.cos(ele)[:, None], nazi, axis=1)], axis=2)          # (nele, nazi, 3)
        self.dpipd_template = self._phase_pairs(unit, sign=1.0)
        self.doa_candidate = [ele, azi]

    def _phase_pairs(self, unit: np.ndarray, sign: float) -> np.ndarray:
        """exp(j * IPD) for all ordered mic pairs, reduced to the pair set of ch_mode.  unit: (..., 3) direction vectors."""
        mics = self.mic_location
        nmic = mics.shape[-2]
        fre = np.linspace(0.0, self.fre_max, self.nf)
        ipd = np.empty(unit.shape[:-1] + (self.nf, nmic, nmic))
        for m1 in range(nmic):
            for m2 in range(nmic):
                itd = np.dot(unit, mics[m2, :] - mics[m1, :]) / self.speed
                ipd[..., m1, m2] = sign * (-2 * np.pi) * fre * itd[..., None]
        return self.data_adjust(np.exp(1j * ipd))

    def data_adjust(self, data):
        if self.ch_mode == 'M':
            return data[..., 0, 1:]
        if self.ch_mode == 'MM':
            nmic = data.shape[-1]
            return np.stack([data[..., i, j] for i in range(nmic - 1) for j in range(i + 1, nmic)], axis=-1).astype(np.complex64)
        raise Exception('Microphone channel mode unrecognised')

    def forward(self, source_doa=None):
        dpipd = None
        if source_doa is not None:                       # (nb, ntimestep, 2, nsource) [ele, azi]
            d = np.asarray(source_doa).transpose(0, 1, 3, 2)
            unit = np.stack([np.sin(d[..., 0]) * np.cos(d[..., 1]), np.sin(d[..., 0]) * np.sin(d[..., 1]), np.cos(d[..., 0])], axis=3)
            # reference :479-482: ITD taken as (m1 - m2) and the phase multiplied by -1 again -> same sign as the template
            dpipd = self._phase_pairs(unit, sign=1.0).transpose(0, 1, 3, 4, 2)   # (nb, ntime, nf, npairs, nsource)
        return self.dpipd_template, dpipd, self.doa_candidate


class SourceDetectLocalize(nn.Module):
    """Iterative source detection and localisation from predicted DP-IPDs (reference :516-646, meth_mode 'IDL').
    forward(pred_ipd (nb,nt,2nf,npairs), dpipd_template (nele,nazi,2nf,npairs), doa_candidate) ->
    (pred_DOAs (nb,nt,2,ns) [ele, azi] in radians, pred_VADs (nb,nt,ns), pred_ss (nb,nt,nele,nazi)).
    The spectrum, arg-max, projection and residual update all run in fnssl_doa_decode_idl (no per-frame host loop)."""

    def __init__(self, max_num_sources, source_num_mode='kNum', meth_mode='IDL'):
        super().__init__()
        self.max_num_sources = max_num_sources
        self.source_num_mode = source_num_mode
        self.meth_mode = meth_mode
        self._tcache = None

    @ops.on_tensor_device
    def forward(self, pred_ipd: Tensor, dpipd_template, doa_candidate):
        if self.meth_mode != 'IDL':
            raise Exception("fn_ssl_b200.SourceDetectLocalize: only meth_mode='IDL' is implemented")
        ops._need_cuda(pred_ipd)
        dev = pred_ipd.device
        templ = torch.as_tensor(dpipd_template)
        nb, nt, nf2, P = pred_ipd.shape
        nele, nazi = templ.shape[:2]
        key = (templ.data_ptr(), tuple(templ.shape), str(dev))
        if self._tcache is None or self._tcache[0] != key:
            t = templ.to(dev, torch.float32).reshape(nele * nazi, nf2 * P).contiguous()
            self._tcache = (key, t, t.t().contiguous())
        _, t, tt = self._tcache
        R, K, ncand, ns = nb * nt, nf2 * P, nele * nazi, int(self.max_num_sources)
        x = pred_ipd.detach().float().reshape(R, K).contiguous()
        cur = torch.empty_like(x)
        smap = torch.empty((R, ncand), dtype=torch.float32, device=dev)
        ss = torch.empty((R, ncand), dtype=torch.float32, device=dev)
        idx = torch.empty((R, ns), dtype=torch.int32, device=dev)
        vad = torch.empty((R, ns), dtype=torch.float32, device=dev)
        vad_mode = {'kNum': 1, 'unkNum': 2}.get(self.source_num_mode, 0)     # anything else leaves the VADs at 0 (:576-579)
        ops._count(2 * ns)
        _lib.check(_lib.load().fnssl_doa_decode_idl(x.data_ptr(), t.data_ptr(), tt.data_ptr(), R, K, ncand, ns, vad_mode,
                                                    cur.data_ptr(), smap.data_ptr(), ss.data_ptr(), idx.data_ptr(),
                                                    vad.data_ptr(), ops._stream()))
        ele = torch.as_tensor(np.asarray(doa_candidate[0]), dtype=torch.float32, device=dev)
        azi = torch.as_tensor(np.asarray(doa_candidate[1]), dtype=torch.float32, device=dev)
        li = idx.long()
        doas = torch.stack((ele[li // nazi], azi[li % nazi]), dim=1)             # (R, 2, ns)
        return doas.reshape(nb, nt, 2, ns), vad.reshape(nb, nt, ns), ss.reshape(nb, nt, nele, nazi)


def pred_ipd_to_doa(pred_batch: Tensor, gerdpipd: DPIPD, sourcelocalize: SourceDetectLocalize, ch_mode: str = 'MM',
                    fre_range_used=range(1, 257)):
    """The prediction half of PredDOA.predgt2DOA (reference :696-733): network output (nb*npairs, nt, 2nf) ->
    {'doa', 'vad_sources', 'spatial_spectrum'} on the horizontal-plane, 0..pi azimuth candidate grid."""
    template, _, _ = gerdpipd()
    t = np.concatenate((template.real[:, :, fre_range_used, :], template.imag[:, :, fre_range_used, :]), axis=2).astype(np.float32)
    nele, nazi = t.shape[:2]
    t = t[int((nele - 1) / 2):int((nele - 1) / 2) + 1, int((nazi - 1) / 2):nazi, :, :]
    doa_candidate = [np.linspace(np.pi / 2, np.pi / 2, 1), np.linspace(0, np.pi, 37)]
    nmic = t.shape[-1]
    nb = pred_batch.shape[0] // nmic
    rebatch = RemoveChFromBatch(ch_mode)(pred_batch.detach(), nb).permute(0, 2, 3, 1)      # (nb, nt, 2nf, npairs)
    doas, vads, ss = sourcelocalize(pred_ipd=rebatch, dpipd_template=torch.from_numpy(t), doa_candidate=doa_candidate)
    return {'doa': doas, 'vad_sources': vads, 'spatial_spectrum': ss}
