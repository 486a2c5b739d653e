"""End-to-end forward: microphone signals -> STFT -> normalised features -> network -> DP-IPD / DOA output.

This is the path `predict_step` runs in the reference (FN-SSL/Lightning/main.py:184-189 and
IPDnet/runIPDnetOn.py:182-186: data_preprocess -> self(in_batch)), with the feature tensor kept on the device
in the channels-last grid the LSTM kernels read (it is never materialised in the reference's (nb,C,F,T) layout).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import config, ops
from .FixedAarryIPDnet import IPDnet
from .Model import FN_SSL

Tensor = torch.Tensor

WIN_LEN, NFFT, HOP = 512, 512, 256          # hard-coded by every reference caller (main.py:38-44)


def data_preprocess_fnssl(mic_sig_batch: Tensor, ch_mode: str = 'MM', eps: float = 1e-6, nor_flag: bool = True,
                          sample_length: int = 298) -> List[Tensor]:
    """Reference-layout features: [(nb*P, 4, 256, nt) f32]  (data_preprocess, FN-SSL/Lightning/main.py:200-225)."""
    spec, magsum = ops.stft(mic_sig_batch, WIN_LEN, HOP, NFFT, want_magsum=True)
    norm = ops.NORM_FORGETTING if nor_flag else ops.NORM_NONE
    _, _, cf = ops.features(spec, magsum, ch_mode, norm, sample_length, eps, torch.float32, want_cfirst=True)
    return [cf]


def data_preprocess_ipdnet(mic_sig_batch: Tensor, eps: float = 1e-6, sample_length: int = 280,
                           offline: bool = False) -> List[Tensor]:
    """[(nb, 2*nch, 256, nt) f32]  (IPDnet/runIPDnetOn.py:240-254; offline norm runIPDnetOff.py:248-251)."""
    spec, magsum = ops.stft(mic_sig_batch, WIN_LEN, HOP, NFFT, want_magsum=True)
    norm = ops.NORM_GLOBAL if offline else ops.NORM_FORGETTING
    _, _, cf = ops.features(spec, magsum, 'ALL', norm, sample_length, eps, torch.float32, want_cfirst=True)
    return [cf]


class _HostStaging:
    """Double-buffered host -> device staging on a side stream: the copy of call i+1 overlaps the kernels of call i
    (every call still copies its own input; a buffer is reused only after the forward that read it has finished)."""

    def __init__(self):
        self.copy = None
        self.bufs = None
        self.idx = 0

    def stage(self, host: Tensor, dev: torch.device) -> Tensor:
        if host.is_cuda:
            return host
        if not host.is_pinned():
            raise RuntimeError("run_host: the input must be a pinned host tensor (tensor.pin_memory())")
        if self.copy is None:
            self.copy = torch.cuda.Stream(device=dev)
        if self.bufs is None or self.bufs[0].shape != host.shape or self.bufs[0].device != dev:
            self.bufs = [torch.empty(host.shape, dtype=torch.float32, device=dev) for _ in range(2)]
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.free = [torch.cuda.Event() for _ in range(2)]
            self.idx = 0
        i = self.idx
        self.idx ^= 1
        main = torch.cuda.current_stream(dev)
        self.copy.wait_event(self.free[i])               # the forward that last read bufs[i] is done (no-op the first time)
        with torch.cuda.stream(self.copy):
            self.bufs[i].copy_(host, non_blocking=True)
            self.ready[i].record(self.copy)
        main.wait_event(self.ready[i])
        self._last = i
        return self.bufs[i]

    def release(self, dev: torch.device) -> None:
        self.free[self._last].record(torch.cuda.current_stream(dev))


class FNSSLPipeline(nn.Module):
    """signal (nb, nsample, nch) f32 [device] -> FN_SSL output (nb*P, nt//12, 512 | 180)."""

    def __init__(self, arch: Optional[FN_SSL] = None, ch_mode: str = 'MM', eps: float = 1e-6, sample_length: int = 298):
        super().__init__()
        self.arch = arch if arch is not None else FN_SSL()
        self.ch_mode, self.eps, self.sample_length = ch_mode, eps, sample_length

    @torch.no_grad()
    def forward(self, signal: Tensor) -> Tensor:
        eng = self.arch._engine()
        # fused front end: the complex spectrum is never written to HBM (ops.stft_features: FFT -> sum|X| -> normaliser ->
        # FFT again -> normalised feature grid)
        g0, _ = ops.stft_features(signal, self.ch_mode, ops.NORM_FORGETTING, self.sample_length, self.eps,
                                  config.grid_dtype(eng), WIN_LEN, HOP, NFFT)
        return self.arch.forward_grid(g0, eng)

    @torch.no_grad()
    def run_host(self, signal_host: Tensor, out_host: Optional[Tensor] = None) -> Tensor:
        """End-to-end call for a serving loop: `signal_host` is a PINNED host tensor (nb, nsample, nch).  The host -> device
        copy runs on a side stream into one of two device buffers, so the copy of the next call overlaps this call's
        kernels; the result is optionally copied back into the pinned `out_host` (asynchronously, on the current stream).
        Returns the device output.  Synchronise the current stream before reading `out_host`."""
        dev = next(self.arch.parameters()).device
        if not hasattr(self, "_staging"):
            self._staging = _HostStaging()
        x = self._staging.stage(signal_host, dev)
        out = self.forward(x)
        if not signal_host.is_cuda:
            self._staging.release(dev)
        if out_host is not None:
            out_host.copy_(out, non_blocking=True)
        return out


class IPDnetPipeline(nn.Module):
    """signal (nb, nsample, nch) -> IPDnet output (nb, nt//12, 512, nch-1, 2).  Online: forgetting norm (L=280);
    offline: utterance-global norm and (optionally) chunk-wise inference."""

    def __init__(self, arch: Optional[IPDnet] = None, eps: float = 1e-6, sample_length: int = 280):
        super().__init__()
        self.arch = arch if arch is not None else IPDnet()
        self.eps, self.sample_length = eps, sample_length

    @torch.no_grad()
    def forward(self, signal: Tensor, offline_inference: bool = False) -> Tensor:
        eng = self.arch._engine()
        offline = not self.arch.is_online
        norm = ops.NORM_GLOBAL if offline else ops.NORM_FORGETTING
        g0, _ = ops.stft_features(signal, 'ALL', norm, self.sample_length, self.eps, config.grid_dtype(eng), WIN_LEN, HOP, NFFT)
        nt = g0.shape[1]
        chunked = offline and offline_inference
        if chunked and nt % self.arch.n:
            pad = self.arch.n - nt % self.arch.n
            g0 = torch.cat((g0, g0.new_zeros(g0.shape[0], pad, g0.shape[2], g0.shape[3])), dim=1)
        return self.arch.forward_grid(g0, eng, nt, chunked)
