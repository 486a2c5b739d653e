"""Chunked (streaming) inference for the ONLINE FN-SSL model -- SURVEY.md §8(f) "next" row: the stateful API.

The reference only ever runs whole clips (predict_step, FN-SSL/Lightning/main.py:184-189), but its online
variant is causal by construction, so the same numbers can be produced chunk by chunk if three pieces of state
are carried:

  * the STFT's 256-sample frame overlap                         (Module.py:48-68, center=False, hop = win/2)
  * the forgetting-norm recursion mu_{t-1} and the frame index  (utils_.py:27-44: a_t depends on t)
  * (h, c) of the three uni-directional narrow-band LSTMs       (Model.py:46; nn.LSTM's h_0/c_0 argument)

The full-band BiLSTM runs along frequency inside one frame (Model.py:38) and the head pools 12 consecutive frames
(Model.py:79-80), so neither carries state as long as chunks are cut at multiples of 12 frames -- which
`FNSSLStream.push` does itself: it accepts any number of samples and emits an output block whenever 12 more
frames are complete.  Feeding a clip in pieces yields exactly the whole-clip output (tests/test_gpu_parity.py).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib, config, ops
from .FixedAarryIPDnet import IPDnet
from .Model import FN_SSL
from .pipeline import HOP, NFFT, WIN_LEN

Tensor = torch.Tensor
POOL = 12          # AvgPool2d((12, 1)), Model.py:69


class FNSSLStream:
    """State of `nb` parallel microphone streams run through one online FN_SSL.

        stream = FNSSLStream(arch, nb=8, nch=2)
        for block in microphone:                      # (nb, n, nch) float32 on the GPU, any n
            out = stream.push(block)                  # None, or (nb*P, k, 512) for the k newly completed output frames
    """

    def __init__(self, arch: FN_SSL, nb: int, nch: int = 2, ch_mode: str = 'MM', eps: float = 1e-6,
                 sample_length: int = 298, device: Optional[torch.device] = None):
        if not arch.is_online:
            raise RuntimeError("FNSSLStream: the offline model (bidirectional narrow-band LSTM) needs the whole clip; "
                               "only FN_SSL(is_online=True) can be streamed")
        if arch.training:
            raise RuntimeError("FNSSLStream: call arch.eval() first")
        self.arch, self.nb, self.nch = arch, nb, nch
        self.ch_mode, self.eps, self.sample_length = ch_mode, eps, sample_length
        self.device = torch.device(device) if device is not None else next(arch.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("FNSSLStream: the model must live on a CUDA device (there is no CPU path)")
        self.rows = _lib.load().fnssl_feature_rows(nb, nch, ops.PAIRING[ch_mode])
        self.reset()

    def reset(self) -> None:
        """Forget everything: the next sample pushed is sample 0 of a new clip."""
        dev = self.device
        self.frames_done = 0
        self._pending = torch.empty((self.nb, 0, self.nch), dtype=torch.float32, device=dev)
        self._mu = torch.zeros((self.rows,), dtype=torch.float32, device=dev)
        nf = NFFT // 2
        self._states: List[Tuple[Tensor, Tensor]] = []
        for blk in (self.arch.block_1, self.arch.block_2, self.arch.block_3):
            H = blk.narr_hidden_size
            self._states.append((torch.zeros((self.rows * nf, H), dtype=torch.float32, device=dev),
                                 torch.zeros((self.rows * nf, H), dtype=torch.float32, device=dev)))

    @property
    def pending_samples(self) -> int:
        return self._pending.shape[1]

    @torch.no_grad()
    def push(self, samples: Tensor) -> Optional[Tensor]:
        if samples.dim() != 3 or samples.shape[0] != self.nb or samples.shape[2] != self.nch:
            raise RuntimeError(f"FNSSLStream.push: expected ({self.nb}, n, {self.nch}), got {tuple(samples.shape)}")
        if samples.device != self.device:
            raise RuntimeError("FNSSLStream.push: samples must be on the stream's CUDA device")
        buf = torch.cat((self._pending, samples.float()), dim=1) if self._pending.shape[1] else samples.float()
        n = buf.shape[1]
        frames = (n - WIN_LEN) // HOP + 1 if n >= WIN_LEN else 0
        k = frames // POOL * POOL
        if k == 0:
            self._pending = buf.contiguous()
            return None
        used = buf[:, :HOP * (k - 1) + WIN_LEN].contiguous()
        self._pending = buf[:, HOP * k:].contiguous()       # the next frame starts HOP*k samples in (keeps the overlap)
        eng = self.arch._engine()
        spec, magsum = ops.stft(used, WIN_LEN, HOP, NFFT, want_magsum=True)
        mu = ops.norm_stream(magsum, self.ch_mode, self.sample_length, self.frames_done, self._mu)
        g0, _, _ = ops.features(spec, None, self.ch_mode, ops.NORM_GIVEN, self.sample_length, self.eps,
                                config.grid_dtype(eng), mu=mu)
        out = self.arch.forward_grid(g0, eng, states=self._states)
        self.frames_done += k
        return out


class IPDnetStream:
    """The same for the online IPDnet (IPDnet/runIPDnetOn.py:182-186, 240-254): carried state = STFT overlap,
    forgetting norm (sample_length 280, all microphones as channels), (h, c) of the two narrow-band LSTMs
    (FixedAarryIPDnet.py:36) and the last 36 input frames of the causal conv block (FixedAarryIPDnet.py:61-73).

        out = stream.push(block)      # None, or (nb, k, 512, M-1, 2) for the k newly completed output frames
    """

    def __init__(self, arch: IPDnet, nb: int, eps: float = 1e-6, sample_length: int = 280,
                 device: Optional[torch.device] = None):
        if not arch.is_online:
            raise RuntimeError("IPDnetStream: the offline model needs the whole clip; only IPDnet(is_online=True) can be streamed")
        if arch.training:
            raise RuntimeError("IPDnetStream: call arch.eval() first")
        self.arch, self.nb, self.nch = arch, nb, arch.input_size // 2
        self.eps, self.sample_length = eps, sample_length
        self.device = torch.device(device) if device is not None else next(arch.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("IPDnetStream: the model must live on a CUDA device (there is no CPU path)")
        self.reset()

    def reset(self) -> None:
        dev = self.device
        self.frames_done = 0
        self._pending = torch.empty((self.nb, 0, self.nch), dtype=torch.float32, device=dev)
        self._mu = torch.zeros((self.nb,), dtype=torch.float32, device=dev)
        nf = NFFT // 2
        lstm = []
        for blk in (self.arch.block_1, self.arch.block_2):
            H = blk.narr_hidden_size
            lstm.append((torch.zeros((self.nb * nf, H), dtype=torch.float32, device=dev),
                         torch.zeros((self.nb * nf, H), dtype=torch.float32, device=dev)))
        self._state = {"lstm": lstm, "conv": None}

    @property
    def pending_samples(self) -> int:
        return self._pending.shape[1]

    @torch.no_grad()
    def push(self, samples: Tensor) -> Optional[Tensor]:
        if samples.dim() != 3 or samples.shape[0] != self.nb or samples.shape[2] != self.nch:
            raise RuntimeError(f"IPDnetStream.push: expected ({self.nb}, n, {self.nch}), got {tuple(samples.shape)}")
        if samples.device != self.device:
            raise RuntimeError("IPDnetStream.push: samples must be on the stream's CUDA device")
        buf = torch.cat((self._pending, samples.float()), dim=1) if self._pending.shape[1] else samples.float()
        n = buf.shape[1]
        frames = (n - WIN_LEN) // HOP + 1 if n >= WIN_LEN else 0
        k = frames // POOL * POOL
        if k == 0:
            self._pending = buf.contiguous()
            return None
        used = buf[:, :HOP * (k - 1) + WIN_LEN].contiguous()
        self._pending = buf[:, HOP * k:].contiguous()
        eng = self.arch._engine()
        spec, magsum = ops.stft(used, WIN_LEN, HOP, NFFT, want_magsum=True)
        mu = ops.norm_stream(magsum, 'ALL', self.sample_length, self.frames_done, self._mu)
        g0, _, _ = ops.features(spec, None, 'ALL', ops.NORM_GIVEN, self.sample_length, self.eps,
                                config.grid_dtype(eng), mu=mu)
        out = self.arch.forward_grid(g0, eng, k, False, stream=self._state)
        self.frames_done += k
        return out
