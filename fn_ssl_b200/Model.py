"""FN-SSL network: drop-in for FN-SSL/Model.py and FN-SSL/Lightning/Model.py (same class names, constructor
arguments, forward layouts and state_dict keys), executed by the sm_100a kernels of libfnssl_b200.so.

    FNblock   (reference FN-SSL/Lightning/Model.py:6-50)   alias FullNarrowBlock
    FN_SSL    (:53-90)
    FN_lightning (FN-SSL/Model.py:92-99)

Eval mode is the accelerated inference path (tensor-core or fp32 engine, fused residuals).  Train mode (`.train()`) runs the
differentiable fp32 path of fn_ssl_b200.training -- every LSTM layer and the DP-IPD head are CUDA kernels with hand-written
backward passes, the residual adds and dropout between them are torch elementwise ops -- so `loss.backward()` fills the
`.grad` of every parameter as it does for the reference modules (FN-SSL/Lightning/main.py:95-109).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import config, ops
from .packing import LSTMParams, run_lstm

Tensor = torch.Tensor


def _require_eval(m: nn.Module) -> None:
    if m.training:
        raise RuntimeError(f"{type(m).__name__}: fn_ssl_b200 implements the inference forward only; call .eval() "
                           "(train-mode dropout / backward are outside the accelerated path)")


class FNblock(nn.Module):
    """Full-band BiLSTM (along frequency) + narrow-band LSTM (along time) fusion block.

    forward(x: (nb, nt, nf, nc), nb_skip=None, fb_skip=None) -> (x (nb,nt,nf,256), fb_skip (nb*nt,nf,256),
    nb_skip (nb*nf,nt,256)) exactly as the reference (Model.py:31-50); note the reference recomputes the
    narrow-band skip from the block input (:34), so the `nb_skip` argument is ignored there and here."""

    def __init__(self, input_size, hidden_size=256, dropout=0.2, is_online=False, is_first=False):
        super().__init__()
        self.input_size = input_size
        self.full_hidden_size = hidden_size // 2
        self.is_first = is_first
        self.is_online = is_online
        self.narr_hidden_size = hidden_size if is_online else hidden_size // 2
        self.dropout = dropout
        self.dropout_full = nn.Dropout(p=dropout)   # kept for module-tree parity; identity in eval
        self.dropout_narr = nn.Dropout(p=dropout)
        self.fullLstm = LSTMParams(input_size, self.full_hidden_size, bidirectional=True)
        narr_in = 2 * self.full_hidden_size + (input_size if is_first else 0)
        self.narrLstm = LSTMParams(narr_in, self.narr_hidden_size, bidirectional=not is_online)
        self.engine = None   # None -> config.DEFAULT_ENGINE

    # -- grid-level forward used by FN_SSL (everything stays channels-last on the device) ----------------
    def _run(self, eng: str, x_full_in: Tensor, c_in: int, raw: Optional[Tensor], narr_addend: Optional[Tensor],
             next_full_addend: bool, need_fb: bool, state=None):
        """x_full_in: operand grid of the full-band pass (block input, or block input + fb_skip).
        raw: first block only -- the raw feature grid concatenated to the narrow-band input.
        narr_addend: non-first blocks -- the block input (narrow-band residual, :44-45); CONSUMED (overwritten with
            full-band output + residual when the tensor-core engine runs the layer).
        state: optional (h, c) of the narrow-band LSTM carried across chunks of a stream (online blocks only).
        Returns (N, N + F [if next_full_addend], F)."""
        fh = self.full_hidden_size
        if self.is_first:
            # F_ is the narrow-band layer's INPUT, so the residual sum N + F cannot be accumulated onto it in place.  The
            # full-band layer writes its output twice (a second TMA tile store per quadrant out of the same shared-memory
            # tile: the 16-channel layer is epilogue-bound, its stores are free) and the sum is accumulated onto the copy
            # with the TMA reduce-add path instead of per-thread read-modify-write stores from the epilogue warps
            # (round 1: a separate HBM-rate copy kernel, 2.8 ms = 3.9 % of a cfg4 step).
            dup = bool(next_full_addend) and eng == "tcgen05"
            F_, F2 = run_lstm(self.fullLstm, eng, ops.ALONG_FREQ, x_full_in, c_in, None, 0, duplicate=dup)
            addend, inplace = (F_ if next_full_addend else None), False
            if dup:
                addend, inplace = F2, True
            N_, S_ = run_lstm(self.narrLstm, eng, ops.ALONG_TIME, F_, 2 * fh, raw, c_in,
                              addend=addend, state=state, inplace_addend=inplace)
        else:
            # Both residual sums are accumulated in place (TMA reduce-add in the tensor-core kernel): narr_addend is dead
            # after the full-band layer and F_ after the narrow-band one, and neither is an input of the layer that
            # overwrites it.  (The first block's narrow-band layer reads F_ as its input, so its sum is a new grid.)
            F_, U_ = run_lstm(self.fullLstm, eng, ops.ALONG_FREQ, x_full_in, c_in, None, 0, addend=narr_addend,
                              want_h=need_fb, inplace_addend=True)
            N_, S_ = run_lstm(self.narrLstm, eng, ops.ALONG_TIME, U_, 2 * fh, None, 0,
                              addend=F_ if next_full_addend else None, state=state, inplace_addend=True)
            if next_full_addend:
                F_ = None        # overwritten by S_
        return N_, S_, F_

    def _run_train(self, x: Tensor, fb_skip: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        """Train-mode block on fp32 grids, the reference's arithmetic line by line (Model.py:31-50): x (nb, nt, nf, ld >= nc),
        fb_skip = the previous block's full-band output before dropout -> (dropout(narrow-band output), this block's fb_skip)."""
        from . import training as T
        nc = self.input_size
        xin = x if self.is_first else x + fb_skip                                   # :36-37 (nb_skip = the block input itself, :34)
        F_ = T.lstm_layer(xin, nc, None, 0, self.fullLstm, ops.ALONG_FREQ)          # :38
        Fd = self.dropout_full(F_)                                                  # :40
        if self.is_first:
            N_ = T.lstm_layer(Fd, 2 * self.full_hidden_size, x, nc, self.narrLstm, ops.ALONG_TIME)    # cat, :42-43,46
        else:
            N_ = T.lstm_layer(Fd + x, 2 * self.full_hidden_size, None, 0, self.narrLstm, ops.ALONG_TIME)   # :44-46
        return self.dropout_narr(N_), F_

    def forward(self, x: Tensor, nb_skip: Optional[Tensor] = None, fb_skip: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor, Tensor]:
        nb, nt, nf, nc = x.shape
        if self.training:
            if not x.is_cuda:
                raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only -- no CPU fallback exists")
            if not self.is_first and fb_skip is None:
                raise RuntimeError("FNblock: fb_skip is required when is_first=False")
            xg = x.float().contiguous()
            y, fb = self._run_train(xg, None if self.is_first else fb_skip.reshape(nb, nt, nf, -1).float())
            N_ = y
            return y, fb.reshape(nb * nt, nf, -1), N_.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
        eng = config.resolve(self.engine, (self.full_hidden_size, self.narr_hidden_size))
        dt = config.grid_dtype(eng)
        xg = ops.grid_copy(x, nc, dt)                                   # (nb,nt,nf,ld) grid of the engine's dtype
        if self.is_first:
            N_, _, F_ = self._run(eng, xg, nc, xg, None, False, True)
        else:
            if fb_skip is None:
                raise RuntimeError("FNblock: fb_skip is required when is_first=False")
            fbg = ops.grid_copy(fb_skip.reshape(nb, nt, nf, -1), nc, dt)
            xin = ops.grid_add(xg, fbg)                                  # x + fb_skip (:36-37)
            N_, _, F_ = self._run(eng, xin, nc, None, xg, False, True)
        y = N_.float()
        fb = F_.float().reshape(nb * nt, nf, -1)
        nbs = y.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
        return y, fb, nbs


class FN_SSL(nn.Module):
    """FN-SSL: 3 FN blocks + DP-IPD head (+ optional DOA classifier).

    forward(x: (nb, 4, nf, nt) f32) -> (nb, nt//12, 2*nf)  [or (nb, nt//12, 180) if is_doa]  (Model.py:72-90)."""

    def __init__(self, input_size=4, hidden_size=256, is_online=True, is_doa=False):
        super().__init__()
        self.is_online = is_online
        self.is_doa = is_doa
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.block_1 = FNblock(input_size=input_size, is_online=is_online, is_first=True)
        self.block_2 = FNblock(input_size=hidden_size, is_online=is_online, is_first=False)
        self.block_3 = FNblock(input_size=hidden_size, is_online=is_online, is_first=False)
        self.emb2ipd = nn.Linear(256, 2)
        self.pooling = nn.AvgPool2d(kernel_size=(12, 1))   # module-tree parity only; fused into the head kernel
        self.tanh = nn.Tanh()
        if self.is_doa:
            self.ipd2doa = nn.Linear(512, 180)
        self.engine = None

    def _engine(self) -> str:
        b = self.block_1
        return config.resolve(self.engine, (b.full_hidden_size, b.narr_hidden_size))

    def forward_grid(self, g0: Tensor, eng: Optional[str] = None, states=None) -> Tensor:
        """g0: feature grid (nb, nt, nf, ld) already in the engine's dtype (fused front-end path).
        states: optional list of three (h, c) pairs -- the narrow-band LSTM states of a stream (fn_ssl_b200.streaming)."""
        if self.training:
            if states is not None or g0.dtype != torch.float32:
                raise RuntimeError("FN_SSL: train mode runs whole clips on float32 grids (fn_ssl_b200.training)")
            return self._forward_train(g0)
        eng = eng or self._engine()
        ci = self.input_size
        st = states or (None, None, None)
        N1, S1, _ = self.block_1._run(eng, g0, ci, g0, None, True, True, state=st[0])
        N2, S2, _ = self.block_2._run(eng, S1, self.hidden_size, None, N1, True, True, state=st[1])
        N3, _, _ = self.block_3._run(eng, S2, self.hidden_size, None, N2, False, False, state=st[2])
        out = ops.ipd_head(N3, N3.shape[-1], self.emb2ipd.weight, self.emb2ipd.bias)
        if self.is_doa:
            out = ops.linear(out, self.ipd2doa.weight, self.ipd2doa.bias)
        return out

    def _forward_train(self, g0: Tensor) -> Tensor:
        """Differentiable forward on an fp32 grid (nb, nt, nf, ld >= input_size) -- Model.py:72-90 in train mode."""
        from . import training as T
        x, fb = self.block_1._run_train(g0, None)
        x, fb = self.block_2._run_train(x, fb)
        x, fb = self.block_3._run_train(x, fb)
        out = T.ipd_head_train(x, self.emb2ipd.weight, self.emb2ipd.bias)
        if self.is_doa:
            out = T.linear_train(out, self.ipd2doa.weight, self.ipd2doa.bias)
        return out

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 4 or x.shape[1] != self.input_size:
            raise RuntimeError(f"FN_SSL: expected (nb, {self.input_size}, nf, nt), got {tuple(x.shape)}")
        if self.training:
            if not x.is_cuda:
                raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only -- no CPU fallback exists")
            return self._forward_train(x.float().permute(0, 3, 2, 1).contiguous())     # Model.py:73
        eng = self._engine()
        g0 = ops.cfirst_to_grid(x, config.grid_dtype(eng))              # x.permute(0,3,2,1), Model.py:73
        return self.forward_grid(g0, eng)


class FN_lightning(nn.Module):
    """Wrapper whose state_dict keys are prefixed `arch.` so Lightning checkpoints load (FN-SSL/Model.py:92-99)."""

    def __init__(self):
        super().__init__()
        self.arch = FN_SSL()

    def forward(self, x):
        return self.arch(x)


FullNarrowBlock = FNblock   # name used by BASELINE.json's north_star
