"""IPDnet (fixed array): drop-in for IPDnet/FixedAarryIPDnet.py (file name spelled as in the reference).

    FNblock       (reference :7-40)      CausCnnBlock (:42-73)  alias CausalConv1dBlock
    IPDnet        (:76-120)              alias FixedArrayIPDnet
Same constructor arguments, forward layouts and state_dict keys.  Eval mode = the accelerated inference path; train mode = the
differentiable fp32 path of fn_ssl_b200.training (LSTM layers and the three causal convs as CUDA kernels with backward passes;
dropout, ReLU, pooling, tanh and the concatenations around them are torch elementwise / view ops).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import config, ops
from .Model import _require_eval
from .packing import LSTMParams, run_lstm

Tensor = torch.Tensor


class FNblock(nn.Module):
    """Full-band BiLSTM + narrow-band LSTM with *concatenated* raw-input skips.
    forward(x (nb,nt,nf,nc), fb_skip (nb*nt,nf,skip), nb_skip (nb*nf,nt,skip)) -> (nb, nt, nf, hidden+skip)."""

    def __init__(self, input_size, hidden_size=128, dropout=0.2, add_skip_dim=4, is_online=False, is_first=False):
        super().__init__()
        self.input_size = input_size
        self.full_hidden_size = hidden_size // 2
        self.is_first = is_first
        self.is_online = is_online
        self.narr_hidden_size = hidden_size if is_online else hidden_size // 2
        self.add_skip_dim = add_skip_dim
        self.dropout = dropout
        self.dropout_full = nn.Dropout(p=dropout)
        self.dropout_narr = nn.Dropout(p=dropout)
        full_in = input_size if is_first else input_size + add_skip_dim
        self.fullLstm = LSTMParams(full_in, self.full_hidden_size, bidirectional=True)
        self.narrLstm = LSTMParams(2 * self.full_hidden_size + add_skip_dim, self.narr_hidden_size,
                                   bidirectional=not is_online)
        self.engine = None

    def _run(self, eng: str, x: Tensor, cx: int, raw: Tensor, craw: int, state=None) -> Tensor:
        """x: grid with cx channels (block 1: the raw grid itself; block 2: previous narrow output), raw: raw grid.
        state: optional (h, c) of the narrow-band LSTM carried across chunks of a stream (online blocks only).
        Returns the narrow-band output grid N (hidden channels); the [N | raw] concat stays virtual."""
        fh = self.full_hidden_size
        if self.is_first:
            F_, _ = run_lstm(self.fullLstm, eng, ops.ALONG_FREQ, x, cx, None, 0)
        else:
            F_, _ = run_lstm(self.fullLstm, eng, ops.ALONG_FREQ, x, cx, raw, craw)
        N_, _ = run_lstm(self.narrLstm, eng, ops.ALONG_TIME, F_, 2 * fh, raw, craw, state=state)
        return N_

    def _run_train(self, x: Tensor, cx: int, x1: Optional[Tensor], cx1: int, raw: Tensor, craw: int) -> Tensor:
        """Train-mode block on fp32 grids (reference :29-39): full-band input = concat(x[cx], x1[cx1]); returns
        dropout(narrow-band output) -- the concatenation with the raw skip (:37) stays virtual, as in `_run`."""
        from . import training as T
        F_ = self.dropout_full(T.lstm_layer(x, cx, x1, cx1, self.fullLstm, ops.ALONG_FREQ))
        N_ = T.lstm_layer(F_, 2 * self.full_hidden_size, raw, craw, self.narrLstm, ops.ALONG_TIME)
        return self.dropout_narr(N_)

    def forward(self, x: Tensor, fb_skip: Tensor, nb_skip: Tensor) -> Tensor:
        nb, nt, nf, nc = x.shape
        if self.training:
            if not x.is_cuda:
                raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only -- no CPU fallback exists")
            skip = fb_skip.reshape(nb, nt, nf, -1).float().contiguous()
            N_ = self._run_train(x.float().contiguous(), nc, None, 0, skip, skip.shape[-1])
            return torch.cat((N_, nb_skip.reshape(nb, nf, nt, -1).permute(0, 2, 1, 3).float()), dim=-1)
        eng = config.resolve(self.engine, (self.full_hidden_size, self.narr_hidden_size))
        dt = config.grid_dtype(eng)
        skip = fb_skip.reshape(nb, nt, nf, -1)
        cs = skip.shape[-1]
        rawg = ops.grid_copy(skip, cs, dt)
        if self.is_first:
            xg, cx = ops.grid_copy(x, nc, dt), nc
        else:   # reference block input = [previous narrow output | raw skip]; feed the two parts separately
            xg, cx = ops.grid_copy(x, nc - cs, dt), nc - cs
        N_ = self._run(eng, xg, cx, rawg, cs)
        return torch.cat((N_.float(), nb_skip.reshape(nb, nf, nt, -1).permute(0, 2, 1, 3).float()), dim=-1)


class CausCnnBlock(nn.Module):
    """3 x (Conv2d 3x3, pad (1,2), no bias, crop 2 frames) with ReLU+AvgPool(1,3), ReLU+AvgPool(1,4), tanh.
    forward(x (nb, C, F, T)) -> (nb, out_dim, F, T//12)."""

    def __init__(self, inp_dim, out_dim, cnn_hidden_dim=128, kernel=(3, 3), stride=(1, 1), padding=(1, 2)):
        super().__init__()
        if tuple(kernel) != (3, 3) or tuple(stride) != (1, 1) or tuple(padding) != (1, 2):
            raise Exception("fn_ssl_b200.CausCnnBlock: only kernel (3,3), stride (1,1), padding (1,2) is implemented")
        self.conv1 = nn.Conv2d(inp_dim, cnn_hidden_dim, kernel_size=kernel, stride=stride, padding=padding, bias=False)
        self.conv2 = nn.Conv2d(cnn_hidden_dim, cnn_hidden_dim, kernel_size=kernel, stride=stride, padding=padding, bias=False)
        self.conv3 = nn.Conv2d(cnn_hidden_dim, out_dim, kernel_size=kernel, stride=stride, padding=padding, bias=False)
        self.pooling1 = nn.AvgPool2d(kernel_size=(1, 3))   # module-tree parity; fused into the conv kernels
        self.pooling2 = nn.AvgPool2d(kernel_size=(1, 4))
        self.pad = padding
        self.relu = nn.ReLU(inplace=True)
        self.tanh = nn.Tanh()

    def forward_grid(self, src0: Tensor, c0: int, src1: Optional[Tensor], c1: int) -> Tensor:
        return ops.causcnn(src0, c0, src1, c1, self.conv1.weight, self.conv2.weight, self.conv3.weight)

    def forward(self, x: Tensor) -> Tensor:
        if self.training:
            from . import training as T
            if not x.is_cuda:
                raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only -- no CPU fallback exists")
            g = x.float().permute(0, 3, 2, 1).contiguous()
            return T.causcnn_train(g, x.shape[1], None, 0, self.conv1.weight, self.conv2.weight, self.conv3.weight)
        g = ops.cfirst_to_grid(x, torch.float32)
        return self.forward_grid(g, x.shape[1], None, 0)


class IPDnet(nn.Module):
    """forward(x (nb, 2M, nf, nt), offline_inference=False) -> (nb, nt//12, 2*nf, M-1, 2)   (reference :91-120)."""

    def __init__(self, input_size=4, hidden_size=128, max_track=2, is_online=True, n_seg=312):
        super().__init__()
        self.is_online = is_online
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.block_1 = FNblock(input_size=input_size, hidden_size=hidden_size, add_skip_dim=input_size,
                               is_online=is_online, is_first=True)
        self.block_2 = FNblock(input_size=hidden_size, hidden_size=hidden_size, add_skip_dim=input_size,
                               is_online=is_online, is_first=False)
        self.cnn_out_dim = 2 * ((input_size // 2) - 1) * max_track
        self.cnn_inp_dim = hidden_size + input_size
        self.conv = CausCnnBlock(inp_dim=self.cnn_inp_dim, out_dim=self.cnn_out_dim)
        self.n = n_seg
        self.engine = None

    def _engine(self) -> str:
        b = self.block_1
        return config.resolve(self.engine, (b.full_hidden_size, b.narr_hidden_size))

    CONV_HISTORY = 36   # frames of conv input a stream keeps: 3 causal 3x3 layers behind pools of 3 and 4 frames

    def forward_grid(self, g0: Tensor, eng: str, nt_real: int, chunked: bool, stream: Optional[dict] = None) -> Tensor:
        """g0: raw feature grid (nb, nt, nf, ld); for chunked offline inference nt is already padded to a multiple of n_seg.
        stream: state of a chunked run of the online model (fn_ssl_b200.streaming.IPDnetStream): "lstm" = two (h, c)
        pairs, "conv" = the last CONV_HISTORY frames of the conv block's two input grids (None at the start of a clip:
        the block is bias-free with ReLU, so an all-zero history is exactly its zero padding)."""
        _require_eval(self)
        nb, nt, nf, _ = g0.shape
        ci = self.input_size
        ou_frame = nt_real // 12
        nseg = 1
        if chunked:                                   # fold ceil(T/n) zero-padded chunks into the batch (:97-101) -- a view
            nseg = nt // self.n
            g0 = g0.reshape(nb * nseg, self.n, nf, g0.shape[-1])
        st = stream["lstm"] if stream is not None else (None, None)
        N1 = self.block_1._run(eng, g0, ci, g0, ci, state=st[0])
        N2 = self.block_2._run(eng, N1, self.hidden_size, g0, ci, state=st[1])
        if stream is not None:
            # conv3's frame q needs pool4 frames q-2..q, each built from 4 pool3 frames that look 2 more frames back:
            # 36 frames (3 output frames) of history make the first NEW output frame exact; the 3 recomputed ones are dropped
            Hh = self.CONV_HISTORY
            hist = stream.get("conv") or [N2.new_zeros((nb, Hh) + tuple(N2.shape[2:])), g0.new_zeros((nb, Hh) + tuple(g0.shape[2:]))]
            N2c, g0c = torch.cat((hist[0], N2), dim=1), torch.cat((hist[1], g0), dim=1)
            stream["conv"] = [N2c[:, -Hh:].contiguous(), g0c[:, -Hh:].contiguous()]
            y = self.conv.forward_grid(N2c, self.hidden_size, g0c, ci)[..., Hh // 12:]
        else:
            y = self.conv.forward_grid(N2, self.hidden_size, g0, ci)                # (nb', cout, nf, nt2)
        nbp, nt2 = y.shape[0], y.shape[3]
        x = y.permute(0, 3, 2, 1).reshape(nbp, nt2, nf, 2, -1).permute(0, 1, 3, 2, 4)   # :114 (tiny output tensor)
        if chunked:
            x = x.reshape(nbp // nseg, nt2 * nseg, 2, nf * 2, -1).permute(0, 1, 3, 4, 2)
            return x[:, :ou_frame, :, :, :]
        return x.reshape(nbp, nt2, 2, nf * 2, -1).permute(0, 1, 3, 4, 2)

    def _forward_train(self, g0: Tensor) -> Tensor:
        """Differentiable forward on an fp32 feature grid (nb, nt, nf, ld >= input_size): reference :91-120 in train mode."""
        from . import training as T
        nb, nt, nf, _ = g0.shape
        ci = self.input_size
        N1 = self.block_1._run_train(g0, ci, None, 0, g0, ci)                     # block 1: x = the raw grid itself
        N2 = self.block_2._run_train(N1, self.hidden_size, g0, ci, g0, ci)        # block 2: x = [N1 | raw]
        y = T.causcnn_train(N2, self.hidden_size, g0, ci, self.conv.conv1.weight, self.conv.conv2.weight, self.conv.conv3.weight)
        nt2 = y.shape[3]
        x = y.permute(0, 3, 2, 1).reshape(nb, nt2, nf, 2, -1).permute(0, 1, 3, 2, 4)     # :114
        return x.reshape(nb, nt2, 2, nf * 2, -1).permute(0, 1, 3, 4, 2)

    def forward(self, x: Tensor, offline_inference: bool = False) -> Tensor:
        if x.dim() != 4 or x.shape[1] != self.input_size:
            raise RuntimeError(f"IPDnet: expected (nb, {self.input_size}, nf, nt), got {tuple(x.shape)}")
        if self.training:
            if not x.is_cuda:
                raise RuntimeError("fn_ssl_b200 runs on CUDA (sm_100a) only -- no CPU fallback exists")
            if offline_inference:
                raise RuntimeError("IPDnet: offline_inference (chunked) is an eval-mode feature")
            return self._forward_train(x.float().permute(0, 3, 2, 1).contiguous())
        eng = self._engine()
        nt = x.shape[3]
        chunked = (not self.is_online) and offline_inference
        nt_alloc = (nt + self.n - 1) // self.n * self.n if chunked else None
        g0 = ops.cfirst_to_grid(x, config.grid_dtype(eng), nt_alloc=nt_alloc)
        return self.forward_grid(g0, eng, nt, chunked)


FixedArrayIPDnet = IPDnet          # names used by BASELINE.json's north_star
CausalConv1dBlock = CausCnnBlock
FullNarrowBlock = FNblock
