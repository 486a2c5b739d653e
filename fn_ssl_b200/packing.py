"""Weight packing: nn.LSTM-shaped parameters -> the engine-specific buffers fnssl_lstm_forward reads.

SIMT engine (FNSSL_ENGINE_SIMT), fp32:
    w4    [dirs][Kp][H] float4 = (W_i, W_f, W_g, W_o)[k][j]   k < I: weight_ih[g*H+j, k]; k >= I: weight_hh[g*H+j, k-I]
    bias4 [dirs][H]     float4 = b_ih + b_hh per gate           Kp = K rounded up to 4, zero rows beyond K = I + H
tcgen05 engine (FNSSL_ENGINE_TCGEN05), fp16: see pack_lstm_tc.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

Tensor = torch.Tensor


class LSTMParams(nn.Module):
    """Parameter holder with nn.LSTM's exact names, shapes, registration order and default init
    (1 layer, batch_first): weight_ih_l0 (4H,in), weight_hh_l0 (4H,H), bias_ih_l0, bias_hh_l0 [+ _reverse],
    U(-1/sqrt(H), 1/sqrt(H)).  Reference checkpoints therefore load with strict=True and the same
    torch.manual_seed gives the same weights as the reference module (FN-SSL/Lightning/Model.py:25-29)."""

    def __init__(self, input_size: int, hidden_size: int, bidirectional: bool = False):
        super().__init__()
        self.input_size, self.hidden_size, self.bidirectional = input_size, hidden_size, bidirectional
        for suf in (["", "_reverse"] if bidirectional else [""]):
            self.register_parameter("weight_ih_l0" + suf, nn.Parameter(torch.empty(4 * hidden_size, input_size)))
            self.register_parameter("weight_hh_l0" + suf, nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
            self.register_parameter("bias_ih_l0" + suf, nn.Parameter(torch.empty(4 * hidden_size)))
            self.register_parameter("bias_hh_l0" + suf, nn.Parameter(torch.empty(4 * hidden_size)))
        self.reset_parameters()
        self._packed = {}

    def reset_parameters(self) -> None:
        k = 1.0 / math.sqrt(self.hidden_size)
        for w in self.parameters():
            nn.init.uniform_(w, -k, k)

    @property
    def num_dirs(self) -> int:
        return 2 if self.bidirectional else 1

    def directions(self) -> List[Tuple[Tensor, Tensor, Tensor, Tensor]]:
        out = []
        for suf in (["", "_reverse"] if self.bidirectional else [""]):
            out.append(tuple(getattr(self, n + suf) for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")))
        return out

    def packed(self, engine: int, splits: Sequence[int]) -> Tensor:
        """Packed device buffer for `engine`; re-packed when a parameter changed (version counters) or moved."""
        ps = list(self.parameters())
        key = (engine, tuple(splits), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), str(ps[0].device))
        hit = self._packed.get(engine)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            dirs = [tuple(t.detach().float() for t in d) for d in self.directions()]
            buf = pack_lstm_simt(dirs) if engine == 0 else pack_lstm_tc(dirs, splits)
        self._packed[engine] = (key, buf)
        return buf

    def extra_repr(self) -> str:
        return f"{self.input_size}, {self.hidden_size}, batch_first=True, bidirectional={self.bidirectional}"


def pack_lstm_simt(dirs: Sequence[Tuple[Tensor, Tensor, Tensor, Tensor]]) -> Tensor:
    w_ih, w_hh = dirs[0][0], dirs[0][1]
    H, I = w_hh.shape[1], w_ih.shape[1]
    K = I + H
    Kp = (K + 3) // 4 * 4
    dev = w_ih.device
    w4 = torch.zeros((len(dirs), Kp, H, 4), dtype=torch.float32, device=dev)
    b4 = torch.empty((len(dirs), H, 4), dtype=torch.float32, device=dev)
    for d, (wi, wh, bi, bh) in enumerate(dirs):
        w4[d, :I] = wi.reshape(4, H, I).permute(2, 1, 0)      # [k][j][gate]
        w4[d, I:K] = wh.reshape(4, H, H).permute(2, 1, 0)
        b4[d] = (bi + bh).reshape(4, H).t()
    return torch.cat((w4.reshape(-1), b4.reshape(-1))).contiguous()


def pack_lstm_tc(dirs, splits) -> Tensor:
    raise RuntimeError("tcgen05 weight packing is not available in this build")
