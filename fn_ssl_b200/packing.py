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


def pack_lstm_whh_t(dirs: Sequence[Tuple[Tensor, Tensor, Tensor, Tensor]]) -> Tensor:
    """weight_hh transposed for the backward pass's recurrent product (fnssl_lstm_backward):
    whh_t [dirs][H (unit j)][H (input k)][4 (gate)] = weight_hh[gate*H + j][k], fp32."""
    H = dirs[0][1].shape[1]
    return torch.stack([wh.detach().float().reshape(4, H, H).permute(1, 2, 0) for (_, wh, _, _) in dirs]).contiguous()


def unpack_lstm_simt_grad(dw: Tensor, num_dirs: int, input_size: int, hidden: int) -> List[Tuple[Tensor, Tensor, Tensor, Tensor]]:
    """Inverse of pack_lstm_simt for a gradient buffer: per direction (d weight_ih (4H, I), d weight_hh (4H, H), d bias_ih (4H),
    d bias_hh (4H)); the two bias gradients are the same tensor values (the forward adds the biases)."""
    I, H = input_size, hidden
    K = I + H
    Kp = (K + 3) // 4 * 4
    w4 = dw[:num_dirs * Kp * H * 4].reshape(num_dirs, Kp, H, 4)
    b4 = dw[num_dirs * Kp * H * 4:].reshape(num_dirs, H, 4)
    out = []
    for d in range(num_dirs):
        gwi = w4[d, :I].permute(2, 1, 0).reshape(4 * H, I).contiguous()        # [k][j][gate] -> [gate*H + j][k]
        gwh = w4[d, I:K].permute(2, 1, 0).reshape(4 * H, H).contiguous()
        gb = b4[d].t().reshape(4 * H).contiguous()
        out.append((gwi, gwh, gb, gb.clone()))
    return out


def pack_lstm_tc(dirs, splits) -> Tensor:
    """tcgen05 engine, fp16.  splits = ((c_real, c_padded), ...) per input source (c_padded % 16 == 0).

    K is cut into 64-column slabs: ceil(c_padded/64) slabs per source (zero columns beyond c_real), then H/64
    slabs for h_{t-1}.  Rows are grouped per accumulator chunk of 32 hidden units in column order
    [chunk][gate i,f,g,o][unit]:   W[d][chunk*128 + gate*32 + u][kcol] = weight[gate*H + chunk*32 + u][k].
    Buffer = fp16 W [dirs][4H][nslabs*64]  ++  fp32 bias (b_ih + b_hh) [dirs][4H] in the same row order."""
    w_ih0, w_hh0 = dirs[0][0], dirs[0][1]
    H = w_hh0.shape[1]
    nch = H // 32
    dev = w_ih0.device
    ws, bs = [], []
    for (wi, wh, bi, bh) in dirs:
        cols, off = [], 0
        for (c_real, c_pad) in splits:
            width = (c_pad + 63) // 64 * 64
            blk = torch.zeros((4 * H, width), dtype=torch.float32, device=dev)
            blk[:, :c_real] = wi[:, off:off + c_real]
            cols.append(blk)
            off += c_real
        assert off == wi.shape[1], "splits do not cover the LSTM input size"
        hw = (H + 63) // 64 * 64
        blk = torch.zeros((4 * H, hw), dtype=torch.float32, device=dev)
        blk[:, :H] = wh
        cols.append(blk)
        wcat = torch.cat(cols, dim=1)                                             # (4H, Kpad), rows gate-major
        kpad = wcat.shape[1]
        wcat = wcat.reshape(4, nch, 32, kpad).permute(1, 0, 2, 3).reshape(4 * H, kpad)
        ws.append(wcat.to(torch.float16).contiguous())
        bs.append((bi + bh).reshape(4, nch, 32).permute(1, 0, 2).reshape(4 * H).float().contiguous())
    wbytes = torch.stack(ws).contiguous().view(torch.uint8).reshape(-1)
    bbytes = torch.stack(bs).contiguous().view(torch.uint8).reshape(-1)
    return torch.cat((wbytes, bbytes)).contiguous()


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def run_lstm(params: "LSTMParams", eng: str, axis: int, src0: Tensor, c0: int, src1, c1: int,
             addend=None, want_h: bool = True, state=None, inplace_addend: bool = False, duplicate: bool = False):
    """Run one LSTM layer with the model-level engine `eng` ("tcgen05" | "simt").  In tcgen05 mode the layer
    uses the tensor-core kernel when it is built for this shape (fnssl_lstm_tc_supported) and the fp32
    CUDA-core kernel (on the same fp16 grids) otherwise.  c0/c1 are the REAL channel counts; fp16 grids are
    zero-padded to multiples of 16 channels, which is what the tensor-core kernel consumes."""
    from . import _lib, ops
    H = params.hidden_size
    if eng == "tcgen05":
        p0, p1 = _pad16(c0), _pad16(c1) if src1 is not None else 0
        ok = (src0.dtype == torch.float16 and src0.shape[-1] >= p0 and (src1 is None or src1.shape[-1] >= p1)
              and _lib.load().fnssl_lstm_tc_supported(H, p0, p1))
        if ok:
            splits = ((c0, p0),) + (((c1, p1),) if src1 is not None else ())
            w = params.packed(ops.ENGINE_TCGEN05, splits)
            return ops.lstm(ops.ENGINE_TCGEN05, axis, src0, p0, src1, p1, w, H, params.num_dirs, addend=addend, want_h=want_h,
                            state=state, inplace_addend=inplace_addend, duplicate=duplicate)
    w = params.packed(ops.ENGINE_SIMT, (c0, c1))
    return ops.lstm(ops.ENGINE_SIMT, axis, src0, c0, src1, c1, w, H, params.num_dirs, addend=addend, want_h=want_h,
                    state=state, duplicate=duplicate)
