"""CPU oracle for the IPDnet2 forward path (OnlineSpatialNet + Mamba, center=True STFT).

THIS FILE IS TEST INFRASTRUCTURE (see oracle/fnssl_oracle.py): only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of the benches may import it.  Nothing under ``fn_ssl_b200/`` may.

Parity status
-------------
* Everything except the Mamba block is PINNED: ``tests/golden/make_golden_ipdnet2.py`` imports the unmodified
  ``IPDnet2/IPDnet2.py`` / ``IPDnet2/Module.py`` / ``IPDnet2/utils_.py`` of the reference and compares them with the
  functions below on seeded inputs (``tests/golden/ipdnet2_golden.npz``).
* The Mamba block is **parity unpinned**: the reference takes it from the third-party package ``mamba_ssm``
  (``IPDnet2/IPDnet2.py:15-19,127,132``), which is neither vendored under /root/reference nor version-pinned anywhere
  in the reference (no entry in any environment file) nor installed in this image.  ``mamba()`` below restates the
  published algorithm of ``mamba_ssm.modules.mamba_simple.Mamba.forward`` (Mamba v1, Gu & Dao 2023, Alg. 2 + the
  package's ``selective_scan_ref``); the golden script injects exactly this function as ``mamba_ssm.Mamba`` into the
  reference model, so the golden vectors pin the reference's *use* of the block (norm, reshape, residual, ordering)
  but not the block's arithmetic.  Parameter names and shapes are pinned by the reference's shipped checkpoint
  ``IPDnet2/checkpoints/ipdnet2_small.ckpt`` (A_log, D, in_proj, conv1d, x_proj, dt_proj, out_proj).

All ``path:line`` citations are relative to the reference checkout.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .fnssl_oracle import forgetting_norm

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# Front end
# --------------------------------------------------------------------------------------

def stft_center_num_frames(nsample: int, hop: int = 320) -> int:
    """nt = floor(nsample / hop + 1)   -- IPDnet2/Module.py:55."""
    return int(math.floor(nsample / hop + 1))


def stft_center(signal: Tensor, win_len: int = 512, win_shift_ratio: float = 0.625, nfft: int = 512) -> Tensor:
    """(nb, nsample, nch) -> (nb, 257, nt, nch) complex64; ``STFT.forward`` IPDnet2/Module.py:46-64:
    torch.stft(center=True) = reflect-pad nfft/2 samples on both sides, then the framing of center=False."""
    nb, nsample, nch = signal.shape
    hop = int(win_len * win_shift_ratio)
    nt = stft_center_num_frames(nsample, hop)
    x = signal.permute(0, 2, 1)                                             # (nb, nch, n)
    x = F.pad(x, (nfft // 2, nfft // 2), mode="reflect")
    window = torch.hann_window(win_len, dtype=signal.dtype)
    frames = x.unfold(-1, win_len, hop)[:, :, :nt, :] * window
    spec = torch.fft.rfft(frames, n=nfft, dim=-1)
    return spec.permute(0, 3, 2, 1).contiguous()


def preprocess_ipdnet2(signal: Tensor, eps: float = 1e-6, sample_length: int = 249) -> Tensor:
    """(nb, nsample, M) -> (nb, 2M, 256, nt); ``data_preprocess`` IPDnet2/run_IPDnet2.py:277-288."""
    spec = stft_center(signal).permute(0, 3, 1, 2)                          # (nb, M, 257, nt)
    mu = forgetting_norm(spec.abs(), sample_length=sample_length)
    feat = torch.cat((spec.real / (mu + eps), spec.imag / (mu + eps)), dim=1)
    return feat[:, :, 1:257, :].contiguous()


# --------------------------------------------------------------------------------------
# Mamba (third party; restated from the published algorithm -- parity unpinned, see header)
# --------------------------------------------------------------------------------------

def mamba(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """(N, L, d_model) -> (N, L, d_model).  mamba_ssm ``Mamba.forward`` without inference cache:
    xz = in_proj(x); x, z = split; x = SiLU(causal depthwise conv1d_k(x) + b); (dt, B, C) = split(x_proj(x));
    delta = softplus(dt_proj(dt) + b_dt); h_t = exp(delta A) h_{t-1} + delta B_t x_t, A = -exp(A_log);
    y_t = <C_t, h_t> + D x_t;  out = out_proj(y * SiLU(z))."""
    A = -torch.exp(sd[prefix + "A_log"].float())                            # (d_inner, d_state)
    D = sd[prefix + "D"].float()
    d_inner, d_state = A.shape
    dt_rank = sd[prefix + "dt_proj.weight"].shape[1]
    N, L, _ = x.shape
    xz = x @ sd[prefix + "in_proj.weight"].t()                              # (N, L, 2 d_inner), no bias
    if prefix + "in_proj.bias" in sd:
        xz = xz + sd[prefix + "in_proj.bias"]
    xi, z = xz[..., :d_inner], xz[..., d_inner:]
    w = sd[prefix + "conv1d.weight"]                                        # (d_inner, 1, k)
    k = w.shape[-1]
    xc = F.conv1d(xi.transpose(1, 2), w, sd[prefix + "conv1d.bias"], padding=k - 1, groups=d_inner)[..., :L]
    xc = F.silu(xc).transpose(1, 2)                                         # (N, L, d_inner)
    x_dbl = xc @ sd[prefix + "x_proj.weight"].t()                           # (N, L, dt_rank + 2 d_state)
    dt, Bm, Cm = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)
    delta = F.softplus(dt @ sd[prefix + "dt_proj.weight"].t() + sd[prefix + "dt_proj.bias"])   # (N, L, d_inner)
    h = x.new_zeros(N, d_inner, d_state)
    ys = []
    for t in range(L):
        dA = torch.exp(delta[:, t, :, None] * A)                            # (N, d_inner, d_state)
        dBx = delta[:, t, :, None] * Bm[:, t, None, :] * xc[:, t, :, None]
        h = dA * h + dBx
        ys.append((h * Cm[:, t, None, :]).sum(-1))
    y = torch.stack(ys, dim=1) + xc * D
    y = y * F.silu(z)
    out = y @ sd[prefix + "out_proj.weight"].t()
    if prefix + "out_proj.bias" in sd:
        out = out + sd[prefix + "out_proj.bias"]
    return out


# --------------------------------------------------------------------------------------
# OnlineSpatialNet
# --------------------------------------------------------------------------------------

def _layer_norm(x: Tensor, sd: StateDict, name: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _fconv(x: Tensor, sd: StateDict, prefix: str, groups: int = 8) -> Tensor:
    """``SpatialNetLayer._fconv`` IPDnet2.py:222-233 on (B, F, T, H): LN over H, grouped Conv1d along F ('same' zero
    padding), per-channel PReLU."""
    B, Fq, T, H = x.shape
    y = _layer_norm(x, sd, prefix + ".0")                                   # LayerNorm(seq_last=True) == LN over H
    y = y.permute(0, 2, 3, 1).reshape(B * T, H, Fq)
    w = sd[prefix + ".1.weight"]
    y = F.conv1d(y, w, sd[prefix + ".1.bias"], padding=w.shape[-1] // 2, groups=groups)
    y = F.prelu(y, sd[prefix + ".2.weight"])
    return y.reshape(B, T, H, Fq).permute(0, 3, 1, 2)


def _full(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """``SpatialNetLayer._full`` IPDnet2.py:235-253: LN, squeeze 1x1 conv + SiLU, Linear along F, unsqueeze + SiLU."""
    B, Fq, T, H = x.shape
    y = _layer_norm(x, sd, prefix + "norm_full")
    y = y.permute(0, 2, 3, 1).reshape(B * T, H, Fq)
    y = F.silu(F.conv1d(y, sd[prefix + "squeeze.0.weight"], sd[prefix + "squeeze.0.bias"]))
    y = F.linear(y, sd[prefix + "full.weight"], sd[prefix + "full.bias"])
    y = F.silu(F.conv1d(y, sd[prefix + "unsqueeze.0.weight"], sd[prefix + "unsqueeze.0.bias"]))
    return y.reshape(B, T, H, Fq).permute(0, 3, 1, 2)


def _pool_f(x: Tensor, k: int) -> Tensor:
    """AvgPool2d((1, k)) over F of (B, F, T, H) -- IPDnet2.py:134-135,148,153."""
    B, Fq, T, H = x.shape
    return x.reshape(B, Fq // k, k, T, H).mean(dim=2)


def spatialnet_layer(x: Tensor, sd: StateDict, prefix: str, is_first: bool) -> Tensor:
    """``SpatialNetLayer.forward`` IPDnet2.py:137-164 with Mamba for both time modules."""
    x = x + _fconv(x, sd, prefix + "fconv1")
    if is_first:
        x = _pool_f(x, 2)
    x = x + _full(x, sd, prefix)
    x = x + _fconv(x, sd, prefix + "fconv2")
    if is_first:
        x = _pool_f(x, 8)
    B, Fq, T, H = x.shape
    for norm, name in (("norm_mhsa", "mhsa."), ("norm_tconvffn", "tconvffn.")):
        y = _layer_norm(x, sd, prefix + norm).reshape(B * Fq, T, H)
        x = x + mamba(y, sd, prefix + name).reshape(B, Fq, T, H)
    return x


def ipdnet2_forward(x: Tensor, sd: StateDict, prefix: str = "", time_compression_ratio: int = 5,
                    fre_compression_ratio: int = 16, n_src: int = 2) -> Tensor:
    """``OnlineSpatialNet.forward`` IPDnet2.py:331-368.  x (B, 2M, 256, T) -> (B, T//5, 512, dim_output/(2 n_src), n_src).

    ``n_src`` generalises the literal 2 of the output reshape (:363-364); the reference only runs with 2."""
    num_layers = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in sd if k.startswith(prefix + "layers."))
    x = x.permute(0, 2, 3, 1)                                               # (B, F, T, C)
    B, Fq, T, C = x.shape
    w = sd[prefix + "encoder.weight"]                                       # (H, C, k) causal conv along T (:66-76,335)
    y = F.conv1d(F.pad(x.reshape(B * Fq, T, C).permute(0, 2, 1), (w.shape[-1] - 1, 0)), w, sd[prefix + "encoder.bias"])
    x = y.permute(0, 2, 1).reshape(B, Fq, T, -1)
    for i in range(num_layers):
        x = spatialnet_layer(x, sd, f"{prefix}layers.{i}.", is_first=(i == 0))
        if i == 0:                                                          # time_compression_layer = 0 (:342-349)
            Bq, Fc, Tq, H = x.shape
            T5 = Tq // time_compression_ratio
            x = x[:, :, :T5 * time_compression_ratio].reshape(Bq, Fc, T5, time_compression_ratio, H).mean(dim=3)
    # FreqInverse (:37-43): per compressed band a 1x1 conv H -> ratio*out, channel index = o*ratio + j
    Bq, Fc, T5, H = x.shape
    wt = sd[prefix + "freq_inverse.trans2.weight"][:, :, 0]                 # (ratio*out, H)
    r = fre_compression_ratio
    dim_out = wt.shape[0] // r
    y = x @ wt.t() + sd[prefix + "freq_inverse.trans2.bias"]                # (B, Fc, T5, out*r)
    y = y.reshape(Bq, Fc, T5, dim_out, r).permute(0, 1, 4, 2, 3).reshape(Bq, Fc * r, T5, dim_out).tanh()   # (B, F, T5, out)
    y = F.linear(y, sd[prefix + "decoder.weight"], sd[prefix + "decoder.bias"])
    nF = Fc * r
    y = y.permute(0, 2, 1, 3).reshape(Bq, T5, nF, n_src, -1).permute(0, 1, 3, 2, 4)
    y = y.reshape(Bq, T5, n_src, nF * 2, -1).permute(0, 1, 3, 4, 2)
    return y.contiguous()


# --------------------------------------------------------------------------------------
# Seeded weights (same key set / shapes as IPDnet2/checkpoints/ipdnet2_small.ckpt minus the "arch." prefix)
# --------------------------------------------------------------------------------------

def ipdnet2_param_shapes(dim_input: int = 10, dim_output: int = 16, num_layers: int = 8, dim_hidden: int = 96,
                         dim_squeeze: int = 8, num_freqs: int = 256, d_state: int = 16, d_conv: int = 4,
                         f_kernel: int = 5, groups: int = 8, fre_compression_ratio: int = 16,
                         encoder_kernel_size: int = 5):
    H, d_inner = dim_hidden, 2 * dim_hidden
    dt_rank = math.ceil(H / 16)
    shapes = {"encoder.weight": (H, dim_input, encoder_kernel_size), "encoder.bias": (H,)}
    for l in range(num_layers):
        p = f"layers.{l}."
        nf = num_freqs // 2 if l == 0 else num_freqs // fre_compression_ratio
        for fc in ("fconv1", "fconv2"):
            shapes.update({p + fc + ".0.weight": (H,), p + fc + ".0.bias": (H,),
                           p + fc + ".1.weight": (H, H // groups, f_kernel), p + fc + ".1.bias": (H,),
                           p + fc + ".2.weight": (H,)})
        shapes.update({p + "norm_full.weight": (H,), p + "norm_full.bias": (H,),
                       p + "squeeze.0.weight": (dim_squeeze, H, 1), p + "squeeze.0.bias": (dim_squeeze,),
                       p + "full.weight": (nf, nf), p + "full.bias": (nf,),
                       p + "unsqueeze.0.weight": (H, dim_squeeze, 1), p + "unsqueeze.0.bias": (H,)})
        for nm, mb in (("norm_mhsa", "mhsa"), ("norm_tconvffn", "tconvffn")):
            shapes.update({p + nm + ".weight": (H,), p + nm + ".bias": (H,),
                           p + mb + ".A_log": (d_inner, d_state), p + mb + ".D": (d_inner,),
                           p + mb + ".in_proj.weight": (2 * d_inner, H),
                           p + mb + ".conv1d.weight": (d_inner, 1, d_conv), p + mb + ".conv1d.bias": (d_inner,),
                           p + mb + ".x_proj.weight": (dt_rank + 2 * d_state, d_inner),
                           p + mb + ".dt_proj.weight": (d_inner, dt_rank), p + mb + ".dt_proj.bias": (d_inner,),
                           p + mb + ".out_proj.weight": (H, d_inner)})
    shapes.update({"freq_inverse.trans2.weight": (fre_compression_ratio * dim_output, H, 1),
                   "freq_inverse.trans2.bias": (fre_compression_ratio * dim_output,),
                   "decoder.weight": (dim_output, dim_output), "decoder.bias": (dim_output,)})
    return shapes


def seeded_ipdnet2_state_dict(seed: int = 0, **cfg) -> StateDict:
    """Deterministic weights with trained-network-like scales (there is no default init to mirror for the Mamba block:
    its constructor is third-party).  uniform(+-1/sqrt(fan_in)) for matrices / conv kernels, LN weights near 1, PReLU
    slopes near 0.25, A_log = log(1..d_state) and dt bias = softplus^-1 of log-uniform [1e-3, 1e-1] as in Mamba."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for name, shape in ipdnet2_param_shapes(**cfg).items():
        leaf = name.split(".")[-1]
        if name.endswith("A_log"):
            t = torch.log(torch.arange(1, shape[1] + 1, dtype=torch.float32)).repeat(shape[0], 1)
            t = t + 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif name.endswith(".D"):
            t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5)
        elif name.endswith("dt_proj.bias"):
            dt = torch.exp(torch.rand(shape, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
            t = dt + torch.log(-torch.expm1(-dt))
        elif len(shape) == 1 and leaf == "weight" and (".0.weight" in name or "norm_" in name):
            t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5)            # LayerNorm gain
        elif len(shape) == 1 and leaf == "weight":
            t = 0.25 + 0.2 * (torch.rand(shape, generator=g) - 0.5)           # PReLU slope
        elif len(shape) == 1:
            t = 0.2 * (torch.rand(shape, generator=g) - 0.5)                  # biases
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = (2.0 * torch.rand(shape, generator=g) - 1.0) / math.sqrt(fan_in)
        sd[name] = t.float().contiguous()
    return sd
