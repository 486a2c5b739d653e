"""IPDnet2: drop-in for IPDnet2/IPDnet2.py (OnlineSpatialNet with Mamba time modules) -- scope row a11.

Same class names, constructor arguments and state_dict keys as the reference, so the shipped checkpoint
``IPDnet2/checkpoints/ipdnet2_small.ckpt`` (keys ``arch.*``) loads with strict=True into ``IPDnet2_lightning``:

    CausalConv1d      (IPDnet2.py:45-82)      parameter holder; fused into the first frequency-stage launch
    FreqInverse       (:23-43)                parameter holder; fused into the head launch
    Mamba             (mamba_ssm.Mamba, third party, :15-19,127,132)   parameter holder with the package's names
    SpatialNetLayer   (:85-256)               3 launches: frequency stage (fnssl_sn_freq_forward), time stage
                                              (fnssl_sn_time_forward: one launch per Mamba block)
    OnlineSpatialNet  (:259-399)              forward(x: (B, 2M, 256, T)) -> (B, T//5, 512, M-1, 2)

Only the configuration the reference runs is built (run_IPDnet2.py:103-119): dim_hidden 96, dim_squeeze 8, conv groups 8,
kernel 5 along F, norms LN, attention 'mamba(16,4)', 256 frequencies compressed by 2 and 8 in layer 0, time compression
5 after layer 0.  Anything else raises.  Inference only (eval mode), fp32 throughout.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

Tensor = torch.Tensor


def _require_eval(m: nn.Module) -> None:
    if m.training:
        raise RuntimeError(f"{type(m).__name__}: fn_ssl_b200 implements the inference forward only; call .eval()")


class CausalConv1d(nn.Conv1d):
    """Parameter holder of the encoder (IPDnet2.py:45-82); computed inside the first frequency-stage kernel."""

    def __init__(self, in_channels, out_channels, kernel_size, look_ahead: int = 0, **kw):
        super().__init__(in_channels, out_channels, kernel_size, **kw)
        if look_ahead != 0:
            raise Exception("CausalConv1d: only look_ahead=0 (the reference's encoder) is implemented")
        self.look_ahead = look_ahead

    def forward(self, x, state=None):
        raise RuntimeError("CausalConv1d is fused into OnlineSpatialNet's first kernel; call OnlineSpatialNet.forward")


class Mamba(nn.Module):
    """Parameter holder with mamba_ssm.Mamba's parameter names / shapes (v1 block: in_proj, depthwise conv1d, x_proj,
    dt_proj, A_log, D, out_proj); init follows the package's documented scheme (A = 1..d_state, dt in [1e-3, 1e-1])."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", layer_idx=None, **kw):
        super().__init__()
        self.d_model, self.d_state, self.d_conv = d_model, d_state, d_conv
        self.d_inner = expand * d_model
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=False)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, d_conv, groups=self.d_inner, padding=d_conv - 1, bias=True)
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True)
        with torch.no_grad():
            dt = torch.exp(torch.rand(self.d_inner) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3)).clamp(min=1e-4)
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.A_log = nn.Parameter(torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).repeat(self.d_inner, 1))
        self.D = nn.Parameter(torch.ones(self.d_inner))
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)

    def forward(self, x, inference_params=None):
        raise RuntimeError("Mamba is executed inside SpatialNetLayer's time-stage kernel; call the layer / network")


class FreqInverse(nn.Module):
    """Parameter holder (IPDnet2.py:23-43): per-band 1x1 conv hidden -> compression_ratio * out_dim, tanh."""

    def __init__(self, nfreq=256, compression_ratio=16, hidden_dim=96, out_dim=16, sample_rate=16000):
        super().__init__()
        self.nfreq, self.nfilters, self.hidden_dim, self.out_dim = nfreq, nfreq // compression_ratio, hidden_dim, out_dim
        self.compression_ratio = compression_ratio
        self.trans2 = nn.Conv1d(hidden_dim, compression_ratio * out_dim, 1)

    def forward(self, x):
        raise RuntimeError("FreqInverse is fused into OnlineSpatialNet's head kernel; call OnlineSpatialNet.forward")


def freq_stage_work(positions: int, nf: int, x_ld: int, first: bool) -> Tuple[float, float]:
    """Algorithmic (FLOP, HBM bytes) of one frequency-stage launch over `positions` = nb * nt frames."""
    H, gc, k, sq = 96, 12, 5, 8
    if first:       # encoder at 256 bins, fconv1 at 256, full-band + fconv2 at 128, pooled to 16
        mac = 256 * H * 5 * x_ld + 256 * H * gc * k + (128 * H * sq * 2 + sq * 128 * 128) + 128 * H * gc * k
        nbytes = 256 * x_ld * 4 + 16 * H * 4
    else:
        mac = 2 * nf * H * gc * k + (nf * H * sq * 2 + sq * nf * nf)
        nbytes = 2 * nf * H * 4
    return 2.0 * mac * positions, float(nbytes) * positions


def time_stage_work(sequences: int, nt: int, pool: int) -> Tuple[float, float]:
    """Algorithmic (FLOP, HBM bytes) of one time-stage launch: two Mamba blocks over `sequences` x nt frames."""
    H, di, ns, dr = 96, 192, 16, 6
    mac = 2 * (H * 2 * di + di * 4 + di * (dr + 2 * ns) + dr * di + 3 * di * ns + di * H)
    return 2.0 * mac * sequences * nt, float(sequences) * nt * H * 4 * (1.0 + 1.0 / pool)


def _versions(mod: nn.Module):
    ps = list(mod.parameters())
    return (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), str(ps[0].device))


class SpatialNetLayer(nn.Module):
    """One layer (IPDnet2.py:85-256): cross-band conv, full-band linear, cross-band conv along F; two Mamba blocks
    along T.  forward(x: (B, F, T, H)) -> ((B, F, T, H), None) for the non-first layers; the first layer is driven by
    OnlineSpatialNet (its kernel also contains the encoder and both frequency pools)."""

    def __init__(self, dim_hidden: int, dim_squeeze: int, num_freqs: int, dropout=(0, 0, 0), kernel_size=(5, 3),
                 conv_groups=(8, 8), norms=("LN", "LN", "GN", "LN", "LN", "LN"), padding: str = 'zeros', full=None,
                 attention: str = 'mamba(16,4)', is_first: bool = False):
        super().__init__()
        if not attention.startswith('mamba'):
            raise Exception(f"SpatialNetLayer: only attention='mamba(d_state,d_conv)' is implemented (got {attention})")
        if any(n.upper() != 'LN' for i, n in enumerate(norms) if i != 2):
            raise Exception("SpatialNetLayer: only LayerNorm ('LN') norms are implemented")
        if any(d > 0 for d in dropout) or padding != 'zeros' or full is not None:
            raise Exception("SpatialNetLayer: dropout / non-zero padding / shared full-band module are not implemented")
        d_state, d_conv = (int(v) for v in attention[6:-1].split(','))
        H, G, K = dim_hidden, conv_groups[0], kernel_size[0]
        self.dim_hidden, self.dim_squeeze, self.num_freqs, self.is_first = H, dim_squeeze, num_freqs, is_first
        self.groups, self.f_kernel = G, K

        def fconv():
            return nn.ModuleList([nn.LayerNorm(H), nn.Conv1d(H, H, K, groups=G, padding='same'), nn.PReLU(H)])
        self.fconv1 = fconv()
        self.norm_full = nn.LayerNorm(H)
        self.squeeze = nn.Sequential(nn.Conv1d(H, dim_squeeze, 1), nn.SiLU())
        self.full = nn.Linear(num_freqs, num_freqs)
        self.unsqueeze = nn.Sequential(nn.Conv1d(dim_squeeze, H, 1), nn.SiLU())
        self.fconv2 = fconv()
        self.norm_mhsa = nn.LayerNorm(H)
        self.mhsa = Mamba(d_model=H, d_state=d_state, d_conv=d_conv, layer_idx=0)
        self.norm_tconvffn = nn.LayerNorm(H)
        self.tconvffn = Mamba(d_model=H, d_state=d_state, d_conv=d_conv, layer_idx=0)
        self._packed = None

    # ---- weight packing (kernel layouts, see include/fnssl_b200.h) ------------------------------------------------
    def _pack(self, encoder: Optional[CausalConv1d], x_ld: int):
        key = (_versions(self), _versions(encoder) if encoder is not None else None, x_ld)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1:]
        keep: List[Tensor] = []

        def dev(t: Tensor) -> int:
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def fconv_w(ml) -> _lib.SnFconvWeights:
            G, gc = self.groups, self.dim_hidden // self.groups
            w = ml[1].weight.detach().float()                                # (H, gc, K): out = g*gc + o, in i
            wp = w.reshape(G, gc, gc, self.f_kernel).permute(0, 3, 2, 1)      # [g][k][i][o]
            return _lib.SnFconvWeights(dev(ml[0].weight), dev(ml[0].bias), dev(wp), dev(ml[1].bias), dev(ml[2].weight))

        with torch.no_grad():
            fa = _lib.SnFreqArgs()
            fa.hidden, fa.squeeze, fa.groups, fa.fkernel = self.dim_hidden, self.dim_squeeze, self.groups, self.f_kernel
            fa.is_first = int(self.is_first)
            if self.is_first:
                ew = encoder.weight.detach().float()                          # (H, cin, k)
                cin, k = ew.shape[1], ew.shape[2]
                wp = torch.zeros((k, x_ld, self.dim_hidden), dtype=torch.float32, device=ew.device)
                wp[:, :cin] = ew.permute(2, 1, 0)
                fa.enc_wp, fa.enc_b, fa.enc_kernel, fa.cin, fa.x_ld = dev(wp), dev(encoder.bias), k, cin, x_ld
            fa.fconv1, fa.fconv2 = fconv_w(self.fconv1), fconv_w(self.fconv2)
            fa.lnf_w, fa.lnf_b = dev(self.norm_full.weight), dev(self.norm_full.bias)
            fa.sq_wt, fa.sq_b = dev(self.squeeze[0].weight[:, :, 0].t()), dev(self.squeeze[0].bias)
            fa.full_wt, fa.full_b = dev(self.full.weight.t()), dev(self.full.bias)
            fa.usq_w, fa.usq_b = dev(self.unsqueeze[0].weight[:, :, 0]), dev(self.unsqueeze[0].bias)
            ta = _lib.SnTimeArgs()
            mb = self.mhsa
            ta.hidden, ta.d_inner, ta.d_state, ta.dt_rank, ta.d_conv = self.dim_hidden, mb.d_inner, mb.d_state, mb.dt_rank, mb.d_conv
            for i, (norm, m) in enumerate(((self.norm_mhsa, self.mhsa), (self.norm_tconvffn, self.tconvffn))):
                ta.m[i] = _lib.MambaWeights(
                    dev(norm.weight), dev(norm.bias), dev(m.in_proj.weight.t()), dev(m.conv1d.weight[:, 0, :]),
                    dev(m.conv1d.bias), dev(m.x_proj.weight.t()), dev(m.dt_proj.weight), dev(m.dt_proj.bias), dev(m.A_log),
                    dev(m.D), dev(m.out_proj.weight.t()))
        self._packed = (key, fa, ta, keep)
        return fa, ta, keep

    @ops.on_tensor_device
    def _run(self, x: Tensor, nb: int, nt: int, encoder: Optional[CausalConv1d], pool: int, t_begin: int = 0,
             state: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
        """x: feature grid (nb, nt, 256, ld) [first layer] or activation (nb, nt, 16, H); returns (nb, (nt - t_begin) // pool,
        16, H).  t_begin (first layer): leading frames that are only history of the causal encoder.  state: the two Mamba
        blocks' carried (nb*16, 19, 192) f32 buffers, read and updated in place (streaming)."""
        ops._need_cuda(x)
        lib = _lib.load()
        H = self.dim_hidden
        fa, ta, _ = self._pack(encoder, x.shape[-1] if self.is_first else 0)
        nf_in = x.shape[2]
        fa.nb, fa.nt, fa.nf, fa.x, fa.t_begin = nb, nt, nf_in, x.data_ptr(), t_begin
        nt = nt - t_begin
        y = torch.empty((nb, nt, 16, H), dtype=torch.float32, device=x.device)
        fa.out = y.data_ptr()
        ops._count(1)
        flops, nbytes = freq_stage_work(nb * nt, nf_in, x.shape[-1] if self.is_first else 0, self.is_first)
        ops.profiled("sn_freq_first" if self.is_first else "sn_freq", flops, nbytes,
                     lambda: _lib.check(lib.fnssl_sn_freq_forward(C.byref(fa), ops._stream())))
        z = torch.empty((nb, nt // pool, 16, H), dtype=torch.float32, device=x.device)
        work = torch.empty_like(y)
        ta.nb, ta.nt, ta.nf, ta.pool, ta.x, ta.work, ta.out = nb, nt, 16, pool, y.data_ptr(), work.data_ptr(), z.data_ptr()
        if state is not None:
            for st in state:
                if st.shape != (nb * 16, 19, 192) or st.dtype != torch.float32 or not st.is_contiguous() or st.device != x.device:
                    raise RuntimeError("SpatialNetLayer: Mamba state must be a contiguous float32 (nb*16, 19, 192) tensor on the device")
            ta.state[0], ta.state[1], ta.state_flags = state[0].data_ptr(), state[1].data_ptr(), 3
        else:
            ta.state[0], ta.state[1], ta.state_flags = None, None, 0
        ops._count(2)
        flops, nbytes = time_stage_work(nb * 16, nt, pool)
        ops.profiled("sn_time_T%d" % nt, flops, nbytes,
                     lambda: _lib.check(lib.fnssl_sn_time_forward(C.byref(ta), ops._stream())))
        return z

    def forward(self, x: Tensor, att_mask=None, chunkwise_recurrent: bool = True, rope: bool = True, state=None,
                inference: bool = False) -> Tuple[Tensor, None]:
        _require_eval(self)
        if self.is_first:
            raise RuntimeError("SpatialNetLayer(is_first=True) runs fused with the encoder; call OnlineSpatialNet.forward")
        B, F, T, H = x.shape
        if F != 16 or H != self.dim_hidden:
            raise RuntimeError(f"SpatialNetLayer: expected (B, 16, T, {self.dim_hidden}), got {tuple(x.shape)}")
        g = x.float().permute(0, 2, 1, 3).contiguous()
        return self._run(g, B, T, None, 1).permute(0, 2, 1, 3).contiguous(), None


class OnlineSpatialNet(nn.Module):
    """IPDnet2 network (IPDnet2.py:259-399).  forward(x: (B, dim_input, 256, T) f32) -> (B, T//5, 512, dim_output/4, 2)."""

    def __init__(self, dim_input: int, dim_output: int, num_layers: int, dim_squeeze: int, num_freqs: int,
                 encoder_kernel_size: int = 5, dim_hidden: int = 192, num_heads: int = 2, dropout=(0, 0, 0),
                 kernel_size=(5, 3), conv_groups=(8, 8), norms=("LN", "LN", "GN", "LN", "LN", "LN"), padding: str = 'zeros',
                 attention: str = 'mhsa(251)', chunkwise_recurrent: bool = True, rope=False,
                 fre_compression_ratio: int = 16, time_compression_ratio: int = 5, time_compression_layer: int = 0,
                 n_src: int = 2):
        super().__init__()
        if (dim_hidden, dim_squeeze, num_freqs, encoder_kernel_size, fre_compression_ratio, time_compression_ratio,
                time_compression_layer, kernel_size[0], conv_groups[0]) != (96, 8, 256, 5, 16, 5, 0, 5, 8):
            raise Exception("OnlineSpatialNet: fn_ssl_b200 is built for the configuration of run_IPDnet2.py:103-119 "
                            "(dim_hidden 96, dim_squeeze 8, 256 freqs, kernels 5, 8 groups, compression 16 / 5 at layer 0)")
        if dim_input > 16 or dim_input % 2:
            raise Exception("OnlineSpatialNet: dim_input = 2 * mics, at most 16")
        if dim_output % (2 * n_src):
            raise Exception("OnlineSpatialNet: dim_output must be 2 * n_src * (mics - 1)")
        self.dim_input, self.dim_output, self.num_layers, self.dim_hidden = dim_input, dim_output, num_layers, dim_hidden
        self.n_src = n_src                      # the literal 2 of the reference's output reshape (:363-364)
        self.time_compression_ratio, self.fre_compression_ratio = time_compression_ratio, fre_compression_ratio
        self.encoder = CausalConv1d(dim_input, dim_hidden, encoder_kernel_size, look_ahead=0)
        self.layers = nn.ModuleList([
            SpatialNetLayer(dim_hidden, dim_squeeze, num_freqs // 2 if l == 0 else num_freqs // fre_compression_ratio,
                            dropout, kernel_size, conv_groups, norms, padding, None, attention, is_first=(l == 0))
            for l in range(num_layers)])
        self.freq_inverse = FreqInverse(num_freqs, fre_compression_ratio, dim_hidden, dim_output)
        self.decoder = nn.Linear(dim_output, dim_output)
        self._head = None

    def _head_weights(self):
        key = (_versions(self.freq_inverse), _versions(self.decoder))
        if self._head is None or self._head[0] != key:
            with torch.no_grad():
                ws = [self.freq_inverse.trans2.weight[:, :, 0].t().detach().float().contiguous(),
                      self.freq_inverse.trans2.bias.detach().float().contiguous(),
                      self.decoder.weight.detach().float().contiguous(), self.decoder.bias.detach().float().contiguous()]
            self._head = (key, ws)
        return self._head[1]

    @ops.on_tensor_device
    def forward_grid(self, g0: Tensor, t_begin: int = 0, states=None) -> Tensor:
        """g0: feature grid (B, T, 256, ld) f32, ld = dim_input rounded up to 4 (zero padded).
        Streaming (fn_ssl_b200.IPDnet2.IPDnet2Stream): frames [0, t_begin) are encoder history, `states` holds one pair of
        Mamba state buffers per layer."""
        _require_eval(self)
        ops._need_cuda(g0)
        B, T, F, ld = g0.shape
        if F != 256 or g0.dtype != torch.float32 or ld < self.dim_input or ld % 4:
            raise RuntimeError(f"OnlineSpatialNet: expected an f32 feature grid (B, T, 256, ld>={self.dim_input}), got {tuple(g0.shape)}")
        r = self.time_compression_ratio
        if T - t_begin < r:
            raise RuntimeError(f"OnlineSpatialNet: needs at least {r} frames")
        st = states if states is not None else [None] * len(self.layers)
        x = self.layers[0]._run(g0.contiguous(), B, T, self.encoder, r, t_begin=t_begin, state=st[0])
        T5 = (T - t_begin) // r
        for i, layer in enumerate(list(self.layers)[1:]):
            x = layer._run(x, B, T5, None, 1, state=st[i + 1])
        tw, tb, dw, db = self._head_weights()
        K2 = self.dim_output // (2 * self.n_src)
        out = torch.empty((B, T5, 2 * F, K2, self.n_src), dtype=torch.float32, device=g0.device)
        ops._count(1)
        _lib.check(_lib.load().fnssl_sn_head_forward(x.data_ptr(), B, T5, 16, self.dim_hidden, tw.data_ptr(), tb.data_ptr(),
                                                     dw.data_ptr(), db.data_ptr(), self.dim_output,
                                                     self.fre_compression_ratio, self.n_src, out.data_ptr(), ops._stream()))
        return out

    def forward(self, x: Tensor, inference: bool = False, return_attn_score: bool = False) -> Tensor:
        _require_eval(self)
        if return_attn_score:
            raise RuntimeError("OnlineSpatialNet: Mamba layers have no attention scores")
        if x.dim() != 4 or x.shape[1] != self.dim_input or x.shape[2] != 256:
            raise RuntimeError(f"OnlineSpatialNet: expected (B, {self.dim_input}, 256, T), got {tuple(x.shape)}")
        return self.forward_grid(ops.cfirst_to_grid(x, torch.float32))


class IPDnet2_lightning(nn.Module):
    """`arch.`-prefixed wrapper so the reference's Lightning checkpoint loads (run_IPDnet2.py:103-119 builds exactly this)."""

    def __init__(self, dim_input: int = 10, dim_output: int = 16, num_layers: int = 8):
        super().__init__()
        self.arch = OnlineSpatialNet(dim_input=dim_input, dim_output=dim_output, num_layers=num_layers, dim_hidden=96,
                                     num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8), dim_squeeze=8, num_freqs=256,
                                     attention='mamba(16,4)', rope=False, time_compression_layer=0,
                                     fre_compression_ratio=16, time_compression_ratio=5)

    def forward(self, x):
        return self.arch(x)


IPDNET2_WIN, IPDNET2_HOP = 512, 320          # run_IPDnet2.py:91-93 (win_shift_ratio 0.625)


@ops.on_tensor_device
def reflect_padded(signal: Tensor) -> Tensor:
    """torch.stft(center=True)'s reflect padding by 256 samples on both sides (IPDnet2/Module.py:46-64)."""
    ops._need_cuda(signal)
    lib = _lib.load()
    x = signal.contiguous().float()
    nb, n, nch = x.shape
    pad = IPDNET2_WIN // 2
    xp = torch.empty((nb, n + 2 * pad, nch), dtype=torch.float32, device=x.device)
    ops._count(1)
    _lib.check(lib.fnssl_reflect_pad(x.data_ptr(), nb, n, nch, pad, xp.data_ptr(), ops._stream()))
    return xp


@ops.on_tensor_device
def stft_center(signal: Tensor, want_magsum: bool = False):
    """IPDnet2's STFT (IPDnet2/Module.py:46-64): torch.stft(center=True) = reflect pad 256 + framing, hop 320."""
    return ops.stft(reflect_padded(signal), IPDNET2_WIN, IPDNET2_HOP, IPDNET2_WIN, want_magsum=want_magsum)


def data_preprocess_ipdnet2(mic_sig_batch: Tensor, eps: float = 1e-6, sample_length: int = 249) -> List[Tensor]:
    """[(nb, 2*nch, 256, nt) f32]  (run_IPDnet2.py:277-288)."""
    spec, magsum = stft_center(mic_sig_batch, want_magsum=True)
    _, _, cf = ops.features(spec, magsum, 'ALL', ops.NORM_FORGETTING, sample_length, eps, torch.float32, want_cfirst=True)
    return [cf]


class IPDnet2Pipeline(nn.Module):
    """signal (nb, nsample, nch) f32 -> OnlineSpatialNet output (nb, nt//5, 512, nch-1, 2), nt = nsample//320 + 1."""

    def __init__(self, arch: OnlineSpatialNet, eps: float = 1e-6, sample_length: int = 249):
        super().__init__()
        self.arch, self.eps, self.sample_length = arch, eps, sample_length

    @torch.no_grad()
    def forward(self, signal: Tensor) -> Tensor:
        # fused front end (ops.stft_features): reflect pad, then FFT -> sum|X| -> normaliser -> FFT again -> feature grid;
        # the (B, 257, T, M) complex spectrum is never written to HBM
        with torch.cuda.device(signal.device):
            g0, _ = ops.stft_features(reflect_padded(signal), 'ALL', ops.NORM_FORGETTING, self.sample_length, self.eps,
                                      torch.float32, IPDNET2_WIN, IPDNET2_HOP, IPDNET2_WIN)
        return self.arch.forward_grid(g0)


class IPDnet2Stream:
    """Chunked (streaming) inference for OnlineSpatialNet -- the model is causal by construction (causal encoder, Mamba
    scans, per-frame frequency modules), so the whole-clip numbers are reproduced chunk by chunk with this state carried:

      * the sample overlap of the center=True STFT (512/320; the left reflect padding is built once from samples 1..256)
      * the forgetting-norm recursion mu_{t-1} and the absolute frame index        (utils_.py, sample_length 249)
      * the last 4 feature frames (history of the kernel-5 causal encoder)         (IPDnet2.py:66-76)
      * per layer and Mamba block: the selective-scan state and the depthwise conv's last 3 frames
        (what mamba_ssm keeps in InferenceParams, IPDnet2.py:170-177)

    Chunks are cut at multiples of 5 frames (the time pooling after layer 0); `push` accepts any number of samples and
    returns the newly completed output frames.  A frame is emitted once its whole 512-sample window has arrived, so the
    last frame(s) of a finite clip -- which the whole-clip STFT completes with right reflect padding -- are never emitted.

        stream = IPDnet2Stream(arch, nb=4)
        out = stream.push(block)          # None, or (nb, k, 512, M-1, 2) for the k newly completed output frames
    """

    def __init__(self, arch: OnlineSpatialNet, nb: int, eps: float = 1e-6, sample_length: int = 249,
                 device: Optional[torch.device] = None):
        if arch.training:
            raise RuntimeError("IPDnet2Stream: call arch.eval() first")
        self.arch, self.nb, self.nch = arch, nb, arch.dim_input // 2
        self.eps, self.sample_length = eps, sample_length
        self.device = torch.device(device) if device is not None else next(arch.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("IPDnet2Stream: the model must live on a CUDA device (there is no CPU path)")
        self.reset()

    def reset(self) -> None:
        dev = self.device
        self.frames_done = 0
        self._started = False
        self._pending = torch.empty((self.nb, 0, self.nch), dtype=torch.float32, device=dev)
        self._mu = torch.zeros((self.nb,), dtype=torch.float32, device=dev)
        self._hist = None
        self._states = [(torch.zeros((self.nb * 16, 19, 192), dtype=torch.float32, device=dev),
                         torch.zeros((self.nb * 16, 19, 192), dtype=torch.float32, device=dev)) for _ in self.arch.layers]

    @property
    def pending_samples(self) -> int:
        return self._pending.shape[1]

    @torch.no_grad()
    def push(self, samples: Tensor) -> Optional[Tensor]:
        if samples.dim() != 3 or samples.shape[0] != self.nb or samples.shape[2] != self.nch:
            raise RuntimeError(f"IPDnet2Stream.push: expected ({self.nb}, n, {self.nch}), got {tuple(samples.shape)}")
        if samples.device != self.device:
            raise RuntimeError("IPDnet2Stream.push: samples must be on the stream's CUDA device")
        buf = torch.cat((self._pending, samples.float()), dim=1) if self._pending.shape[1] else samples.float()
        pad = IPDNET2_WIN // 2
        if not self._started:
            if buf.shape[1] <= pad:
                self._pending = buf.contiguous()
                return None
            buf = torch.cat((buf[:, 1:pad + 1].flip(1), buf), dim=1)      # torch.stft(center=True)'s left reflect padding
            self._started = True
        n = buf.shape[1]
        frames = (n - IPDNET2_WIN) // IPDNET2_HOP + 1 if n >= IPDNET2_WIN else 0
        r = self.arch.time_compression_ratio
        k = frames // r * r
        if k == 0:
            self._pending = buf.contiguous()
            return None
        used = buf[:, :IPDNET2_HOP * (k - 1) + IPDNET2_WIN].contiguous()
        self._pending = buf[:, IPDNET2_HOP * k:].contiguous()
        spec, magsum = ops.stft(used, IPDNET2_WIN, IPDNET2_HOP, IPDNET2_WIN, want_magsum=True)
        mu = ops.norm_stream(magsum, 'ALL', self.sample_length, self.frames_done, self._mu)
        g, _, _ = ops.features(spec, None, 'ALL', ops.NORM_GIVEN, self.sample_length, self.eps, torch.float32, mu=mu)
        if self._hist is None:        # zero history == the causal encoder's zero padding at the start of a clip
            self._hist = torch.zeros((self.nb, 4, g.shape[2], g.shape[3]), dtype=torch.float32, device=self.device)
        gh = torch.cat((self._hist, g), dim=1)
        self._hist = gh[:, -4:].contiguous()
        out = self.arch.forward_grid(gh, t_begin=4, states=self._states)
        self.frames_done += k
        return out
