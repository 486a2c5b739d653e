"""CPU oracle for the FN-SSL / IPDnet forward hot path.

THIS FILE IS TEST INFRASTRUCTURE.  It is a functional (state_dict in, tensors out) CPU
restatement of the reference algorithm, used only by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the *checker* and the
*reported CPU baseline*.  Nothing under ``fn_ssl_b200/`` may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
modules from /root/reference (in the build container), runs them on seeded inputs / seeded
default-init weights and stores their outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

All ``path:line`` citations are relative to the reference checkout (Audio-WestlakeU/FN-SSL).
Arithmetic is fp32 on CPU, exactly what the reference executes (torch.stft -> rfft,
LSTM cell = sigmoid/tanh gates in PyTorch's i,f,g,o order, Conv2d, AvgPool2d, Linear).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# Front end
# --------------------------------------------------------------------------------------


def stft_num_frames(nsample: int, win_len: int = 512, hop: int = 256) -> int:
    """nt = floor((nsample - win_len) / hop + 1)   -- FN-SSL/Module.py:56, IPDnet/Module.py:53."""
    return int(math.floor((nsample - win_len) / hop + 1))


def stft(signal: Tensor, win_len: int = 512, win_shift_ratio: float = 0.5, nfft: int = 512) -> Tensor:
    """(nb, nsample, nch) f32 -> (nb, nfft/2+1, nt, nch) complex64.

    Restates ``STFT.forward`` (FN-SSL/Lightning/Module.py:48-68): per channel
    ``torch.stft(center=False, periodic Hann, normalized=False, onesided)``.  Written here as
    explicit framing  x[t*hop : t*hop+win]  * hann  -> rfft, which SURVEY.md §8c measured to be
    bit-identical to torch.stft(center=False).
    """
    nb, nsample, nch = signal.shape
    hop = int(win_len * win_shift_ratio)
    nt = stft_num_frames(nsample, win_len, hop)
    window = torch.hann_window(win_len, dtype=signal.dtype)           # periodic (Module.py:61)
    x = signal.permute(0, 2, 1)                                       # (nb, nch, nsample)
    frames = x.unfold(-1, win_len, hop)[:, :, :nt, :] * window        # (nb, nch, nt, win)
    spec = torch.fft.rfft(frames, n=nfft, dim=-1)                     # (nb, nch, nt, nf)
    return spec.permute(0, 3, 2, 1).contiguous()                      # (nb, nf, nt, nch)


def add_ch_to_batch(data: Tensor, ch_mode: str = "MM") -> Tensor:
    """(nb, nch, nf, nt) -> (nb*P, 2, nf, nt); ``AddChToBatch.forward`` FN-SSL/Lightning/Module.py:384-405.

    'M' : rows (ref=0, m) for m = 1..nch-1;  'MM': all pairs i<j in lexicographic order."""
    nb, nch = data.shape[:2]
    rows = []
    for b in range(nb):
        if ch_mode == "M":
            pairs = [(0, m) for m in range(1, nch)]
        elif ch_mode == "MM":
            pairs = [(i, j) for i in range(nch - 1) for j in range(i + 1, nch)]
        else:
            raise Exception("ch_mode unrecognised")
        for i, j in pairs:
            rows.append(torch.stack((data[b, i], data[b, j]), 0))
    return torch.stack(rows, 0).contiguous()


def forgetting_norm(mag: Tensor, sample_length: int = 298) -> Tensor:
    """(R, C, nf, nt) magnitudes -> (R, 1, 1, nt) recursive mean; FN-SSL/Lightning/utils_.py:9-55.

    m_t = mean over (C*nf) of mag[..., t];  mu_t = a_t*mu_{t-1} + (1-a_t)*m_t with
    a_t = min((t-1)/(t+1), (L-1)/(L+1)) for t < L else (L-1)/(L+1);  mu_{-1} = 0
    (so a_0 = -1 -> mu_0 = 2*m_0, a_1 = 0 -> mu_1 = m_1)."""
    assert mag.ndim == 4
    R, C, nf, nt = mag.shape
    flat = mag.reshape(R, C * nf, nt)
    alpha = (sample_length - 1) / (sample_length + 1)
    mu = 0
    out = []
    for t in range(nt):
        if t < sample_length:
            a = torch.min(torch.tensor([(t - 1) / (t + 1), alpha]))   # fp32 like the reference (:31)
            mu = a * mu + (1 - a) * torch.mean(flat[:, :, t], dim=1).reshape(R, 1)
        else:
            mu = alpha * mu + (1 - alpha) * torch.mean(flat[:, :, t], dim=1).reshape(R, 1)
        out.append(mu)
    return torch.stack(out, dim=-1).reshape(R, 1, 1, nt)


def preprocess_fnssl(signal: Tensor, ch_mode: str = "MM", eps: float = 1e-6,
                     sample_length: int = 298) -> Tensor:
    """(nb, nsample, nch) -> (nb*P, 4, 256, nt) network input.

    ``data_preprocess`` FN-SSL/Lightning/main.py:200-225: STFT -> (nb,nch,nf,nt) -> pair re-batch ->
    |.| -> forgetting_norm -> re/(mu+eps), im/(mu+eps) -> cat dim=1 -> bins 1..256."""
    spec = stft(signal).permute(0, 3, 1, 2)
    reb = add_ch_to_batch(spec, ch_mode)
    mu = forgetting_norm(torch.abs(reb), sample_length)
    re = torch.real(reb) / (mu + eps)
    im = torch.imag(reb) / (mu + eps)
    return torch.cat((re, im), dim=1)[:, :, 1:257, :].contiguous()


def preprocess_ipdnet(signal: Tensor, eps: float = 1e-6, sample_length: int = 280,
                      offline: bool = False) -> Tensor:
    """(nb, nsample, nch) -> (nb, 2*nch, 256, nt).

    Online: IPDnet/runIPDnetOn.py:240-254 (forgetting_norm, sample_length=280, all mics as channels).
    Offline: IPDnet/runIPDnetOff.py:248-251 (one utterance-global mean of |X| per batch row)."""
    spec = stft(signal).permute(0, 3, 1, 2)
    mag = torch.abs(spec)
    if offline:
        mu = torch.mean(mag.reshape(mag.shape[0], -1), dim=1)[:, None, None, None]
    else:
        mu = forgetting_norm(mag, sample_length)
    re = torch.real(spec) / (mu + eps)
    im = torch.imag(spec) / (mu + eps)
    return torch.cat((re, im), dim=1)[:, :, 1:257, :].contiguous()


# --------------------------------------------------------------------------------------
# LSTM (PyTorch nn.LSTM semantics, 1 layer, batch_first)
# --------------------------------------------------------------------------------------


def _lstm_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                    reverse: bool) -> Tensor:
    """Explicit time loop.  Gate order i,f,g,o; two bias vectors; zero initial state
    (nn.LSTM as instantiated at FN-SSL/Lightning/Model.py:25-29, IPDnet/FixedAarryIPDnet.py:24-28)."""
    N, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    out = x.new_empty(N, L, H)
    gx = x @ w_ih.t() + (b_ih + b_hh)
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        g = gx[:, t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def lstm(x: Tensor, sd: StateDict, prefix: str, fast: bool = False) -> Tensor:
    """(N, L, I) -> (N, L, dirs*H) with parameters ``prefix + weight_ih_l0[_reverse]`` etc.

    ``fast=True`` evaluates the same recurrence through ATen's fused CPU LSTM (``torch._VF.lstm``,
    the kernel the reference's nn.LSTM dispatches to) -- used for the timed CPU baseline and
    cross-checked against the explicit loop in tests/test_oracle_golden.py."""
    bidir = (prefix + "weight_ih_l0_reverse") in sd
    sufs = ["", "_reverse"] if bidir else [""]
    if fast:
        flat = []
        for s in sufs:
            flat += [sd[prefix + "weight_ih_l0" + s], sd[prefix + "weight_hh_l0" + s],
                     sd[prefix + "bias_ih_l0" + s], sd[prefix + "bias_hh_l0" + s]]
        H = flat[1].shape[1]
        zeros = x.new_zeros(len(sufs), x.shape[0], H)
        out, _, _ = torch._VF.lstm(x, (zeros, zeros), flat, True, 1, 0.0, False, bidir, True)
        return out
    outs = []
    for s in sufs:
        outs.append(_lstm_direction(x, sd[prefix + "weight_ih_l0" + s], sd[prefix + "weight_hh_l0" + s],
                                    sd[prefix + "bias_ih_l0" + s], sd[prefix + "bias_hh_l0" + s],
                                    reverse=(s == "_reverse")))
    return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------------------
# FN-SSL
# --------------------------------------------------------------------------------------


def fnssl_block(x: Tensor, sd: StateDict, prefix: str, is_first: bool,
                nb_skip: Optional[Tensor] = None, fb_skip: Optional[Tensor] = None,
                fast: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """``FNblock.forward`` FN-SSL/Lightning/Model.py:31-50 (eval mode: dropout = identity).

    x: (nb, nt, nf, nc).  Note :34 -- the narrow-band skip is ALWAYS recomputed from the block
    input, the ``nb_skip`` argument is ignored."""
    nb, nt, nf, nc = x.shape
    nb_skip = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)                   # :34
    x = x.reshape(nb * nt, nf, -1)
    if not is_first:
        x = x + fb_skip                                                        # :36-37
    x = lstm(x, sd, prefix + "fullLstm.", fast)                                # :38
    fb_skip = x
    x = x.view(nb, nt, nf, -1).permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)    # :41
    x = torch.cat((x, nb_skip), dim=-1) if is_first else x + nb_skip           # :42-45
    x = lstm(x, sd, prefix + "narrLstm.", fast)                                # :46
    nb_skip = x
    x = x.view(nb, nf, nt, -1).permute(0, 2, 1, 3)                             # :49
    return x, fb_skip, nb_skip


def fnssl_forward(x: Tensor, sd: StateDict, prefix: str = "", fast: bool = False) -> Tensor:
    """``FN_SSL.forward`` FN-SSL/Lightning/Model.py:72-90.  x: (nb, 4, nf, nt) ->
    (nb, nt//12, 2*nf) or, if ``ipd2doa.weight`` is in the state dict, (nb, nt//12, 180)."""
    x = x.permute(0, 3, 2, 1)
    nb, nt, nf, nc = x.shape
    x, fb, nbs = fnssl_block(x, sd, prefix + "block_1.", True, fast=fast)
    x, fb, nbs = fnssl_block(x, sd, prefix + "block_2.", False, nbs, fb, fast)
    x, fb, nbs = fnssl_block(x, sd, prefix + "block_3.", False, nbs, fb, fast)
    x = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)                         # :79
    ipd = F.avg_pool2d(x, kernel_size=(12, 1))                                 # :80  (T -> T//12)
    ipd = torch.tanh(F.linear(ipd, sd[prefix + "emb2ipd.weight"], sd[prefix + "emb2ipd.bias"]))
    nt2 = ipd.shape[1]
    ipd = ipd.view(nb, nf, nt2, -1).permute(0, 2, 1, 3)
    result = torch.cat((ipd[:, :, :, 0], ipd[:, :, :, 1]), dim=2)              # :85-87
    if (prefix + "ipd2doa.weight") in sd:
        result = F.linear(result, sd[prefix + "ipd2doa.weight"], sd[prefix + "ipd2doa.bias"])
    return result


# --------------------------------------------------------------------------------------
# IPDnet (fixed array)
# --------------------------------------------------------------------------------------


def ipdnet_block(x: Tensor, sd: StateDict, prefix: str, fb_skip: Tensor, nb_skip: Tensor,
                 fast: bool = False) -> Tensor:
    """IPDnet ``FNblock.forward`` IPDnet/FixedAarryIPDnet.py:29-40 (skips are *concatenated* raw input)."""
    nb, nt, nf, nc = x.shape
    x = x.reshape(nb * nt, nf, -1)
    x = lstm(x, sd, prefix + "fullLstm.", fast)
    x = torch.cat((x, fb_skip), dim=-1)
    x = x.view(nb, nt, nf, -1).permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
    x = lstm(x, sd, prefix + "narrLstm.", fast)
    x = torch.cat((x, nb_skip), dim=-1)
    return x.view(nb, nf, nt, -1).permute(0, 2, 1, 3)


def causcnn(x: Tensor, sd: StateDict, prefix: str = "conv.", pad: Tuple[int, int] = (1, 2)) -> Tensor:
    """``CausCnnBlock.forward`` IPDnet/FixedAarryIPDnet.py:61-73.  x: (nb, C, F, T) -> (nb, out, F, T//12)."""
    out = F.conv2d(x, sd[prefix + "conv1.weight"], None, stride=(1, 1), padding=pad)
    out = torch.relu(out)[:, :, :, :-pad[1]]
    out = F.avg_pool2d(out, kernel_size=(1, 3))
    out = F.conv2d(out, sd[prefix + "conv2.weight"], None, stride=(1, 1), padding=pad)
    out = torch.relu(out)[:, :, :, :-pad[1]]
    out = F.avg_pool2d(out, kernel_size=(1, 4))
    out = F.conv2d(out, sd[prefix + "conv3.weight"], None, stride=(1, 1), padding=pad)
    out = out[:, :, :, :-pad[1]]
    return torch.tanh(out)


def _split_segments(x: Tensor, seg_len: int) -> Tensor:
    """IPDnet/utils_.py:152-167: zero-pad nt to a multiple of seg_len, reshape to (nb, nseg, seg, nf, nc)."""
    nb, nt, nf, nc = x.shape
    pad_len = (seg_len - (nt % seg_len)) % seg_len
    if pad_len > 0:
        x = torch.cat([x, x.new_zeros(nb, pad_len, nf, nc)], dim=1)
    return x.reshape(nb, x.shape[1] // seg_len, seg_len, nf, nc)


def ipdnet_forward(x: Tensor, sd: StateDict, is_online: bool = True, offline_inference: bool = False,
                   n_seg: int = 312, fast: bool = False) -> Tensor:
    """``IPDnet.forward`` IPDnet/FixedAarryIPDnet.py:91-120.  x: (nb, 2M, nf, nt) ->
    (nb, nt//12, 2*nf, M-1, 2)."""
    x = x.permute(0, 3, 2, 1)
    nb, nt, nf, nc = x.shape
    ou_frame = nt // 12
    chunked = (not is_online) and offline_inference
    if chunked:
        x = _split_segments(x, n_seg)
        nb, nseg, seg_nt, nf, nc = x.shape
        x = x.reshape(nb * nseg, seg_nt, nf, nc)
        nb, nt, nf, nc = x.shape
    fb_skip = x.reshape(nb * nt, nf, nc)
    nb_skip = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, nc)
    x = ipdnet_block(x, sd, "block_1.", fb_skip, nb_skip, fast)
    x = ipdnet_block(x, sd, "block_2.", fb_skip, nb_skip, fast)
    nb, nt, nf, nc = x.shape
    x = x.permute(0, 3, 2, 1)
    nt2 = nt // 12
    x = causcnn(x, sd).permute(0, 3, 2, 1).reshape(nb, nt2, nf, 2, -1).permute(0, 1, 3, 2, 4)
    if chunked:
        x = x.reshape(nb // nseg, nt2 * nseg, 2, nf * 2, -1).permute(0, 1, 3, 4, 2)
        return x[:, :ou_frame, :, :, :]
    return x.reshape(nb, nt2, 2, nf * 2, -1).permute(0, 1, 3, 4, 2)


# --------------------------------------------------------------------------------------
# Seeded default-init weights (same RNG stream as constructing the reference module)
# --------------------------------------------------------------------------------------


def _lstm_init(sd: StateDict, prefix: str, inp: int, H: int, bidir: bool) -> None:
    """nn.LSTM.reset_parameters: U(-1/sqrt(H), 1/sqrt(H)) for every tensor in registration order."""
    k = 1.0 / math.sqrt(H)
    for suf in (["", "_reverse"] if bidir else [""]):
        sd[prefix + "weight_ih_l0" + suf] = torch.empty(4 * H, inp).uniform_(-k, k)
        sd[prefix + "weight_hh_l0" + suf] = torch.empty(4 * H, H).uniform_(-k, k)
        sd[prefix + "bias_ih_l0" + suf] = torch.empty(4 * H).uniform_(-k, k)
        sd[prefix + "bias_hh_l0" + suf] = torch.empty(4 * H).uniform_(-k, k)


def _linear_init(sd: StateDict, prefix: str, inp: int, out: int) -> None:
    """nn.Linear.reset_parameters (kaiming_uniform(a=sqrt 5) == U(-1/sqrt(in), 1/sqrt(in)) for both)."""
    w = torch.empty(out, inp)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    k = 1.0 / math.sqrt(inp)
    sd[prefix + "weight"] = w
    sd[prefix + "bias"] = torch.empty(out).uniform_(-k, k)


def _conv_init(sd: StateDict, name: str, cin: int, cout: int) -> None:
    w = torch.empty(cout, cin, 3, 3)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    sd[name] = w


def seeded_fnssl_state_dict(seed: int = 0, is_online: bool = True, is_doa: bool = False) -> StateDict:
    """State dict equal to ``torch.manual_seed(seed); FN_SSL(is_online=..., is_doa=...).state_dict()``
    (FN-SSL/Lightning/Model.py:56-71); equality is asserted by tests/golden/make_golden.py."""
    torch.manual_seed(seed)
    sd: StateDict = {}
    nh = 256 if is_online else 128
    _lstm_init(sd, "block_1.fullLstm.", 4, 128, True)
    _lstm_init(sd, "block_1.narrLstm.", 260, nh, not is_online)
    for b in ("block_2.", "block_3."):
        _lstm_init(sd, b + "fullLstm.", 256, 128, True)
        _lstm_init(sd, b + "narrLstm.", 256, nh, not is_online)
    _linear_init(sd, "emb2ipd.", 256, 2)
    if is_doa:
        _linear_init(sd, "ipd2doa.", 512, 180)
    return sd


def seeded_ipdnet_state_dict(seed: int = 0, input_size: int = 4, hidden_size: int = 128,
                             max_track: int = 2, is_online: bool = True) -> StateDict:
    """Equal to ``torch.manual_seed(seed); IPDnet(...).state_dict()`` (IPDnet/FixedAarryIPDnet.py:80-90)."""
    torch.manual_seed(seed)
    sd: StateDict = {}
    fh = hidden_size // 2
    nh = hidden_size if is_online else hidden_size // 2
    _lstm_init(sd, "block_1.fullLstm.", input_size, fh, True)
    _lstm_init(sd, "block_1.narrLstm.", 2 * fh + input_size, nh, not is_online)
    _lstm_init(sd, "block_2.fullLstm.", hidden_size + input_size, fh, True)
    _lstm_init(sd, "block_2.narrLstm.", 2 * fh + input_size, nh, not is_online)
    cin = hidden_size + input_size
    cout = 2 * ((input_size // 2) - 1) * max_track
    _conv_init(sd, "conv.conv1.weight", cin, 128)
    _conv_init(sd, "conv.conv2.weight", 128, 128)
    _conv_init(sd, "conv.conv3.weight", 128, cout)
    return sd


def white_noise(nb: int, nsample: int, nch: int, seed: int = 1234) -> Tensor:
    """Synthetic input of SURVEY.md §8d: torch.randn(B, nsample, M) from Generator(seed)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(nb, nsample, nch, generator=g, dtype=torch.float32)


# --------------------------------------------------------------------------------------
# "Next" row #1: DP-IPD templates and IPD -> DOA decoding
# --------------------------------------------------------------------------------------


def _pair_select(data, ch_mode: str):
    """``DPIPD.data_adjust`` FN-SSL/Lightning/Module.py:500-514: (..., nmic, nmic) -> (..., npairs)."""
    import numpy as np
    nmic = data.shape[-1]
    if ch_mode == "M":
        return data[..., 0, 1:]
    if ch_mode == "MM":
        cols = [data[..., i, j] for i in range(nmic - 1) for j in range(i + 1, nmic)]
        return np.stack(cols, axis=-1).astype(np.complex64)
    raise Exception("Microphone channel mode unrecognised")


def dpipd_template(ndoa_candidate, mic_location, nf: int = 257, fre_max: float = 8000, ch_mode: str = "M",
                   speed: float = 343.0):
    """Far-field direct-path IPD templates, ``DPIPD.__init__`` FN-SSL/Lightning/Module.py:428-462.
    Returns (template (nele, nazi, nf, npairs) complex, [ele_candidate, azi_candidate])."""
    import numpy as np
    nele, nazi = ndoa_candidate
    ele = np.linspace(0, np.pi, nele)
    azi = np.linspace(-np.pi, np.pi, nazi)
    nmic = mic_location.shape[-2]
    r = np.stack([np.outer(np.sin(ele), np.cos(azi)), np.outer(np.sin(ele), np.sin(azi)),
                  np.tile(np.cos(ele), [nazi, 1]).transpose()], axis=2)                    # (nele, nazi, 3) unit vectors
    fre = np.linspace(0.0, fre_max, nf)
    ipd = np.empty((nele, nazi, nf, nmic, nmic))
    for m1 in range(nmic):
        for m2 in range(nmic):
            itd = np.dot(r, mic_location[m2, :] - mic_location[m1, :]) / speed            # :449
            ipd[:, :, :, m1, m2] = -2 * np.pi * fre[None, None, :] * itd[:, :, None]      # :450-451
    return _pair_select(np.exp(1j * ipd), ch_mode), [ele, azi]


def doa_templates_for_decode(template, fre_range_used=range(1, 257)):
    """``PredDOA.predgt2DOA`` FN-SSL/Lightning/Module.py:701-716: [re | im] over the used bins, elevation fixed to the
    horizontal plane, azimuths 0..pi.  (nele,nazi,nf,P) complex -> ((1, nazi_half, 2*len(bins), P) f32, [ele, azi])."""
    import numpy as np
    t = np.concatenate((template.real[:, :, fre_range_used, :], template.imag[:, :, fre_range_used, :]), axis=2).astype(np.float32)
    nele, nazi = t.shape[:2]
    t = t[int((nele - 1) / 2):int((nele - 1) / 2) + 1, int((nazi - 1) / 2):nazi, :, :]
    return t, [np.linspace(np.pi / 2, np.pi / 2, 1), np.linspace(0, np.pi, 37)]


def source_detect_localize_idl(pred_ipd: Tensor, template: Tensor, doa_candidate, max_num_sources: int = 1,
                               source_num_mode: str = "kNum"):
    """Iterative detection/localisation, ``SourceDetectLocalize.forward`` (meth_mode 'IDL')
    FN-SSL/Lightning/Module.py:525-577.  pred_ipd (nb, nt, 2nf, P), template (nele, nazi, 2nf, P) ->
    (DOAs (nb,nt,2,ns) [ele, azi] radians, VADs (nb,nt,ns), spatial spectrum (nb,nt,nele,nazi))."""
    import numpy as np
    nb, nt, nf2, P = pred_ipd.shape
    nele, nazi = template.shape[:2]
    tmat = template.reshape(nele * nazi, nf2 * P)                       # (cand, K): same (2nf, P) flattening as :537
    cur = pred_ipd.reshape(nb * nt, nf2 * P).clone()
    scale = P * nf2 / 2
    doas = torch.zeros(nb * nt, 2, max_num_sources)
    vads = torch.zeros(nb * nt, max_num_sources)
    ss0 = None
    for sidx in range(max_num_sources):
        smap = cur @ tmat.t() / scale                                   # :553-556
        if ss0 is None:
            ss0 = smap.clone()
        idx = smap.argmax(dim=1)                                        # :558
        e_idx, a_idx = np.unravel_index(idx.numpy(), (nele, nazi))
        doas[:, 0, sidx] = torch.from_numpy(np.asarray(doa_candidate[0])[e_idx]).float()
        doas[:, 1, sidx] = torch.from_numpy(np.asarray(doa_candidate[1])[a_idx]).float()
        best = tmat[idx]                                                # (R, K)
        ratio = (best * cur).sum(1) / (best * best).sum(1)              # :571-574
        if source_num_mode == "kNum":
            vads[:, sidx] = 1
        elif source_num_mode == "unkNum":
            vads[:, sidx] = ratio
        cur = cur - ratio[:, None] * best                               # :575,:580
    return (doas.reshape(nb, nt, 2, max_num_sources), vads.reshape(nb, nt, max_num_sources),
            ss0.reshape(nb, nt, nele, nazi))
