"""Golden vectors for the IPDnet2 row, from the UNMODIFIED reference ``IPDnet2/IPDnet2.py``.

Run in the build container only:   python tests/golden/make_golden_ipdnet2.py   -> tests/golden/ipdnet2_golden.npz

Shims (nothing of the reference is edited):
* ``matplotlib`` / ``soundfile`` / ``webrtcvad`` / ``scipy.signal`` users: empty modules where absent (plotting / IO).
* ``mamba_ssm``: the third-party package the reference imports at IPDnet2.py:16 is not in this image and not under
  /root/reference.  A stand-in ``mamba_ssm.Mamba`` nn.Module with the parameter names / shapes of the reference's
  shipped checkpoint is injected; its forward is ``oracle.ipdnet2_oracle.mamba``.  Hence the golden outputs pin the
  reference's front end, encoder, cross-band / full-band modules, pooling, FreqInverse, decoder, output reshape and its
  USE of the Mamba block -- not the block's arithmetic.
* SECOND PIN of the block's arithmetic (round 2): the same unmodified reference model is run again with ``mamba_ssm.Mamba``
  := a thin subclass of Hugging Face transformers' ``MambaMixer`` (transformers 5.5, ``slow_forward`` -- HF's own port of the
  state-spaces/mamba algorithm, with mamba_ssm's parameter names, validated upstream against mamba_ssm checkpoints).  Those
  outputs are stored as ``net_*_out_hfmamba`` and a single-block fixture as ``mamba_hf_{x,y}``: an implementation of the block
  that this repository did not write.  (tests/test_gpu_ipdnet2.py additionally checks the scan against vLLM's CUDA port of
  mamba_ssm's selective_scan_fwd kernel on the GPU box.)
The script also loads the reference checkpoint ``IPDnet2/checkpoints/ipdnet2_small.ckpt`` strictly into the reference
model and checks the oracle against it (no fixture stored for it: 326 tensors / 5 MB).
"""
import contextlib
import io
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/IPDnet2"
sys.path.insert(0, ROOT)
from oracle import ipdnet2_oracle as orc2  # noqa: E402


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


class MambaBase(nn.Module):
    """What the reference sees as `mamba_ssm.Mamba` (it also does isinstance checks against it, IPDnet2.py:155,160):
    constructing it yields whichever implementation is selected at that moment."""
    impl = None

    def __new__(cls, *a, **kw):
        return super().__new__(MambaBase.impl if cls is MambaBase else cls)


class MambaStandIn(MambaBase):
    """Parameter container with mamba_ssm.Mamba's names; forward = the oracle's restatement."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, layer_idx=None, **kw):
        super().__init__()
        d_inner = expand * d_model
        dt_rank = math.ceil(d_model / 16)
        self.in_proj = nn.Linear(d_model, 2 * d_inner, bias=False)
        self.conv1d = nn.Conv1d(d_inner, d_inner, d_conv, groups=d_inner, padding=d_conv - 1, bias=True)
        self.x_proj = nn.Linear(d_inner, dt_rank + 2 * d_state, bias=False)
        self.dt_proj = nn.Linear(dt_rank, d_inner, bias=True)
        self.A_log = nn.Parameter(torch.zeros(d_inner, d_state))
        self.D = nn.Parameter(torch.ones(d_inner))
        self.out_proj = nn.Linear(d_inner, d_model, bias=False)

    def forward(self, x, inference_params=None):
        assert inference_params is None
        return orc2.mamba(x, dict(self.state_dict()), "")


def hf_mamba_class():
    """mamba_ssm.Mamba's constructor signature over transformers' MambaMixer (independent third-party implementation)."""
    from transformers.models.mamba.modeling_mamba import MambaConfig, MambaMixer

    class HFMamba(MambaBase, MambaMixer):
        def __init__(self, d_model, d_state=16, d_conv=4, expand=2, layer_idx=None, **kw):
            cfg = MambaConfig(hidden_size=d_model, state_size=d_state, conv_kernel=d_conv, expand=expand,
                              time_step_rank=math.ceil(d_model / 16), use_bias=False, use_conv_bias=True,
                              num_hidden_layers=1, vocab_size=8)
            MambaMixer.__init__(self, cfg, layer_idx or 0)

        def forward(self, x, inference_params=None):
            assert inference_params is None
            return self.slow_forward(x)

    return HFMamba


MambaBase.impl = MambaStandIn


def _import_reference():
    for m in ["matplotlib", "matplotlib.pyplot", "soundfile", "webrtcvad"]:
        sys.modules.setdefault(m, types.ModuleType(m))
    ms = types.ModuleType("mamba_ssm")
    ms.Mamba = MambaBase
    gen = types.ModuleType("mamba_ssm.utils.generation")
    gen.InferenceParams = object
    sys.modules["mamba_ssm"] = ms
    sys.modules["mamba_ssm.utils"] = types.ModuleType("mamba_ssm.utils")
    sys.modules["mamba_ssm.utils.generation"] = gen
    sys.path.insert(0, REF)
    import IPDnet2 as ref_net          # IPDnet2/IPDnet2.py
    import Module as ref_module        # IPDnet2/Module.py
    import utils_ as ref_utils         # IPDnet2/utils_.py
    return ref_net, ref_module, ref_utils


def _run_quiet(model, x):
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):   # SpatialNetLayer.forward prints a shape (:149)
        return model(x)


def main():
    hf_mamba_class()      # import transformers BEFORE the empty stand-in modules exist (its availability probes trip over them)
    ref_net, ref_module, ref_utils = _import_reference()
    out = {}
    # ---- front end: center=True STFT 512/320 + forgetting_norm(249) + re/im features (run_IPDnet2.py:277-288)
    sig = _randn((2, 320 * 24 + 101, 3), 21)
    stft = ref_module.STFT(win_len=512, win_shift_ratio=0.625, nfft=512)(sig)
    out["fe_stft_re"], out["fe_stft_im"] = stft.real.numpy(), stft.imag.numpy()
    st = stft.permute(0, 3, 1, 2)
    mu = ref_utils.forgetting_norm(torch.abs(st), sample_length=249)
    feat = torch.cat((torch.real(st) / (mu + 1e-6), torch.imag(st) / (mu + 1e-6)), dim=1)[:, :, range(1, 257), :]
    out["fe_feat"] = feat.numpy()
    o = orc2.stft_center(sig)
    assert o.shape == stft.shape and float((o - stft).abs().max()) <= 1e-5 * float(stft.abs().max())
    assert float((orc2.preprocess_ipdnet2(sig) - feat).abs().max()) <= 1e-5 * float(feat.abs().max())

    # ---- network: small (3 mics, 3 layers) and the reference's own configuration (5 mics, 8 layers), seeded weights
    for tag, cfg, xshape, seed in (
            ("small", dict(dim_input=6, dim_output=8, num_layers=3), (2, 6, 256, 27), 22),
            ("default", dict(dim_input=10, dim_output=16, num_layers=8), (1, 10, 256, 40), 23)):
        model = ref_net.OnlineSpatialNet(dim_hidden=96, num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8),
                                         norms=["LN", "LN", "GN", "LN", "LN", "LN"], dim_squeeze=8, num_freqs=256,
                                         attention='mamba(16,4)', rope=False, time_compression_layer=0,
                                         fre_compression_ratio=16, time_compression_ratio=5, **cfg).eval()
        sd = orc2.seeded_ipdnet2_state_dict(seed, **cfg)
        model.load_state_dict(sd, strict=True)
        x = _randn(xshape, seed + 100)
        y = _run_quiet(model, x)
        yo = orc2.ipdnet2_forward(x, sd)
        err = float((y - yo).abs().max()) / float(y.abs().max())
        print(f"[{tag}] reference out {tuple(y.shape)}  oracle rel-to-max err {err:.2e}")
        assert y.shape == yo.shape and err <= 2e-5
        out[f"net_{tag}_out"] = y.numpy()
        # the same unmodified reference model with Hugging Face's Mamba implementation inside
        MambaBase.impl = hf_mamba_class()
        try:
            model_hf = ref_net.OnlineSpatialNet(dim_hidden=96, num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8),
                                                norms=["LN", "LN", "GN", "LN", "LN", "LN"], dim_squeeze=8, num_freqs=256,
                                                attention='mamba(16,4)', rope=False, time_compression_layer=0,
                                                fre_compression_ratio=16, time_compression_ratio=5, **cfg).eval()
        finally:
            MambaBase.impl = MambaStandIn
        model_hf.load_state_dict(sd, strict=True)
        yh = _run_quiet(model_hf, x)
        errh = float((yh - yo).abs().max()) / float(yh.abs().max())
        print(f"[{tag}] reference + HF MambaMixer: oracle rel-to-max err {errh:.2e}; vs stand-in run {float((yh - y).abs().max()):.2e}")
        assert errh <= 2e-5
        out[f"net_{tag}_out_hfmamba"] = yh.numpy()

    # ---- the shipped checkpoint loads strictly and the oracle follows it
    ck = torch.load(os.path.join(REF, "checkpoints", "ipdnet2_small.ckpt"), map_location="cpu", weights_only=False)
    sd = {k[len("arch."):]: v.float() for k, v in ck["state_dict"].items() if k.startswith("arch.")}
    cfg = dict(dim_input=10, dim_output=16, num_layers=8)
    shapes = orc2.ipdnet2_param_shapes(**cfg)
    assert set(shapes) == set(sd) and all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
    model = ref_net.OnlineSpatialNet(dim_hidden=96, num_heads=4, dim_squeeze=8, num_freqs=256,
                                     attention='mamba(16,4)', rope=False, **cfg).eval()
    model.load_state_dict(sd, strict=True)
    x = _randn((1, 10, 256, 35), 31)
    y = _run_quiet(model, x)
    yo = orc2.ipdnet2_forward(x, sd)
    err = float((y - yo).abs().max()) / float(y.abs().max())
    print(f"[ckpt] reference out {tuple(y.shape)}  oracle rel-to-max err {err:.2e}")
    assert err <= 2e-5
    # ---- single Mamba block fixture from the HF implementation (shapes of the shipped checkpoint; mamba_ssm-style init of
    # A_log = log(1..d_state), dt bias ~ softplus^-1 of dt in [1e-3, 1e-1], everything else N(0, 0.2))
    blk = hf_mamba_class()(96).eval()
    g = torch.Generator().manual_seed(41)
    bsd = {k: 0.2 * torch.randn(v.shape, generator=g) for k, v in blk.state_dict().items()}
    bsd["A_log"] = torch.log(torch.arange(1, 17).float())[None].repeat(192, 1)
    dt = torch.exp(torch.rand(192, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
    bsd["dt_proj.bias"] = dt + torch.log(-torch.expm1(-dt))
    bsd["D"] = torch.ones(192)
    blk.load_state_dict(bsd)
    bx = _randn((3, 61, 96), 42)
    with torch.no_grad():
        by = blk(bx)
    bo = orc2.mamba(bx, {"m." + k: v for k, v in bsd.items()}, "m.")
    berr = float((by - bo).abs().max()) / float(by.abs().max())
    print(f"[mamba block] HF MambaMixer vs oracle.mamba rel-to-max err {berr:.2e}")
    assert berr <= 1e-5
    out["mamba_hf_x"], out["mamba_hf_y"] = bx.numpy(), by.numpy()
    for k, v in bsd.items():
        out["mamba_hf_sd." + k] = v.numpy()

    # ---- the shipped checkpoint through reference + HF Mamba as well
    MambaBase.impl = hf_mamba_class()
    try:
        model_hf = ref_net.OnlineSpatialNet(dim_hidden=96, num_heads=4, dim_squeeze=8, num_freqs=256,
                                            attention='mamba(16,4)', rope=False, **cfg).eval()
    finally:
        MambaBase.impl = MambaStandIn
    model_hf.load_state_dict(sd, strict=True)
    yh = _run_quiet(model_hf, x)
    errh = float((yh - yo).abs().max()) / float(yh.abs().max())
    print(f"[ckpt] reference + HF MambaMixer: oracle rel-to-max err {errh:.2e}")
    assert errh <= 2e-5
    np.savez_compressed(os.path.join(HERE, "ipdnet2_golden.npz"), **out)
    print("wrote ipdnet2_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
