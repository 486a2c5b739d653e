"""GPU parity tests: every kernel of the hot path, called through the C ABI (ctypes), against the CPU oracle
and the committed golden vectors of the unmodified reference.

Tolerances: max|y - y_ref| <= tol * max|y_ref| per tensor.
    STFT                : 1e-5 (values), frame indexing bit-exact (shape + per-frame alignment test)
    SIMT engine (fp32)  : 2e-5
    tcgen05 engine      : 1e-3  (fp16 operands, fp32 accumulate; BASELINE.json's bar)
"""
import numpy as np
import pytest
import torch

from oracle import fnssl_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def _relerr(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).float()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _engines():
    from fn_ssl_b200 import config
    return ["simt", "tcgen05"] if config.TC_AVAILABLE else ["simt"]


TOL = {"simt": 2e-5, "tcgen05": 1e-3}


# ---------------------------------------------------------------------------------------------
# front end
# ---------------------------------------------------------------------------------------------

def test_stft_matches_reference_golden(golden_fnssl):
    import fn_ssl_b200 as F
    sig = _randn((2, 512 + 256 * 30 + 77, 3), 11)
    spec = F.STFT(512, 0.5, 512)(sig.to(DEV))
    assert spec.shape == (2, 257, 31, 3) and spec.dtype == torch.complex64
    ref = torch.complex(torch.from_numpy(golden_fnssl["fe_stft_re"]), torch.from_numpy(golden_fnssl["fe_stft_im"]))
    assert _relerr(torch.view_as_real(spec), torch.view_as_real(ref)) <= 1e-5


@pytest.mark.parametrize("nsample,nch", [(512, 1), (512 + 255, 2), (512 + 256, 2), (64000, 2), (9001, 5), (20000, 8)])
def test_stft_shapes_and_edges(nsample, nch):
    import fn_ssl_b200 as F
    sig = _randn((2, nsample, nch), 5)
    spec = F.STFT(512, 0.5, 512)(sig.to(DEV))
    ref = orc.stft(sig)
    assert spec.shape == ref.shape                                  # nt = floor((n-512)/256 + 1): bit-exact framing
    assert _relerr(torch.view_as_real(spec), torch.view_as_real(ref)) <= 1e-5


def test_stft_frame_indexing_is_exact():
    """A unit impulse at sample n lands in exactly the frames t with t*256 <= n < t*256+512, weighted by hann[n - 256 t]."""
    import fn_ssl_b200 as F
    n = 1000
    sig = torch.zeros(1, 4096, 1)
    sig[0, n, 0] = 1.0
    spec = F.STFT(512, 0.5, 512)(sig.to(DEV)).cpu()
    mag = spec.abs()[0, :, :, 0]                                    # (257, nt)
    hann = torch.hann_window(512)
    for t in range(mag.shape[1]):
        inside = t * 256 <= n < t * 256 + 512
        expect = float(hann[n - 256 * t]) if inside else 0.0
        assert torch.allclose(mag[:, t], torch.full((257,), expect), atol=1e-6), t


def test_too_short_signal_raises():
    import fn_ssl_b200 as F
    with pytest.raises(RuntimeError, match="shorter"):
        F.STFT(512, 0.5, 512)(torch.zeros(1, 300, 2, device=DEV))


@pytest.mark.parametrize("mode", ["M", "MM"])
def test_fnssl_features_match_reference_golden(golden_fnssl, mode):
    import fn_ssl_b200 as F
    sig = _randn((2, 512 + 256 * 30 + 77, 3), 11)
    feat = F.data_preprocess_fnssl(sig.to(DEV), ch_mode=mode)[0]
    assert _relerr(feat, golden_fnssl[f"fe_feat_{mode}"]) <= 1e-5
    # the standalone normaliser, default and short sample_length (exercises the t >= L branch)
    spec = F.STFT(512, 0.5, 512)(sig.to(DEV)).permute(0, 3, 1, 2)
    reb = F.AddChToBatch(mode)(spec)
    assert _relerr(F.forgetting_norm(reb.abs()), golden_fnssl[f"fe_mu_{mode}"]) <= 1e-5
    assert _relerr(F.forgetting_norm(reb.abs(), 8), golden_fnssl[f"fe_mu8_{mode}"]) <= 1e-5


def test_ipdnet_features_match_reference_golden(golden_ipdnet):
    import fn_ssl_b200 as F
    sig = _randn((2, 512 + 256 * 20 + 5, 4), 21)
    assert _relerr(F.data_preprocess_ipdnet(sig.to(DEV))[0], golden_ipdnet["fe_feat_on"]) <= 1e-5
    assert _relerr(F.data_preprocess_ipdnet(sig.to(DEV), offline=True)[0], golden_ipdnet["fe_feat_off"]) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_feature_grid_layout(dtype):
    """Grid (R, nt, 256, ld) == the reference (R, C, 256, nt) tensor permuted; padding channels are zero."""
    from fn_ssl_b200 import ops
    sig = _randn((2, 6000, 2), 3).to(DEV)
    spec, magsum = ops.stft(sig, want_magsum=True)
    g, mu, cf = ops.features(spec, magsum, "MM", ops.NORM_FORGETTING, 298, 1e-6, dtype, want_cfirst=True)
    ref = orc.preprocess_fnssl(sig.cpu())
    assert _relerr(cf, ref) <= 1e-5
    tol = 1e-5 if dtype == torch.float32 else 1e-3
    assert _relerr(g[..., :4].float().permute(0, 3, 2, 1), ref) <= tol
    assert float(g[..., 4:].abs().max()) == 0.0 if g.shape[-1] > 4 else True
    back = ops.grid_to_cfirst(g, 4)
    assert _relerr(back, ref) <= tol
    g2 = ops.cfirst_to_grid(cf, dtype)
    assert torch.equal(g2, g)


@pytest.mark.parametrize("nch,pairing", [(2, "MM"), (3, "MM"), (3, "M"), (4, "ALL"), (8, "ALL"), (5, "ALL")])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_fused_stft_features_equals_two_step_front_end(nch, pairing, dtype):
    """ops.stft_features (the pipelines' front end: two FFT passes, the complex spectrum never written) is bit-identical to
    ops.stft + ops.features for every pairing / normaliser, including a ragged last frame tile and hop 320 (IPDnet2)."""
    from fn_ssl_b200 import ops
    sig = _randn((2, 512 + 256 * 37 + 19, nch), 31).to(DEV)
    for norm, hop in ((ops.NORM_FORGETTING, 256), (ops.NORM_GLOBAL, 256), (ops.NORM_NONE, 256), (ops.NORM_FORGETTING, 320)):
        spec, magsum = ops.stft(sig, 512, hop, 512, want_magsum=True)
        g_ref, mu_ref, _ = ops.features(spec, magsum, pairing, norm, 31, 1e-6, dtype)      # sample_length 31 < nt: both branches
        g, mu = ops.stft_features(sig, pairing, norm, 31, 1e-6, dtype, 512, hop, 512)
        assert g.shape == g_ref.shape and torch.equal(g, g_ref), (norm, hop)
        if norm != ops.NORM_NONE:
            assert torch.equal(mu, mu_ref)
    # the reference oracle on top (FN-SSL pairing only: the oracle's preprocess is the 'MM' path)
    if pairing == "MM":
        g, _ = ops.stft_features(sig, "MM", ops.NORM_FORGETTING, 298, 1e-6, torch.float32)
        ref = orc.preprocess_fnssl(sig.cpu())
        assert _relerr(g[..., :4].permute(0, 3, 2, 1), ref) <= 1e-5


# ---------------------------------------------------------------------------------------------
# LSTM layer, both axes, both engines
# ---------------------------------------------------------------------------------------------

def _tc4_only(monkeypatch):
    """Route every tensor-core layer to lstm_tc4.cu (the CTA-pair kernels lstm_tc5.cu / lstm_tc6.cu off)."""
    monkeypatch.setenv("FNSSL_TC_PAIR", "0")
    monkeypatch.setenv("FNSSL_TC_PAIR256", "0")


def _lstm_case(engine, axis, nb, nt, nf, c0, c1, H, bidir, use_addend, seed=0, inplace=False):
    from fn_ssl_b200 import config, ops
    from fn_ssl_b200.packing import LSTMParams, run_lstm
    dt = config.grid_dtype(engine)
    torch.manual_seed(seed)
    p = LSTMParams(c0 + c1, H, bidirectional=bidir).to(DEV)
    x0 = _randn((nb, nt, nf, c0), seed + 1)
    x1 = _randn((nb, nt, nf, c1), seed + 2) if c1 else None
    oc = H * (2 if bidir else 1)
    add = _randn((nb, nt, nf, oc), seed + 3) if use_addend else None
    g0 = ops.grid_copy(x0.to(DEV), c0, dt)
    g1 = ops.grid_copy(x1.to(DEV), c1, dt) if c1 else None
    ga = ops.grid_copy(add.to(DEV), oc, dt) if use_addend else None
    before = ops.LAUNCHES
    ga_ref = ga.float().cpu() if use_addend else None
    h, hs = run_lstm(p, engine, axis, g0, c0, g1, c1, addend=ga, inplace_addend=inplace)
    assert ops.LAUNCHES == before + 1
    if inplace and use_addend and engine == "tcgen05":
        assert hs.data_ptr() == ga.data_ptr()      # the sum really was accumulated into the residual operand
    if engine == "tcgen05":   # the layer really ran on the tensor-core kernel, not the CUDA-core fallback
        from fn_ssl_b200 import _lib
        assert _lib.load().fnssl_lstm_tc_supported(H, (c0 + 15) // 16 * 16, (c1 + 15) // 16 * 16) == 1
    # oracle on the same (dtype-rounded) inputs
    x = torch.cat([t for t in (g0[..., :c0].float().cpu(), g1[..., :c1].float().cpu() if c1 else None) if t is not None], -1)
    sd = {"l." + k: v.detach().cpu() for k, v in p.state_dict().items()}
    if axis == 0:
        ref = orc.lstm(x.reshape(nb * nt, nf, -1), sd, "l.").reshape(nb, nt, nf, oc)
    else:
        ref = orc.lstm(x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1), sd, "l.").reshape(nb, nf, nt, oc).permute(0, 2, 1, 3)
    e = _relerr(h, ref)
    if use_addend:
        e = max(e, _relerr(hs, ref + ga_ref))
    return e


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("H,bidir,c0,c1,addend", [
    (32, True, 4, 0, False), (64, False, 64, 4, False), (128, True, 256, 0, True),
    (256, False, 256, 4, True), (128, True, 5, 3, False), (256, False, 256, 0, False)])
def test_lstm_layer_simt(axis, H, bidir, c0, c1, addend):
    # ragged: rows not a multiple of the CTA tile, odd channel counts
    assert _lstm_case("simt", axis, 2, 7, 19, c0, c1, H, bidir, addend) <= 2e-5


@pytest.mark.parametrize("rows", ["64", "128"])
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("H,bidir,c0,c1,addend", [
    (128, True, 4, 0, False), (128, True, 256, 0, True), (128, True, 256, 4, True), (128, False, 256, 0, False),
    (128, True, 256, 8, False), (64, True, 8, 0, False), (128, False, 128, 8, False), (64, True, 128, 8, True),
    (256, False, 256, 4, True), (256, False, 256, 0, False), (256, False, 256, 8, False)])
def test_lstm_layer_tcgen05(monkeypatch, rows, axis, H, bidir, c0, c1, addend):
    """The tensor-core cluster kernel (lstm_tc4.cu): both row-tile shapes (128 / 64 sequences per sub-tile), ragged tiles
    (70 frames / 40 bins are not multiples of the tile), two-source inputs and the fused residual output."""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    if H == 256 and rows == "128":
        pytest.skip("H = 256 only exists with 64-row tiles")
    monkeypatch.setenv("FNSSL_TC_ROWS", rows)      # (a knob of lstm_tc4.cu: keep the pair kernels out of this test)
    _tc4_only(monkeypatch)
    nb, nt, nf = (2, 70, 256) if axis == 1 else (2, 70, 40)
    assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, H, bidir, addend) <= 1e-3
    if addend:   # residual sum accumulated in place (TMA reduce-add onto the residual operand)
        assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, H, bidir, addend, inplace=True) <= 1e-3


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("bidir,c0,c1,addend", [(True, 4, 0, False), (True, 256, 0, True), (True, 256, 4, True), (False, 256, 0, False),
                                                 (True, 256, 8, False), (False, 128, 8, False), (True, 64, 0, True)])
def test_lstm_layer_tcgen05_pair_kernel(monkeypatch, axis, bidir, c0, c1, addend):
    """The CTA-pair kernel (lstm_tc5.cu, tcgen05 cta_group::2; H = 128) forced on small layers: ragged chains (70 frames /
    40 bins against 256-row chains), an absent second chain, two-source inputs, the in-place residual output."""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    monkeypatch.setenv("FNSSL_TC_PAIR", "1")
    monkeypatch.setenv("FNSSL_TC_PAIR_MIN", "1")
    for nb, nt, nf in (((3, 5, 256), (2, 70, 300)) if axis == 1 else ((2, 70, 40), (5, 130, 7))):
        assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, 128, bidir, addend, inplace=addend) <= 1e-3


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("bidir,c0,c1,addend", [(False, 256, 0, True), (False, 256, 0, False), (False, 64, 0, False), (True, 16, 0, False),
                                                 (False, 128, 0, True), (False, 256, 4, True), (False, 256, 8, False), (True, 64, 16, True)])
def test_lstm_layer_tcgen05_pair_kernel_h256(monkeypatch, axis, bidir, c0, c1, addend):
    """The H = 256 CTA-pair kernel (lstm_tc6.cu: cta_group::2 with M = 128, 8-CTA clusters, 2x2 TMEM accumulator layout) forced on
    small layers: ragged 128-row chains, an absent second chain, the in-place residual output, a narrow (<= 16 channel) second
    source through its own ring.  (A layer with a carried state runs lstm_tc4.cu -- covered by the other LSTM tests.)"""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    monkeypatch.setenv("FNSSL_TC_PAIR256_MIN", "1")
    for nb, nt, nf in (((3, 5, 256), (2, 33, 300)) if axis == 1 else ((2, 70, 40), (5, 130, 7))):
        assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, 256, bidir, addend, inplace=addend) <= 1e-3


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("bidir,c0,c1,addend", [(True, 16, 0, False), (True, 256, 0, True), (True, 256, 4, True), (False, 64, 0, False),
                                                 (True, 256, 8, False), (False, 128, 16, True)])
def test_lstm_layer_tcgen05_pair_kernel_m128_h128(monkeypatch, axis, bidir, c0, c1, addend):
    """lstm_tc6.cu instantiated for H = 128 (clusters of 4 = 2 pairs, 128-row chains, M = 128 cta_group::2 MMAs) -- the kernel of
    mid-size H = 128 layers (cfg2's 16 utterances) -- forced on small layers: ragged chains, an absent second chain, both
    directions, the narrow second source, the in-place residual output."""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    monkeypatch.setenv("FNSSL_TC_PAIR128_MIN", "1")
    for nb, nt, nf in (((3, 5, 256), (2, 33, 300)) if axis == 1 else ((2, 70, 40), (5, 130, 7))):
        assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, 128, bidir, addend, inplace=addend) <= 1e-3


@pytest.mark.parametrize("kernel", ["simt", "tc4", "tc4_threads", "tc5", "tc6_h256", "tc6_h128"])
@pytest.mark.parametrize("axis", [0, 1])
def test_lstm_second_copy_of_the_output(monkeypatch, kernel, axis):
    """fnssl_lstm_forward with out1 and NO residual operand: out1 is a second copy of h, written by the kernel itself (TMA tile
    stores out of the same shared-memory tile; per-thread stores on the fallback paths) -- what FNblock 1 uses instead of a
    copy kernel.  Every LSTM kernel: both outputs are bit-identical, equal the single-output run, and match the oracle."""
    from fn_ssl_b200 import config, ops
    from fn_ssl_b200.packing import LSTMParams, run_lstm
    if kernel != "simt" and not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    eng = "simt" if kernel == "simt" else "tcgen05"
    H, c0, c1, bidir = {"simt": (64, 20, 4, True), "tc4": (128, 16, 0, True), "tc4_threads": (128, 64, 0, True), "tc5": (128, 16, 0, True),
                        "tc6_h256": (256, 64, 0, False), "tc6_h128": (128, 64, 16, True)}[kernel]
    _tc4_only(monkeypatch)
    if kernel == "tc4_threads":
        monkeypatch.setenv("FNSSL_TC_NO_TMA_OUT", "1")
    elif kernel == "tc5":
        monkeypatch.setenv("FNSSL_TC_PAIR", "1"); monkeypatch.setenv("FNSSL_TC_PAIR_MIN", "1")
    elif kernel == "tc6_h256":
        monkeypatch.setenv("FNSSL_TC_PAIR256", "1"); monkeypatch.setenv("FNSSL_TC_PAIR256_MIN", "1")
    elif kernel == "tc6_h128":
        monkeypatch.setenv("FNSSL_TC_PAIR", "1"); monkeypatch.setenv("FNSSL_TC_PAIR_MIN", "1000000"); monkeypatch.setenv("FNSSL_TC_PAIR128_MIN", "1")
    nb, nt, nf = (2, 70, 40) if axis == 0 else (3, 5, 300)
    dt = config.grid_dtype(eng)
    torch.manual_seed(5)
    p = LSTMParams(c0 + c1, H, bidirectional=bidir).to(DEV)
    g0 = ops.grid_copy(_randn((nb, nt, nf, c0), 6).to(DEV), c0, dt)
    g1 = ops.grid_copy(_randn((nb, nt, nf, c1), 7).to(DEV), c1, dt) if c1 else None
    h, h2 = run_lstm(p, eng, axis, g0, c0, g1, c1, duplicate=True)
    h_single, none = run_lstm(p, eng, axis, g0, c0, g1, c1)
    assert none is None and h2 is not None and h2.data_ptr() != h.data_ptr()
    assert torch.equal(h, h2) and torch.equal(h, h_single) and torch.isfinite(h).all()
    oc = H * (2 if bidir else 1)
    x = torch.cat([t for t in (g0[..., :c0].float().cpu(), g1[..., :c1].float().cpu() if c1 else None) if t is not None], -1)
    sd = {"l." + k: v.detach().cpu() for k, v in p.state_dict().items()}
    if axis == 0:
        ref = orc.lstm(x.reshape(nb * nt, nf, -1), sd, "l.").reshape(nb, nt, nf, oc)
    else:
        ref = orc.lstm(x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1), sd, "l.").reshape(nb, nf, nt, oc).permute(0, 2, 1, 3)
    assert _relerr(h2, ref) <= TOL[eng]
    with pytest.raises(RuntimeError, match="duplicate"):
        run_lstm(p, eng, axis, g0, c0, g1, c1, addend=h, duplicate=True)


@pytest.mark.parametrize("small1", ["1", "0"])
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("H,bidir,c0,c1,addend", [(128, True, 256, 4, True), (64, True, 128, 8, True), (256, False, 256, 8, True),
                                                   (128, False, 64, 16, False)])
def test_lstm_narrow_second_source_ring(monkeypatch, small1, axis, H, bidir, c0, c1, addend):
    """Generation 4's handling of a <= 16-channel second source (own two-entry ring of 32B-swizzled [rows x 16] slabs, 4 KB
    weight slab, every CTA fetching its own x slabs), forced on / off for every H (the dispatcher uses it for H = 256)."""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    monkeypatch.setenv("FNSSL_TC_SMALL1", small1)
    _tc4_only(monkeypatch)
    nb, nt, nf = (2, 70, 256) if axis == 1 else (2, 70, 40)
    assert _lstm_case("tcgen05", axis, nb, nt, nf, c0, c1, H, bidir, addend, inplace=addend) <= 1e-3


@pytest.mark.parametrize("axis,nb,nt,nf", [(0, 3, 100, 7), (1, 1, 6, 300), (1, 2, 5, 515)])
def test_lstm_layer_tcgen05_multi_tile(monkeypatch, axis, nb, nt, nf):
    """Generation 4 with more than one cluster tile and ragged last tiles on both sub-tiles (300 rows = 256 + 44,
    515 bins = 2 x 256 + 3), outputs through TMA tile stores / in-place reduce-add."""
    from fn_ssl_b200 import config
    if not config.TC_AVAILABLE:
        pytest.skip("tcgen05 engine not built")
    for rows in ("128", "64"):
        monkeypatch.setenv("FNSSL_TC_ROWS", rows)      # (a knob of lstm_tc4.cu: keep the pair kernels out of this test)
        _tc4_only(monkeypatch)
        assert _lstm_case("tcgen05", axis, nb, nt, nf, 64, 4, 128, True, True, inplace=True) <= 1e-3
        assert _lstm_case("tcgen05", axis, nb, nt, nf, 16, 0, 64, False, False) <= 1e-3


# ---------------------------------------------------------------------------------------------
# networks against the reference's golden outputs
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag,kw", [("on", dict(is_online=True)), ("off", dict(is_online=False)),
                                    ("doa", dict(is_online=True, is_doa=True))])
def test_fnssl_network_matches_reference_golden(golden_fnssl, tag, kw):
    import fn_ssl_b200 as F
    x = _randn((2, 4, 256, 26), 12).to(DEV)
    for eng in _engines():
        torch.manual_seed(3)
        net = F.FN_SSL(**kw).to(DEV).eval()
        net.engine = eng
        assert _relerr(net(x), golden_fnssl[f"net_{tag}"]) <= TOL[eng], eng


def test_fnblock_module_api_matches_reference_golden(golden_fnssl):
    import fn_ssl_b200 as F
    g = golden_fnssl
    blk = F.FNblock(input_size=4, hidden_size=64, is_online=True, is_first=True).eval()
    blk.load_state_dict({k[len("blk_first_sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("blk_first_sd.")})
    blk2 = F.FNblock(input_size=64, hidden_size=64, is_online=False, is_first=False).eval()
    blk2.load_state_dict({k[len("blk_next_sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("blk_next_sd.")})
    blk.to(DEV); blk2.to(DEV)
    blk.engine = blk2.engine = "simt"
    xb = _randn((1, 10, 16, 4), 13).to(DEV)
    y, fb, nbs = blk(xb)
    assert _relerr(y, g["blk_first_y"]) <= 2e-5 and _relerr(fb, g["blk_first_fb"]) <= 2e-5 and _relerr(nbs, g["blk_first_nb"]) <= 2e-5
    y2, fb2, nbs2 = blk2(y, fb_skip=fb, nb_skip=nbs)
    assert _relerr(y2, g["blk_next_y"]) <= 2e-5 and _relerr(fb2, g["blk_next_fb"]) <= 2e-5 and _relerr(nbs2, g["blk_next_nb"]) <= 2e-5


def test_fnblock_module_api_hidden256_both_engines(golden_fnssl):
    """The public FNblock.forward (grid_copy / grid_add / .float() glue, Model.py:31-50) at the paper's hidden size on BOTH
    engines -- i.e. including the tcgen05 product engine -- against the unmodified reference's outputs.  Weights:
    `torch.manual_seed(s); FNblock(...)` reproduces the reference module's default init (asserted in make_golden.py)."""
    import fn_ssl_b200 as F
    g = golden_fnssl
    xb = _randn((1, 6, 8, 4), 17).to(DEV)
    for eng in _engines():
        torch.manual_seed(15)
        blk = F.FNblock(input_size=4, hidden_size=256, is_online=True, is_first=True).eval().to(DEV)
        torch.manual_seed(16)
        blk2 = F.FNblock(input_size=256, hidden_size=256, is_online=False, is_first=False).eval().to(DEV)
        blk.engine = blk2.engine = eng
        before = _tc_launches()
        y, fb, nbs = blk(xb)
        tol = TOL[eng]
        assert y.shape == (1, 6, 8, 256) and fb.shape == (6, 8, 256) and nbs.shape == (8, 6, 256)
        assert _relerr(y, g["blk256_first_y"]) <= tol and _relerr(fb, g["blk256_first_fb"]) <= tol and _relerr(nbs, g["blk256_first_nb"]) <= tol
        # second block fed with the REFERENCE's first-block outputs, so its error is not compounded
        y2, fb2, nbs2 = blk2(torch.from_numpy(g["blk256_first_y"]).to(DEV), fb_skip=torch.from_numpy(g["blk256_first_fb"]).to(DEV),
                             nb_skip=torch.from_numpy(g["blk256_first_nb"]).to(DEV))
        assert _relerr(y2, g["blk256_next_y"]) <= tol and _relerr(fb2, g["blk256_next_fb"]) <= tol and _relerr(nbs2, g["blk256_next_nb"]) <= tol
        if eng == "tcgen05":
            assert _tc_launches() - before == 4      # all four LSTM layers ran on the tensor-core kernel


def _tc_launches():
    from fn_ssl_b200 import ops
    return ops.TC_LSTM_LAUNCHES


@pytest.mark.parametrize("tag,kw", [("d2", dict(input_size=4, hidden_size=128, max_track=2, is_online=True)),
                                    ("m4", dict(input_size=8, hidden_size=256, max_track=2, is_online=True)),
                                    ("off", dict(input_size=4, hidden_size=128, max_track=2, is_online=False))])
def test_ipdnet_network_matches_reference_golden(golden_ipdnet, tag, kw):
    import fn_ssl_b200 as F
    x = _randn((2, kw["input_size"], 64, 26), 22).to(DEV)
    for eng in _engines():
        torch.manual_seed(4)
        net = F.IPDnet(**kw).to(DEV).eval()
        net.engine = eng
        assert _relerr(net(x), golden_ipdnet[f"net_{tag}"]) <= TOL[eng], eng
        if tag == "off":
            net.n = 12
            assert _relerr(net(x, offline_inference=True), golden_ipdnet["net_off_chunked"]) <= TOL[eng], eng


def test_causcnn_matches_reference_golden(golden_ipdnet):
    import fn_ssl_b200 as F
    g = golden_ipdnet
    cnn = F.CausCnnBlock(inp_dim=20, out_dim=4, cnn_hidden_dim=128).eval()
    cnn.load_state_dict({k[len("cnn_sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cnn_sd.")})
    y = cnn.to(DEV)(_randn((2, 20, 12, 37), 23).to(DEV))
    assert _relerr(y, g["cnn_y"]) <= 2e-5


@pytest.mark.parametrize("cin0,cin1", [(32, 0), (256, 8), (128, 4)])
def test_causcnn_tensor_core_path(monkeypatch, cin0, cin1):
    """CausCnnBlock on fp16 grids: tcgen05 implicit-GEMM conv1/conv2 (two-source input, ragged bins/frames) against the
    oracle, and bit-for-bit agreement of its dispatch with the CUDA-core kernels' semantics (same input grids)."""
    import fn_ssl_b200 as F
    from fn_ssl_b200 import ops
    cin = cin0 + cin1
    torch.manual_seed(21)
    cnn = F.CausCnnBlock(inp_dim=cin, out_dim=6, cnn_hidden_dim=128).eval().to(DEV)
    x = _randn((2, cin, 150, 38), 24)                      # 150 bins: two 128-bin tiles, the second ragged
    g0 = ops.cfirst_to_grid(x[:, :cin0].to(DEV), torch.float16)
    g1 = ops.cfirst_to_grid(x[:, cin0:].to(DEV), torch.float16) if cin1 else None
    xq = torch.cat([g0[..., :cin0].float().cpu()] + ([g1[..., :cin1].float().cpu()] if cin1 else []), -1).permute(0, 3, 2, 1)
    ref = orc.causcnn(xq, {"conv." + k: v.detach().cpu() for k, v in cnn.state_dict().items()})
    y_tc = cnn.forward_grid(g0, cin0, g1, cin1)
    assert y_tc.shape == ref.shape == (2, 6, 150, 3)
    assert _relerr(y_tc, ref) <= 1e-3
    monkeypatch.setenv("FNSSL_CONV_ENGINE", "simt")
    y_simt = cnn.forward_grid(g0, cin0, g1, cin1)
    assert _relerr(y_simt, ref) <= 2e-5
    assert _relerr(y_tc, y_simt) <= 1e-3


# ---------------------------------------------------------------------------------------------
# end to end at BASELINE sizes (4 s @ 16 kHz): oracle on a small batch + size-independent properties
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dtype,C,ld", [(torch.float16, 256, 256), (torch.float16, 264, 272), (torch.float16, 20, 32),
                                         (torch.float32, 256, 256)])
def test_ipd_head_kernels(dtype, C, ld):
    """AvgPool(12) -> Linear(C, 2) -> tanh -> [ch0 | ch1] (Model.py:79-87): the 16-byte-per-lane fp16 kernel (C % 8 == 0),
    the generic fp16 kernel and the fp32 kernel against the same arithmetic in torch on the dtype-rounded grid."""
    from fn_ssl_b200 import ops
    nb, nt, nf = 2, 38, 40                                                    # 38 frames -> 3 pooled frames, 2 dropped
    g = torch.zeros((nb, nt, nf, ld), dtype=dtype, device=DEV)
    g[..., :C] = _randn((nb, nt, nf, C), 3).to(DEV).to(dtype)
    w, b = _randn((2, C), 4).to(DEV) * 0.2, _randn((2,), 5).to(DEV)
    out = ops.ipd_head(g, C, w, b)
    x = g[..., :C].float().cpu()[:, :36].reshape(nb, 3, 12, nf, C).mean(2)
    y = torch.tanh(x @ w.cpu().t() + b.cpu())                                 # (nb, 3, nf, 2)
    ref = torch.cat((y[..., 0], y[..., 1]), dim=-1)
    assert out.shape == ref.shape == (nb, 3, 2 * nf)
    assert _relerr(out, ref) <= 2e-5


@pytest.mark.parametrize("online", [True, False])
def test_fnssl_end_to_end_4s(online):
    import fn_ssl_b200 as F
    sig = orc.white_noise(2, 64000, 2)
    sd = orc.seeded_fnssl_state_dict(0, is_online=online)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = orc.fnssl_forward(orc.preprocess_fnssl(sig), sd, fast=True)
    for eng in _engines():
        net = F.FN_SSL(is_online=online).eval()
        net.load_state_dict(sd)
        net.to(DEV).engine = eng
        pipe = F.FNSSLPipeline(net)
        out = pipe(sig.to(DEV))
        assert out.shape == (2, 20, 512)
        assert _relerr(out, ref) <= TOL[eng], eng
        # same path through the reference-shaped API: data_preprocess -> FN_SSL.forward
        out2 = net(F.data_preprocess_fnssl(sig.to(DEV))[0])
        assert _relerr(out2, ref) <= TOL[eng], eng


def test_fnssl_batch16_properties(monkeypatch):
    """cfg2 size (B=16, 4 s), the code-default (online) model: utterances are independent -- any utterance of the batch equals
    its solo run (bit-exact when the same kernel serves both: at B = 16 the H = 256 layers run the CTA-pair kernel lstm_tc6.cu,
    a solo utterance lstm_tc4.cu, whose h-part accumulates in a different order -- so the bit-for-bit comparison is made with
    the pair kernels off and the two dispatches are compared at the engine's tolerance); outputs are tanh-bounded and finite."""
    import fn_ssl_b200 as F
    sig = orc.white_noise(16, 64000, 2).to(DEV)
    net = F.FN_SSL().eval()
    net.load_state_dict(orc.seeded_fnssl_state_dict(0))
    pipe = F.FNSSLPipeline(net.to(DEV))
    out = pipe(sig)
    assert out.shape == (16, 20, 512) and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0
    monkeypatch.setenv("FNSSL_TC_PAIR", "0")
    monkeypatch.setenv("FNSSL_TC_PAIR256", "0")
    out4 = pipe(sig)
    assert _relerr(out, out4) <= TOL[net._engine()]
    for b in (0, 7, 15):
        solo = pipe(sig[b:b + 1])
        assert torch.equal(solo[0], out4[b]), b


@pytest.mark.parametrize("tag,kw,B", [("cfg2 offline", dict(is_online=False), 16), ("cfg4 doa", dict(is_online=False, is_doa=True), 32),
                                      ("cfg4 doa online", dict(is_online=True, is_doa=True), 32)])
def test_fnssl_benched_configs_at_size(tag, kw, B, monkeypatch):
    """The configurations bench.py measures, at their per-GPU batch sizes (cfg2: 16 x 4 s offline; cfg4: the 32-utterance
    shard of the global 256 batch at 8 GPUs, DOA head): one utterance of the batch against the oracle, and every probed
    utterance bit-identical to its solo run (utterances are independent)."""
    import fn_ssl_b200 as F
    sig = orc.white_noise(B, 64000, 2)
    sd = orc.seeded_fnssl_state_dict(0, **kw)
    net = F.FN_SSL(**kw).eval()
    net.load_state_dict(sd)
    pipe = F.FNSSLPipeline(net.to(DEV))
    out = pipe(sig.to(DEV))
    width = 180 if kw.get("is_doa") else 512
    assert out.shape == (B, 20, width) and torch.isfinite(out).all()
    k = B - 3
    ref = orc.fnssl_forward(orc.preprocess_fnssl(sig[k:k + 1]), sd, fast=True)
    assert _relerr(out[k:k + 1], ref) <= TOL[net._engine()]
    # At these sizes the H = 128 layers run the CTA-pair kernel (lstm_tc5.cu); a solo utterance is a small grid and runs
    # lstm_tc4.cu, whose h-part accumulates the chunks in a different order -- so the bit-for-bit solo comparison is made with
    # the pair kernel switched off, and the two kernels are compared with each other at the engine's tolerance.
    monkeypatch.setenv("FNSSL_TC_PAIR", "0")
    monkeypatch.setenv("FNSSL_TC_PAIR256", "0")      # (the online model's H = 256 layers: lstm_tc6.cu at batch size, lstm_tc4.cu solo)
    out4 = pipe(sig.to(DEV))
    assert _relerr(out, out4) <= TOL[net._engine()]
    for b in (0, k, B - 1):
        assert torch.equal(pipe(sig[b:b + 1].to(DEV))[0], out4[b]), b


def test_ipdnet_cfg3_at_size(monkeypatch):
    """BASELINE configs[2]: IPDnet 4-mic, hidden 256, online, batch 32 x 4 s -- one utterance against the oracle, solo runs
    bit-identical."""
    import fn_ssl_b200 as F
    B = 32
    sig = orc.white_noise(B, 64000, 4)
    kw = dict(input_size=8, hidden_size=256, max_track=2, is_online=True)
    sd = orc.seeded_ipdnet_state_dict(0, **kw)
    net = F.IPDnet(**kw).eval()
    net.load_state_dict(sd)
    pipe = F.IPDnetPipeline(net.to(DEV))
    out = pipe(sig.to(DEV))
    assert out.shape == (B, 20, 512, 3, 2) and torch.isfinite(out).all()
    k = 21
    ref = orc.ipdnet_forward(orc.preprocess_ipdnet(sig[k:k + 1]), sd, fast=True)
    assert _relerr(out[k:k + 1], ref) <= TOL[net._engine()]
    # (as above: the full-band BLSTM(2x128) layers run the CTA-pair kernel at this size and lstm_tc4.cu solo)
    monkeypatch.setenv("FNSSL_TC_PAIR", "0")
    monkeypatch.setenv("FNSSL_TC_PAIR256", "0")
    out4 = pipe(sig.to(DEV))
    assert _relerr(out, out4) <= TOL[net._engine()]
    for b in (0, k, B - 1):
        assert torch.equal(pipe(sig[b:b + 1].to(DEV))[0], out4[b]), b


def test_silence_then_tone_stays_finite_on_fp16_grids():
    """A tonal onset after digital silence drives re/(mu+eps) beyond the fp16 range (mu is still ~0 when the tone starts):
    the fp16 feature grid saturates at +-65504 instead of overflowing to inf, so the along-time LSTM state stays finite and the
    output stays close to the fp32 engine's."""
    import fn_ssl_b200 as F
    n = 64000
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    sig = torch.zeros(1, n, 2)
    sig[0, 20000:, 0] = 1000.0 * torch.sin(2 * np.pi * 1000.0 * t[20000:])
    sig[0, 20000:, 1] = 1000.0 * torch.sin(2 * np.pi * 1000.0 * t[20000:] + 0.3)
    sd = orc.seeded_fnssl_state_dict(0)
    outs = {}
    for eng in _engines():
        net = F.FN_SSL().eval()
        net.load_state_dict(sd)
        net.to(DEV).engine = eng
        outs[eng] = F.FNSSLPipeline(net)(sig.to(DEV))
        assert torch.isfinite(outs[eng]).all(), eng
    from fn_ssl_b200 import ops
    spec, magsum = ops.stft(sig.to(DEV), want_magsum=True)
    g16, _, _ = ops.features(spec, magsum, "MM", ops.NORM_FORGETTING, 298, 1e-6, torch.float16)
    assert torch.isfinite(g16).all() and float(g16.float().abs().max()) <= 65504.0


def test_tensors_on_a_non_current_device():
    """Models / tensors living on cuda:1 while cuda:0 is the current device: every wrapper makes the tensors' device
    current for the launch (ops.on_tensor_device), so results equal the cuda:0 run bit for bit."""
    import fn_ssl_b200 as F
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sig = orc.white_noise(2, 512 + 256 * 25, 2, seed=9)
    sd = orc.seeded_fnssl_state_dict(0)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        net = F.FN_SSL().eval()
        net.load_state_dict(sd)
        outs.append(F.FNSSLPipeline(net.to(dev))(sig.to(dev)).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])


def test_ipdnet_end_to_end_4mic():
    import fn_ssl_b200 as F
    sig = orc.white_noise(1, 64000, 4)
    kw = dict(input_size=8, hidden_size=256, max_track=2, is_online=True)
    sd = orc.seeded_ipdnet_state_dict(0, **kw)
    ref = orc.ipdnet_forward(orc.preprocess_ipdnet(sig), sd, fast=True)
    for eng in _engines():
        net = F.IPDnet(**kw).eval()
        net.load_state_dict(sd)
        net.to(DEV).engine = eng
        out = F.IPDnetPipeline(net)(sig.to(DEV))
        assert out.shape == (1, 20, 512, 3, 2)
        assert _relerr(out, ref) <= TOL[eng], eng


# ---------------------------------------------------------------------------------------------
# "next" row: IPD -> DOA decoding
# ---------------------------------------------------------------------------------------------

def test_dpipd_templates_match_reference_golden(golden_fnssl):
    import fn_ssl_b200 as F
    g = golden_fnssl
    mic3 = np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0), (0.0, 0.05, 0.01)))
    for mode in ("M", "MM"):
        d = F.DPIPD(ndoa_candidate=[7, 13], mic_location=mic3, nf=33, fre_max=8000, ch_mode=mode, speed=340)
        tpl, gt, cand = d(source_doa=np.array([[[[1.2, 0.7], [0.4, -2.0]]]]))
        assert _relerr(tpl.real.astype(np.float32), g[f"dec_template_{mode}_re"]) <= 1e-6
        assert _relerr(tpl.imag.astype(np.float32), g[f"dec_template_{mode}_im"]) <= 1e-6
        assert _relerr(gt.real.astype(np.float32), g[f"dec_gt_{mode}_re"]) <= 1e-6
        assert _relerr(gt.imag.astype(np.float32), g[f"dec_gt_{mode}_im"]) <= 1e-6
        assert len(cand[0]) == 7 and len(cand[1]) == 13


@pytest.mark.parametrize("snm", ["kNum", "unkNum", "KNum"])
def test_source_detect_localize_matches_reference_golden(golden_fnssl, snm):
    import fn_ssl_b200 as F
    g = golden_fnssl
    d = F.DPIPD(ndoa_candidate=[37, 73], mic_location=np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), nf=257, fre_max=8000,
                ch_mode="MM", speed=340)
    t2, cand2 = orc.doa_templates_for_decode(d.dpipd_template)      # the slicing PredDOA.predgt2DOA applies
    sdl = F.SourceDetectLocalize(max_num_sources=2, source_num_mode=snm, meth_mode="IDL")
    doa, vad, ss = sdl(torch.from_numpy(g["dec_pred_ipd"]).to(DEV), torch.from_numpy(t2), cand2)
    assert np.array_equal(doa.cpu().numpy(), g[f"dec_doa_{snm}"])          # identical candidates picked (index work: exact)
    assert _relerr(ss, g["dec_ss"]) <= 1e-5
    if snm == "unkNum":
        assert _relerr(vad, g[f"dec_vad_{snm}"]) <= 1e-5
    else:
        assert np.array_equal(vad.cpu().numpy(), g[f"dec_vad_{snm}"])      # 'KNum' (main.py's spelling) leaves the VADs at 0


def test_pred_ipd_to_doa_matches_reference_golden(golden_fnssl):
    import fn_ssl_b200 as F
    g = golden_fnssl
    gd = F.DPIPD(ndoa_candidate=[37, 73], mic_location=np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), nf=257, fre_max=8000,
                 ch_mode="MM", speed=340)
    sdl = F.SourceDetectLocalize(max_num_sources=1, source_num_mode="kNum", meth_mode="IDL")
    netout = _randn((3, 4, 512), 16).tanh().to(DEV)
    out = F.pred_ipd_to_doa(netout, gd, sdl, ch_mode="MM")
    assert np.array_equal(out["doa"].cpu().numpy(), g["dec_preddoa_doa"])
    assert np.array_equal(out["vad_sources"].cpu().numpy(), g["dec_preddoa_vad"])
    assert _relerr(out["spatial_spectrum"], g["dec_preddoa_ss"]) <= 1e-5


def test_decode_many_pairs_and_sources():
    """8-mic 'MM' (28 pairs, K = 14336 > one shared-memory chunk) and 3 sources against the oracle."""
    import fn_ssl_b200 as F
    rng = np.random.RandomState(3)
    mics = rng.uniform(-0.1, 0.1, size=(8, 3))
    d = F.DPIPD(ndoa_candidate=[9, 19], mic_location=mics, nf=257, fre_max=8000, ch_mode="MM", speed=343.0)
    tpl = d.dpipd_template
    t = np.concatenate((tpl.real[:, :, 1:257, :], tpl.imag[:, :, 1:257, :]), axis=2).astype(np.float32)
    T = torch.from_numpy(t)
    pred = 0.6 * T[2, 3][None, None] + 0.5 * T[6, 15][None, None] + 0.3 * T[4, 9][None, None] + 0.02 * _randn((2, 5, 512, 28), 8)
    ref = orc.source_detect_localize_idl(pred, T, d.doa_candidate, 3, "unkNum")
    sdl = F.SourceDetectLocalize(max_num_sources=3, source_num_mode="unkNum", meth_mode="IDL")
    doa, vad, ss = sdl(pred.to(DEV), T, d.doa_candidate)
    assert np.array_equal(doa.cpu().numpy(), ref[0].numpy())
    assert _relerr(vad, ref[1]) <= 1e-4 and _relerr(ss, ref[2]) <= 1e-4


def test_source_detect_localize_degenerate_rows_follow_torch_argmax():
    """Rows whose spatial spectrum is all NaN / all -inf (e.g. a NaN network output): torch.argmax -- what the reference calls
    at Module.py:558 -- returns the first NaN (NaN counts as the maximum) or index 0; the kernel must do the same and never
    index the templates out of range."""
    import fn_ssl_b200 as F
    d = F.DPIPD(ndoa_candidate=[37, 73], mic_location=np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), nf=257, fre_max=8000,
                ch_mode="MM", speed=340)
    t2, cand = orc.doa_templates_for_decode(d.dpipd_template)      # horizontal plane, azimuth 0..pi: (1, 37, 512, 1)
    T = torch.from_numpy(t2)
    pred = 0.8 * T[0, 11][None, None].repeat(1, 4, 1, 1) + 0.01 * _randn((1, 4, 512, 1), 3)
    pred[0, 1] = float("nan")                       # whole row NaN -> spectrum row all NaN
    pred[0, 2, 5, 0] = float("inf")                 # inf * (+/-) template values -> NaN / +-inf mix
    sdl = F.SourceDetectLocalize(max_num_sources=1, source_num_mode="kNum", meth_mode="IDL")
    doa, vad, ss = sdl(pred.to(DEV), T, cand)
    torch.cuda.synchronize()                        # a bad index would have faulted by now
    smap = (pred.reshape(4, 512) @ T.reshape(37, 512).t()) / 256.0
    want = smap.argmax(dim=1)
    assert int(want[0]) == 11 and int(want[1]) == 0
    azi = torch.as_tensor(np.asarray(cand[1]), dtype=torch.float32)
    got = doa.cpu()[0, :, 1, 0]
    assert torch.equal(got[[0, 3]], azi[want[[0, 3]]])          # ordinary rows
    assert float(got[1]) == float(azi[want[1]])                 # all-NaN row: first NaN = candidate 0, as torch.argmax
    assert 0.0 <= float(got[2]) <= float(azi[-1])               # mixed row: a valid candidate


# ------------------------------------------------------------------------------------------------
# "next" row: stateful (streaming) API -- carried LSTM state, forgetting-norm state, STFT overlap
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("rows", ["64", "128"])
@pytest.mark.parametrize("engine,H,c0,c1", [("simt", 64, 20, 0), ("simt", 256, 256, 4), ("tcgen05", 64, 64, 0),
                                            ("tcgen05", 128, 256, 16), ("tcgen05", 256, 256, 0), ("tcgen05", 256, 128, 16)])
def test_lstm_carried_state_equals_whole_sequence(monkeypatch, rows, engine, H, c0, c1):
    """nn.LSTM semantics of (h_0, c_0) -> (h_n, c_n): running a narrow-band layer over 3 chunks with the state carried
    gives bit-identical outputs to one run over the whole sequence, and the final state matches the oracle."""
    from fn_ssl_b200 import config, ops
    from fn_ssl_b200.packing import LSTMParams, run_lstm
    monkeypatch.setenv("FNSSL_TC_ROWS", rows)      # (a knob of lstm_tc4.cu: keep the pair kernels out of this test)
    _tc4_only(monkeypatch)
    if engine == "simt" and rows == "64":
        pytest.skip("row-tile shapes only exist in the tensor-core engine")
    dt = config.grid_dtype(engine)
    nb, nt, nf = 2, 21, 75                           # 150 rows: ragged against both tile sizes
    torch.manual_seed(7)
    p = LSTMParams(c0 + c1, H, bidirectional=False).to(DEV)
    g0 = ops.grid_copy(_randn((nb, nt, nf, c0), 70).to(DEV), c0, dt)
    g1 = ops.grid_copy(_randn((nb, nt, nf, c1), 71).to(DEV), c1, dt) if c1 else None
    ga = ops.grid_copy(_randn((nb, nt, nf, H), 72).to(DEV), H, dt)
    whole_h, whole_s = run_lstm(p, engine, ops.ALONG_TIME, g0, c0, g1, c1, addend=ga)
    state = (torch.zeros(nb * nf, H, device=DEV), torch.zeros(nb * nf, H, device=DEV))
    hs, ss = [], []
    for a, b in ((0, 1), (1, 9), (9, 21)):          # a one-step chunk exercises the steps == 1 path
        h, s = run_lstm(p, engine, ops.ALONG_TIME, g0[:, a:b].contiguous(), c0, g1[:, a:b].contiguous() if c1 else None, c1,
                        addend=ga[:, a:b].contiguous(), state=state)
        hs.append(h); ss.append(s)
    assert torch.equal(torch.cat(hs, 1), whole_h)
    assert torch.equal(torch.cat(ss, 1), whole_s)
    # final state against the oracle (h_n is the last output; c_n from the fp32 recursion)
    x = torch.cat([t for t in (g0[..., :c0].float().cpu(), g1[..., :c1].float().cpu() if c1 else None) if t is not None], -1)
    sd = {"l." + k: v.detach().cpu() for k, v in p.state_dict().items()}
    xs = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
    lstm = torch.nn.LSTM(c0 + c1, H, batch_first=True)
    lstm.load_state_dict({k[2:]: v for k, v in sd.items()})
    with torch.no_grad():
        ref_y, (ref_h, ref_c) = lstm(xs)
    assert _relerr(orc.lstm(xs, sd, "l."), ref_y) <= 1e-6      # the torch module and the oracle agree (oracle is the checker)
    tol = 2e-5 if engine == "simt" else 1e-3
    assert _relerr(state[0], ref_h[0]) <= tol
    assert _relerr(state[1], ref_c[0]) <= tol
    assert torch.equal(state[0].to(whole_h.dtype), whole_h[:, -1].reshape(nb * nf, -1)[:, :H])


def test_lstm_state_rejects_bidirectional():
    from fn_ssl_b200 import ops
    from fn_ssl_b200.packing import LSTMParams, run_lstm
    p = LSTMParams(8, 32, bidirectional=True).to(DEV)
    g = ops.grid_copy(_randn((1, 4, 6, 8), 1).to(DEV), 8, torch.float32)
    st = (torch.zeros(6, 32, device=DEV), torch.zeros(6, 32, device=DEV))
    with pytest.raises(RuntimeError, match="uni-directional"):
        run_lstm(p, "simt", ops.ALONG_TIME, g, 8, None, 0, state=st)


def test_norm_stream_continues_the_recursion():
    """forgetting_norm over frames [0, nt) in three pieces == one pass (bit-exact), across the t < L / t >= L switch."""
    from fn_ssl_b200 import ops
    sig = _randn((2, 512 + 256 * 39, 3), 81).to(DEV)
    spec, magsum = ops.stft(sig, want_magsum=True)
    _, mu_whole, _ = ops.features(spec, magsum, "MM", ops.NORM_FORGETTING, 16, 1e-6, torch.float32)
    state = torch.zeros(6, device=DEV)
    parts, t0 = [], 0
    for a, b in ((0, 5), (5, 17), (17, 40)):
        parts.append(ops.norm_stream(magsum[:, :, a:b].contiguous(), "MM", 16, t0, state))
        t0 = b
    mu = torch.cat(parts, 1)
    assert torch.equal(mu, mu_whole)
    assert torch.equal(state, mu_whole[:, -1])
    g_whole, _, _ = ops.features(spec, magsum, "MM", ops.NORM_FORGETTING, 16, 1e-6, torch.float32)
    g_given, _, _ = ops.features(spec, None, "MM", ops.NORM_GIVEN, 16, 1e-6, torch.float32, mu=mu)
    assert torch.equal(g_whole, g_given)


@pytest.mark.parametrize("engine", ["tcgen05", "simt"])
def test_fnssl_stream_equals_whole_clip(engine):
    """Feeding a clip to FNSSLStream in uneven pieces reproduces the whole-clip pipeline output exactly."""
    import fn_ssl_b200 as F
    from fn_ssl_b200.streaming import FNSSLStream
    torch.manual_seed(11)
    net = F.FN_SSL(is_online=True).eval().to(DEV)
    net.engine = engine
    nb, nsample = 2, 512 + 256 * 40 + 100            # 41 frames -> 3 output frames, 5 frames + 100 samples left over
    sig = _randn((nb, nsample, 2), 82).to(DEV)
    whole = F.FNSSLPipeline(net)(sig)                # (2, 3, 512)
    st = FNSSLStream(net, nb=nb, nch=2)
    outs, pos = [], 0
    for n in (300, 3000, 3072, 1, 2500, nsample):    # last piece: whatever is left
        piece = sig[:, pos:pos + n]
        pos = min(nsample, pos + n)
        o = st.push(piece)
        if o is not None:
            outs.append(o)
    got = torch.cat(outs, 1)
    assert got.shape == whole.shape
    assert torch.equal(got, whole)
    assert st.frames_done == 36 and st.pending_samples == nsample - 36 * 256
    # oracle check of the streamed result (the whole-clip pipeline is itself checked against the oracle elsewhere)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = orc.fnssl_forward(orc.preprocess_fnssl(sig.cpu(), "MM"), sd, fast=True)
    assert _relerr(got, ref) <= (2e-5 if engine == "simt" else 1e-3)
    st.reset()
    assert st.push(sig[:, :512 + 256 * 11]).shape == (nb, 1, 512)


def test_fnssl_stream_rejects_offline_model():
    import fn_ssl_b200 as F
    from fn_ssl_b200.streaming import FNSSLStream
    with pytest.raises(RuntimeError, match="offline"):
        FNSSLStream(F.FN_SSL(is_online=False).eval().to(DEV), nb=1)


@pytest.mark.parametrize("kw,nch", [(dict(input_size=4, hidden_size=128, max_track=2, is_online=True), 2),
                                    (dict(input_size=8, hidden_size=256, max_track=2, is_online=True), 4)])
def test_ipdnet_stream_equals_whole_clip(kw, nch):
    """IPDnetStream (carried LSTM state + 36 frames of conv history) reproduces the whole-clip output exactly."""
    import fn_ssl_b200 as F
    torch.manual_seed(12)
    net = F.IPDnet(**kw).eval().to(DEV)
    nb, nsample = 2, 512 + 256 * 63 + 17             # 64 frames -> 5 output frames (4 frames left over)
    sig = _randn((nb, nsample, nch), 83).to(DEV)
    whole = F.IPDnetPipeline(net)(sig)               # (2, 5, 512, nch-1, 2)
    st = F.IPDnetStream(net, nb=nb)
    outs, pos = [], 0
    for n in (3328, 3072, 100, 6144, 2972, nsample):
        o = st.push(sig[:, pos:pos + n])
        pos = min(nsample, pos + n)
        if o is not None:
            outs.append(o)
    got = torch.cat(outs, 1)
    assert got.shape == whole.shape == (nb, 5, 512, nch - 1, 2)
    assert torch.equal(got, whole)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = orc.ipdnet_forward(orc.preprocess_ipdnet(sig.cpu()), sd)
    assert _relerr(got, ref) <= 1e-3


def test_run_host_staging_matches_device_call():
    """FNSSLPipeline.run_host: pinned host input through the double-buffered side-stream copy == the device-tensor call,
    for a sequence of different inputs (buffer reuse must wait for the forward that read the buffer)."""
    import fn_ssl_b200 as F
    torch.manual_seed(0)
    net = F.FN_SSL(is_online=True).eval().to(DEV)
    pipe = F.FNSSLPipeline(net)
    sigs = [_randn((2, 512 + 256 * 23, 2), 70 + i).pin_memory() for i in range(5)]
    outs_host = [torch.empty((2, 2, 512), dtype=torch.float32).pin_memory() for _ in range(5)]
    outs = [pipe.run_host(s, o).clone() for s, o in zip(sigs, outs_host)]
    torch.cuda.synchronize()
    for s, o, oh in zip(sigs, outs, outs_host):
        ref = pipe(s.to(DEV))
        assert torch.equal(o, ref) and torch.equal(oh.to(DEV), ref)
    with pytest.raises(RuntimeError, match="pinned"):
        pipe.run_host(torch.zeros(2, 8192, 2))
