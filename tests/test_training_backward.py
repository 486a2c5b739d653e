"""Training step of the hot path: backward pass of the LSTM layers and the DP-IPD head.

Goldens (tests/golden/grad_golden.npz, made by tests/golden/make_golden.py grad): the UNMODIFIED reference FN_SSL / FNblock in
train mode (dropout probability 0 so the step is deterministic), MSE loss, loss.backward().
CPU: the oracle's autograd against those goldens.  GPU: the CUDA backward kernels (through the C ABI, torch.autograd.Function
wrappers of fn_ssl_b200.training) against torch's autograd of nn.LSTM on the CPU and against the goldens.
Tolerance: max |g - g_ref| <= 1e-4 max |g_ref| per gradient tensor (fp32 kernels; weight gradients are reduced with fp32 atomics)."""
import os

import numpy as np
import pytest
import torch

from oracle import fnssl_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "grad_golden.npz"))


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _randn(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def _check_against_golden(g, tag, named_grads, out, loss):
    assert _rel(out, g[f"{tag}_out"]) <= 2e-5
    assert abs(float(loss.detach()) - float(g[f"{tag}_loss"])) <= 2e-5 * float(g[f"{tag}_loss"])
    names = [str(n) for n in g[f"{tag}_names"]]
    assert list(named_grads.keys()) == names
    for i, n in enumerate(names):
        gr = named_grads[n].detach().cpu().double()
        assert abs(float(gr.norm()) - g[f"{tag}_grad_norm"][i]) <= TOL * g[f"{tag}_grad_norm"][i], n
        assert abs(float(gr.sum()) - g[f"{tag}_grad_sum"][i]) <= 5 * TOL * g[f"{tag}_grad_norm"][i], n
    for k in g.files:
        if k.startswith(f"{tag}_grad_") and k not in (f"{tag}_grad_norm", f"{tag}_grad_sum"):
            n = k[len(f"{tag}_grad_"):]
            if n.endswith("[:, :4]"):
                assert _rel(named_grads[n[:-7]][:, :4], g[k]) <= TOL, k
            else:
                assert _rel(named_grads[n], g[k]) <= TOL, k


# ---- CPU: the oracle's autograd is pinned to the reference's ------------------------------------------------------------------

@pytest.mark.parametrize("tag,kw", [("off", dict(is_online=False)), ("on", dict(is_online=True))])
def test_oracle_autograd_matches_reference_gradients(g, tag, kw):
    x, tgt = _randn((1, 4, 256, 24), 21), _randn((1, 2, 512), 22).tanh()
    sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_fnssl_state_dict(3, **kw).items()}
    y = orc.fnssl_forward(x, sd, fast=True)
    loss = torch.nn.functional.mse_loss(y, tgt)
    loss.backward()
    _check_against_golden(g, tag, {k: v.grad for k, v in sd.items()}, y, loss)


def test_backward_abi_symbols_and_validation():
    import ctypes
    from fn_ssl_b200 import _lib
    lib = _lib.load()
    assert lib.fnssl_lstm_train_saved_bytes(2, 3, 5, 32, 2) == 2 * 2 * 3 * 5 * 32 * 5 * 4
    a = _lib.LstmArgs()
    a.engine, a.dtype = _lib.ENGINE_TCGEN05, _lib.F16                 # the training path is the fp32 engine: rejected before any CUDA call
    assert lib.fnssl_lstm_forward_train(ctypes.byref(a), None, 0, None) != 0 and b"fp32" in lib.fnssl_last_error()
    assert lib.fnssl_lstm_backward(ctypes.byref(a), None, 0, None, None, 0, None, 0, None, 0, None, None) != 0
    assert lib.fnssl_ipd_head_backward(None, 0, 1, 12, 4, 8, None, None, None, None, 0, None, None, None) != 0


def test_gradient_unpacking_is_the_inverse_of_weight_packing():
    from fn_ssl_b200.packing import pack_lstm_simt, pack_lstm_whh_t, unpack_lstm_simt_grad
    H, I = 32, 5
    dirs = [tuple(_randn(s, 40 + i) for i, s in enumerate(((4 * H, I), (4 * H, H), (4 * H,), (4 * H,)))) for _ in range(2)]
    un = unpack_lstm_simt_grad(pack_lstm_simt(dirs), 2, I, H)
    for d in range(2):
        assert torch.equal(un[d][0], dirs[d][0]) and torch.equal(un[d][1], dirs[d][1])
        assert torch.allclose(un[d][2], dirs[d][2] + dirs[d][3]) and torch.equal(un[d][2], un[d][3])
    t = pack_lstm_whh_t(dirs)
    assert t.shape == (2, H, H, 4) and float(t[1, 3, 7, 2]) == float(dirs[1][1][2 * H + 3, 7])


# ---- GPU: one layer against torch's autograd of nn.LSTM ----------------------------------------------------------------------

def _sequences(grid, axis):
    """(nb, nt, nf, C) -> (rows, steps, C) as the reference reshapes it (Model.py:35,41)."""
    nb, nt, nf, C = grid.shape
    return grid.reshape(nb * nt, nf, C) if axis == 0 else grid.permute(0, 2, 1, 3).reshape(nb * nf, nt, C)


def _grid_from_sequences(seq, axis, nb, nt, nf):
    return seq.reshape(nb, nt, nf, -1) if axis == 0 else seq.reshape(nb, nf, nt, -1).permute(0, 2, 1, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("H,bidir,c0,ld0,c1,ld1", [(32, True, 4, 4, 0, 0), (64, False, 20, 24, 0, 0), (128, True, 24, 24, 4, 8),
                                                  (256, False, 12, 12, 3, 4), (128, True, 256, 256, 0, 0)])
def test_lstm_layer_gradients_match_torch_autograd(axis, H, bidir, c0, ld0, c1, ld1):
    from fn_ssl_b200 import training as T
    from fn_ssl_b200.packing import LSTMParams
    nb, nt, nf = 2, 9, 11                                    # 18 / 22 sequences: a partial row tile in every CTA
    torch.manual_seed(7)
    ref = torch.nn.LSTM(c0 + c1, H, batch_first=True, bidirectional=bidir)
    params = LSTMParams(c0 + c1, H, bidirectional=bidir)
    params.load_state_dict(ref.state_dict())
    params = params.cuda()
    s0, s1 = _randn((nb, nt, nf, ld0), 1), (_randn((nb, nt, nf, ld1), 2) if c1 else None)
    dout = _randn((nb, nt, nf, H * (2 if bidir else 1)), 3)
    # reference: nn.LSTM on the CPU over the reference's sequence layout
    r0 = s0.clone().requires_grad_(True)
    r1 = s1.clone().requires_grad_(True) if c1 else None
    xin = r0[..., :c0] if not c1 else torch.cat((r0[..., :c0], r1[..., :c1]), dim=-1)
    href = _grid_from_sequences(ref(_sequences(xin, axis))[0], axis, nb, nt, nf)
    (href * dout).sum().backward()
    # CUDA path
    d0 = s0.cuda().requires_grad_(True)
    d1 = s1.cuda().requires_grad_(True) if c1 else None
    h = T.lstm_layer(d0, c0, d1, c1, params, axis)
    (h * dout.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert _rel(h, href) <= 2e-5
    assert _rel(d0.grad, r0.grad) <= TOL                      # padding channels (ld0 > c0) carry zero gradient on both sides
    if c1:
        assert _rel(d1.grad, r1.grad) <= TOL
    for (n, p_), (_, q_) in zip(ref.named_parameters(), params.named_parameters()):
        assert _rel(q_.grad, p_.grad) <= TOL, n
    with pytest.raises(RuntimeError):                        # the saved activations are overwritten by the first backward
        h.backward(dout.cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("rpt", ["8", "16"])
@pytest.mark.parametrize("axis", [0, 1])
def test_lstm_layer_rows_per_thread_variants(monkeypatch, rpt, axis):
    """The 8- / 16-rows-per-thread instantiations of the training forward and the BPTT kernel (chosen by grid coverage from batch
    8 / 16 up, i.e. never on test-sized grids) forced onto a small grid with a partial row tile."""
    monkeypatch.setenv("FNSSL_TRAIN_RPT", rpt)
    test_lstm_layer_gradients_match_torch_autograd(axis, 128, True, 24, 24, 4, 8)


@pytest.mark.gpu
def test_lstm_layer_without_input_gradient_and_weight_only_sources():
    """First layer of the network: the feature grid needs no gradient (dsrc0 = NULL); second source with a gradient only."""
    from fn_ssl_b200 import training as T
    from fn_ssl_b200.packing import LSTMParams
    torch.manual_seed(8)
    ref = torch.nn.LSTM(36, 64, batch_first=True, bidirectional=True)
    params = LSTMParams(36, 64, bidirectional=True)
    params.load_state_dict(ref.state_dict())
    params = params.cuda()
    nb, nt, nf = 1, 7, 70
    s0, s1, dout = _randn((nb, nt, nf, 32), 4), _randn((nb, nt, nf, 4), 5), _randn((nb, nt, nf, 128), 6)
    r1 = s1.clone().requires_grad_(True)
    href = _grid_from_sequences(ref(_sequences(torch.cat((s0, r1), -1), 1))[0], 1, nb, nt, nf)
    (href * dout).sum().backward()
    d1 = s1.cuda().requires_grad_(True)
    h = T.lstm_layer(s0.cuda(), 32, d1, 4, params, 1)
    (h * dout.cuda()).sum().backward()
    assert _rel(d1.grad, r1.grad) <= TOL
    for (n, p_), (_, q_) in zip(ref.named_parameters(), params.named_parameters()):
        assert _rel(q_.grad, p_.grad) <= TOL, n


@pytest.mark.gpu
@pytest.mark.parametrize("nt", [24, 31])
def test_ipd_head_train_gradients(nt):
    from fn_ssl_b200 import training as T
    nb, nf, Cc = 2, 19, 256
    x, w, b = _randn((nb, nt, nf, Cc), 9), 0.1 * _randn((2, Cc), 10), _randn((2,), 11)
    dy = _randn((nb, nt // 12, 2 * nf), 12)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    seq = xr.permute(0, 2, 1, 3).reshape(nb * nf, nt, Cc)                           # Model.py:79-87
    ipd = torch.tanh(torch.nn.functional.linear(torch.nn.AvgPool2d(kernel_size=(12, 1))(seq), wr, br))
    ipd = ipd.view(nb, nf, nt // 12, 2).permute(0, 2, 1, 3)
    yref = torch.cat((ipd[..., 0], ipd[..., 1]), dim=2)
    (yref * dy).sum().backward()
    xd, wd, bd = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = T.ipd_head_train(xd, wd, bd)
    (y * dy.cuda()).sum().backward()
    assert _rel(y, yref) <= 1e-5
    assert _rel(xd.grad, xr.grad) <= TOL and _rel(wd.grad, wr.grad) <= TOL and _rel(bd.grad, br.grad) <= TOL


# ---- GPU: the network's training step against the reference's ----------------------------------------------------------------

def _no_dropout(net):
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return net


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", [("off", dict(is_online=False)), ("on", dict(is_online=True))])
def test_fnssl_training_step_matches_reference_golden(g, tag, kw):
    import fn_ssl_b200 as F
    net = F.FN_SSL(**kw)
    net.load_state_dict(orc.seeded_fnssl_state_dict(3, **kw))
    net = _no_dropout(net.cuda().train())
    x, tgt = _randn((1, 4, 256, 24), 21).cuda(), _randn((1, 2, 512), 22).tanh().cuda()
    y = net(x)
    loss = torch.nn.functional.mse_loss(y, tgt)
    loss.backward()
    torch.cuda.synchronize()
    _check_against_golden(g, tag, {n: p_.grad for n, p_ in net.named_parameters()}, y, loss)
    # eval mode on the same module is still the inference path and agrees with the train-mode forward (dropout off)
    with torch.no_grad():
        net.eval().engine = "simt"
        assert _rel(net(x), y) <= 2e-5


@pytest.mark.gpu
def test_fnblock_train_module_api_matches_reference_golden(g):
    import fn_ssl_b200 as F
    blk = F.FNblock(input_size=4, hidden_size=64, is_online=True, is_first=True)
    blk2 = F.FNblock(input_size=64, hidden_size=64, is_online=False, is_first=False)
    blk.load_state_dict({k[len("blk1_w_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("blk1_w_")})
    blk2.load_state_dict({k[len("blk2_w_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("blk2_w_")})
    blk, blk2 = _no_dropout(blk.cuda().train()), _no_dropout(blk2.cuda().train())
    xb = _randn((1, 10, 16, 4), 23).cuda().requires_grad_(True)
    y, fb, nbs = blk(xb)
    y2, fb2, nbs2 = blk2(y, fb_skip=fb, nb_skip=nbs)
    wy, wf = _randn(tuple(y2.shape), 24).cuda(), _randn(tuple(fb2.shape), 25).cuda()
    ((y2 * wy).sum() + (fb2 * wf).sum()).backward()
    assert _rel(y2, g["blk_y2"]) <= 2e-5 and _rel(fb2, g["blk_fb2"]) <= 2e-5
    assert _rel(xb.grad, g["blk_dx"]) <= TOL
    for n, p_ in blk.named_parameters():
        assert _rel(p_.grad, g[f"blk1_grad_{n}"]) <= TOL, n
    for n, p_ in blk2.named_parameters():
        assert _rel(p_.grad, g[f"blk2_grad_{n}"]) <= TOL, n


@pytest.mark.gpu
def test_training_steps_reduce_the_loss_with_dropout_on():
    """A few optimiser steps of the whole pipeline in train mode (default dropout 0.2): DP-IPD targets from the CUDA target
    kernel, MSE loss kernel, backward kernels, torch.optim.Adam -- the loss on the fixed batch goes down, every gradient is finite."""
    import fn_ssl_b200 as F
    from fn_ssl_b200 import training as T
    torch.manual_seed(0)
    net = F.FN_SSL(is_online=False).cuda().train()
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    sig = orc.white_noise(2, 512 + 256 * 23, 2, seed=3).cuda()
    feats = F.data_preprocess_fnssl(sig)[0]                                         # (2, 4, 256, 24) reference layout
    doa = torch.stack((torch.full((2, 2, 1), 1.5), torch.tensor([[[0.3], [0.35]], [[2.0], [2.1]]])), dim=2).cuda()   # (nb, nt2, 2, ns)
    tgt = T.dpipd_targets(doa, [[-0.04, 0, 0], [0.04, 0, 0]], ch_mode="MM")         # (nb, nt2, 512, 1)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = T.ipd_mse_loss(net(feats), tgt)
        loss.backward()
        assert all(torch.isfinite(p_.grad).all() for p_ in net.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


# ---- IPDnet: causal conv block and network training step ---------------------------------------------------------------------

@pytest.fixture(scope="module")
def gi():
    return np.load(os.path.join(ROOT, "tests", "golden", "grad_ipdnet_golden.npz"))


IPD_CFGS = [("d2", dict(input_size=4, hidden_size=128, max_track=2, is_online=True)),
            ("off6", dict(input_size=6, hidden_size=64, max_track=2, is_online=False))]


@pytest.mark.parametrize("tag,kw", IPD_CFGS)
def test_oracle_autograd_matches_reference_ipdnet_gradients(gi, tag, kw):
    x = _randn((1, kw["input_size"], 40, 25), 31)
    sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_ipdnet_state_dict(4, **kw).items()}
    y = orc.ipdnet_forward(x, sd, is_online=kw["is_online"], fast=True)
    loss = torch.nn.functional.mse_loss(y, _randn(tuple(y.shape), 32).tanh())
    loss.backward()
    _check_against_golden(gi, tag, {k: v.grad for k, v in sd.items()}, y.detach(), loss.detach())


@pytest.mark.gpu
@pytest.mark.parametrize("c0,ld0,c1,ld1,O", [(5, 5, 0, 0, 7), (20, 24, 3, 4, 128), (128, 128, 0, 0, 130), (6, 8, 130, 132, 4)])
def test_conv3x3_causal_gradients_match_torch_autograd(c0, ld0, c1, ld1, O):
    from fn_ssl_b200 import training as T
    nb, nt, nf = 2, 13, 9
    s0, s1 = _randn((nb, nt, nf, ld0), 50), (_randn((nb, nt, nf, ld1), 51) if c1 else None)
    w, dy = 0.2 * _randn((O, c0 + c1, 3, 3), 52), _randn((nb, nt, nf, O), 53)
    r0, rw = s0.clone().requires_grad_(True), w.clone().requires_grad_(True)
    r1 = s1.clone().requires_grad_(True) if c1 else None
    xin = r0[..., :c0] if not c1 else torch.cat((r0[..., :c0], r1[..., :c1]), dim=-1)
    yref = torch.nn.functional.conv2d(xin.permute(0, 3, 2, 1), rw, padding=(1, 2))[:, :, :, :-2].permute(0, 3, 2, 1)   # (nb,C,F,T) conv
    (yref * dy).sum().backward()
    d0, dw = s0.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    d1 = s1.cuda().requires_grad_(True) if c1 else None
    y = T.conv3x3_causal(d0, c0, d1, c1, dw)
    (y * dy.cuda()).sum().backward()
    assert _rel(y, yref) <= 2e-5
    assert _rel(d0.grad, r0.grad) <= TOL and _rel(dw.grad, rw.grad) <= TOL
    if c1:
        assert _rel(d1.grad, r1.grad) <= TOL


@pytest.mark.gpu
def test_causcnn_block_training_matches_reference_golden(gi):
    import fn_ssl_b200 as F
    cnn = F.CausCnnBlock(inp_dim=20, out_dim=4, cnn_hidden_dim=128)
    cnn.load_state_dict({k[len("cnn_sd."):]: torch.from_numpy(gi[k]) for k in gi.files if k.startswith("cnn_sd.")})
    cnn = cnn.cuda().train()
    xc = _randn((2, 20, 13, 38), 33).cuda().requires_grad_(True)
    yc = cnn(xc)
    (yc * _randn(tuple(yc.shape), 34).cuda()).sum().backward()
    assert _rel(yc, gi["cnn_y"]) <= 2e-5 and _rel(xc.grad, gi["cnn_dx"]) <= TOL
    for n, p_ in cnn.named_parameters():
        assert _rel(p_.grad, gi["cnn_grad." + n]) <= TOL, n
    with torch.no_grad():                                    # eval mode on the same module = the fused inference kernels
        assert _rel(cnn.eval()(xc.detach()), yc) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", IPD_CFGS)
def test_ipdnet_training_step_matches_reference_golden(gi, tag, kw):
    import fn_ssl_b200 as F
    net = F.IPDnet(**kw)
    net.load_state_dict(orc.seeded_ipdnet_state_dict(4, **kw))
    net = _no_dropout(net.cuda().train())
    x = _randn((1, kw["input_size"], 40, 25), 31).cuda()
    y = net(x)
    loss = torch.nn.functional.mse_loss(y, _randn(tuple(y.shape), 32).tanh().cuda())
    loss.backward()
    torch.cuda.synchronize()
    _check_against_golden(gi, tag, {n: p_.grad for n, p_ in net.named_parameters()}, y.detach(), loss.detach())


@pytest.mark.gpu
def test_pit_loss_gradient_matches_oracle_autograd():
    from fn_ssl_b200 import training as T
    from oracle import training_oracle as tro
    gt = _randn((2, 5, 64, 2, 3), 60)
    order = torch.stack([torch.randperm(3, generator=torch.Generator().manual_seed(70 + r)) for r in range(10)])
    pred = torch.stack([gt.reshape(10, -1, 3)[r][:, order[r]] for r in range(10)]).reshape(gt.shape) + 0.1 * _randn(tuple(gt.shape), 61)
    pr = pred.clone().requires_grad_(True)
    lref, pref = tro.ipdnet_pit_loss(pr, gt)
    lref.backward()
    pd = pred.cuda().requires_grad_(True)
    loss, perm = T.ipd_pit_mse_loss(pd, gt.cuda())
    (3.0 * loss).backward()
    assert np.array_equal(perm.cpu().numpy(), pref.numpy())
    assert abs(float(loss.detach()) - float(lref.detach())) <= 1e-5 * float(lref.detach())
    assert _rel(pd.grad, 3.0 * pr.grad) <= 1e-5


# ---- the reference's LightningModule surface (training_step) -----------------------------------------------------------------

def _train_batch():
    gen = torch.Generator().manual_seed(5)
    sig = orc.white_noise(2, 512 + 256 * 23, 2, seed=3)
    doa = torch.stack((torch.rand(2, 2, 1, generator=gen) * 3.1, torch.rand(2, 2, 1, generator=gen) * 3.1), dim=2)     # (nb, nseg, 2, ns)
    vad = torch.rand(2, 2, 12, 1, generator=gen)
    return sig, {'doa': doa, 'vad_sources': vad}


def test_training_module_surface():
    from fn_ssl_b200 import training as T
    mod = T.FNSSLTrainModule()
    for name in ("forward", "data_preprocess", "cal_loss", "training_step", "validation_step", "predict_step", "configure_optimizers"):
        assert callable(getattr(mod, name))                                          # MyModel's names, FN-SSL/Lightning/main.py:80-279
    cfg = mod.configure_optimizers()
    assert isinstance(cfg['optimizer'], torch.optim.Adam) and cfg['optimizer'].defaults['lr'] == 0.001
    assert isinstance(cfg['lr_scheduler']['scheduler'], torch.optim.lr_scheduler.ExponentialLR) and cfg['lr_scheduler']['monitor'] == 'valid/loss'
    assert list(mod.state_dict().keys())[0].startswith("arch.")                     # checkpoints carry the `arch.` prefix, as Lightning's
    with pytest.raises(RuntimeError, match="CUDA"):
        mod.training_step(_train_batch(), 0)                                         # no CPU fallback here either


@pytest.mark.gpu
def test_training_module_step_matches_the_oracle_composition():
    """training_step = CUDA front end -> train-mode network -> CUDA DP-IPD targets -> CUDA MSE loss; against the same chain of
    oracle functions (main.py:95-103,191-266), loss and gradients."""
    import fn_ssl_b200 as F
    from fn_ssl_b200 import training as T
    from oracle import training_oracle as tro
    net = F.FN_SSL(is_online=False)
    net.load_state_dict(orc.seeded_fnssl_state_dict(0, is_online=False))
    mod = T.FNSSLTrainModule(arch=_no_dropout(net)).cuda().train()
    sig, gt = _train_batch()
    loss = mod.training_step((sig, gt), 0)["loss"]
    sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_fnssl_state_dict(0, is_online=False).items()}
    ref_gt = tro.fnssl_targets(gt['doa'].numpy(), gt['vad_sources'].mean(2).numpy(), np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), 'MM')
    ref_loss = tro.fnssl_loss(orc.fnssl_forward(orc.preprocess_fnssl(sig), sd, fast=True), ref_gt)
    assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 1e-4 * float(ref_loss.detach())
    loss.backward()
    ref_loss.backward()
    for n, p_ in net.named_parameters():
        assert _rel(p_.grad, sd[n].grad) <= 2e-4, n
    cfg = mod.configure_optimizers()
    cfg['optimizer'].step()
    cfg['lr_scheduler']['scheduler'].step()
    assert float(mod.validation_step((sig, gt))) != float(loss.detach())             # the parameters moved
    with torch.no_grad():
        assert tuple(mod.eval().predict_step(sig.permute(0, 2, 1).cuda()).shape) == (2, 2, 512)


def _ipdnet_train_batch():
    gen = torch.Generator().manual_seed(6)
    sig = orc.white_noise(2, 512 + 256 * 23, 3, seed=3)
    dp = 0.3 * torch.randn(2, sig.shape[1], 3, 2, generator=gen)
    dp[0, :, :, 1] = 0.0                                                             # source 1 of utterance 0 is silent: non-source target
    doa = torch.stack((torch.full((2, 2, 2), 1.57), torch.rand(2, 2, 2, generator=gen) * 3.1), dim=2)     # (nb, nt2, 2, nsrc)
    return sig, {'doa': doa, 'dp_signal': dp}


IPD_MIC3 = np.array(((0.0, 0.0, 0.0), (0.03, 0.0, 0.0), (0.0, 0.04, 0.0)))


def test_ipdnet_training_module_surface():
    from fn_ssl_b200 import training as T
    mod = T.IPDnetTrainModule(mic_pos=torch.tensor(IPD_MIC3))
    cfg = mod.configure_optimizers()
    assert cfg['optimizer'].defaults['lr'] == 0.0005 and isinstance(cfg['lr_scheduler']['scheduler'], torch.optim.lr_scheduler.ExponentialLR)
    assert tuple(mod.non_source_tar.shape) == (512, 2) and "non_source_tar" not in mod.state_dict()
    assert all(k.startswith("arch.") for k in mod.state_dict())
    with pytest.raises(RuntimeError, match="CUDA"):
        mod.training_step(_ipdnet_train_batch(), 0)


@pytest.mark.gpu
def test_ipdnet_training_module_step_matches_the_oracle_composition():
    """training_step of IPDnet = CUDA front end + dp-VAD (CUDA STFTs) -> train-mode network -> per-source targets with the
    non-source target -> frame-level PIT loss; against the reference's lines restated with oracle functions
    (runIPDnetOn.py:144-154,196-206,224-283), loss and gradients."""
    import fn_ssl_b200 as F
    from fn_ssl_b200 import training as T
    from oracle import training_oracle as tro
    kw = dict(input_size=6, hidden_size=64, max_track=2, is_online=True)
    net = F.IPDnet(**kw)
    net.load_state_dict(orc.seeded_ipdnet_state_dict(4, **kw))
    mod = T.IPDnetTrainModule(mic_pos=IPD_MIC3, arch=_no_dropout(net)).cuda().train()
    sig, scene = _ipdnet_train_batch()
    loss = mod.training_step((sig, scene), 0)["loss"]
    st = orc.stft(sig)
    vad = torch.zeros(2, st.shape[2], 2)
    for s in range(2):                                                               # cal_vad, :224-235
        vad[:, :, s] = torch.mean(torch.abs(orc.stft(scene['dp_signal'][:, :, :, s]))[:, :, :, 0] / torch.abs(st[:, :, :, 0]), dim=1)
    vad = torch.nn.AvgPool2d(kernel_size=(12, 1))(vad)
    tgt = tro.ipdnet_targets(scene['doa'].numpy(), vad.numpy(), IPD_MIC3, T.non_source_target(IPD_MIC3), 'M')
    sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_ipdnet_state_dict(4, **kw).items()}
    ref_loss, _ = tro.ipdnet_pit_loss(orc.ipdnet_forward(orc.preprocess_ipdnet(sig), sd, is_online=True, fast=True), tgt)
    assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 1e-4 * float(ref_loss.detach())
    loss.backward()
    ref_loss.backward()
    # 1.6 M ReLU pre-activations behind conv1: two fp32 implementations may disagree on the sign of one that is ~1e-7 from zero, which
    # moves every gradient upstream of it by ~1e-3 of its largest element (measured under the host emulation: one such flip).  The
    # layer-level tests hold the kernels to 1e-4 on smooth paths; this composition test therefore bounds the relative L2 error.
    for n, p_ in net.named_parameters():
        a, b = p_.grad.detach().cpu().double(), sd[n].grad.double()
        assert float((a - b).norm() / b.norm()) <= 5e-3, n
    cfg = mod.configure_optimizers()                                                 # keep the scheduler alive: it wraps optimizer.step
    cfg['optimizer'].step()
    assert float(mod.validation_step((sig, scene))) != float(loss.detach())


@pytest.mark.gpu
@pytest.mark.parametrize("first,nc", [(True, 6), (False, 70)])
def test_ipdnet_fnblock_train_module_api_matches_oracle_autograd(first, nc):
    """IPDnet's FNblock through its module API in train mode (FixedAarryIPDnet.py:29-40): output and the gradient w.r.t. the
    block input against the oracle's autograd."""
    from fn_ssl_b200.FixedAarryIPDnet import FNblock
    torch.manual_seed(1)
    blk = _no_dropout(FNblock(input_size=6 if first else 64, hidden_size=64, add_skip_dim=6, is_online=True, is_first=first))
    nb, nt, nf = 2, 5, 7
    x, raw, wy = _randn((nb, nt, nf, nc), 80), _randn((nb, nt, nf, 6), 81), _randn((nb, nt, nf, 70), 82)
    fb, nbs = raw.reshape(nb * nt, nf, 6), raw.permute(0, 2, 1, 3).reshape(nb * nf, nt, 6)
    sd = {("b." + k): v.clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    (orc.ipdnet_block(xr, sd, "b.", fb, nbs, fast=True) * wy).sum().backward()
    blk = blk.cuda().train()
    xd = x.cuda().requires_grad_(True)
    y = blk(xd, fb.cuda(), nbs.cuda())
    (y * wy.cuda()).sum().backward()
    assert _rel(y, orc.ipdnet_block(x, {k: v.detach() for k, v in sd.items()}, "b.", fb, nbs, fast=True)) <= 2e-5
    assert _rel(xd.grad, xr.grad) <= TOL
    for n, p_ in blk.named_parameters():
        assert _rel(p_.grad, sd["b." + n].grad) <= TOL, n


@pytest.mark.gpu
def test_fnssl_doa_head_training_step_matches_oracle_autograd():
    """`is_doa=True` (BASELINE configs[3]'s head) in train mode: the DOA classifier Linear(512,180) runs the head.cu kernels forward
    and backward (fnssl_linear_backward); every gradient against the oracle's autograd."""
    import fn_ssl_b200 as F
    kw = dict(is_online=True, is_doa=True)
    net = F.FN_SSL(**kw)
    net.load_state_dict(orc.seeded_fnssl_state_dict(3, **kw))
    net = _no_dropout(net.cuda().train())
    x, tgt = _randn((2, 4, 256, 24), 95), _randn((2, 2, 180), 96)
    sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_fnssl_state_dict(3, **kw).items()}
    yref = orc.fnssl_forward(x, sd, fast=True)
    torch.nn.functional.mse_loss(yref, tgt).backward()
    y = net(x.cuda())
    torch.nn.functional.mse_loss(y, tgt.cuda()).backward()
    assert tuple(y.shape) == (2, 2, 180) and _rel(y, yref) <= 2e-5
    for n, p_ in net.named_parameters():
        assert _rel(p_.grad, sd[n].grad) <= TOL, n
