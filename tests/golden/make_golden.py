"""Generate golden vectors by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Each reference sub-project is imported in its own process (they share flat module names
``Module``/``Model``/``utils_``).  Shims: empty ``matplotlib``/``soundfile`` modules (plotting / IO
only).  The script also asserts that ``oracle.seeded_*_state_dict`` reproduces the
reference modules' default-init weights bit for bit, so fixtures only store inputs' seeds and the
reference OUTPUTS (weights are re-generated from the seed at test time).
"""
import os
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _shim():
    for m in ["matplotlib", "matplotlib.pyplot", "soundfile", "webrtcvad"]:
        sys.modules[m] = types.ModuleType(m)


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def gen_fnssl():
    _shim()
    sys.path.insert(0, os.path.join(REF, "FN-SSL", "Lightning"))
    import Model as ref_model          # FN-SSL/Lightning/Model.py
    import Module as ref_module        # FN-SSL/Lightning/Module.py
    import utils_ as ref_utils         # FN-SSL/Lightning/utils_.py
    from oracle import fnssl_oracle as orc

    out = {}
    # ---- front end: STFT / AddChToBatch / forgetting_norm / data_preprocess (main.py:206-225)
    sig = _randn((2, 512 + 256 * 30 + 77, 3), 11)
    stft = ref_module.STFT(win_len=512, win_shift_ratio=0.5, nfft=512)(sig)
    out["fe_stft_re"], out["fe_stft_im"] = stft.real.numpy(), stft.imag.numpy()
    for mode in ("M", "MM"):
        st = stft.permute(0, 3, 1, 2)
        reb = ref_module.AddChToBatch(ch_mode=mode)(st)
        mag = torch.abs(reb)
        mu = ref_utils.forgetting_norm(mag)               # default sample_length=298
        mu8 = ref_utils.forgetting_norm(mag, sample_length=8)
        re = torch.real(reb) / (mu + 1e-6)
        im = torch.imag(reb) / (mu + 1e-6)
        feat = torch.cat((re, im), dim=1)[:, :, range(1, 257), :]
        out[f"fe_mu_{mode}"] = mu.numpy()
        out[f"fe_mu8_{mode}"] = mu8.numpy()
        out[f"fe_feat_{mode}"] = feat.numpy()
    # ---- network: FN_SSL online / offline / doa on a short clip (nt=26 -> 2 pooled frames)
    x = _randn((2, 4, 256, 26), 12)
    for tag, kw in (("on", dict(is_online=True)), ("off", dict(is_online=False)),
                    ("doa", dict(is_online=True, is_doa=True))):
        torch.manual_seed(3)
        net = ref_model.FN_SSL(**kw).eval()
        sd = orc.seeded_fnssl_state_dict(3, **kw)
        ref_sd = net.state_dict()
        assert list(sd.keys()) == list(ref_sd.keys()), (tag, "state_dict key order")
        for k in sd:
            assert torch.equal(sd[k], ref_sd[k]), (tag, k)
        with torch.no_grad():
            out[f"net_{tag}"] = net(x).numpy()
    # ---- one FNblock (cfg1), first and non-first, small hidden
    torch.manual_seed(5)
    blk = ref_model.FNblock(input_size=4, hidden_size=64, is_online=True, is_first=True).eval()
    xb = _randn((1, 10, 16, 4), 13)
    with torch.no_grad():
        y, fb, nbs = blk(xb)
    out["blk_first_y"], out["blk_first_fb"], out["blk_first_nb"] = y.numpy(), fb.numpy(), nbs.numpy()
    for k, v in blk.state_dict().items():
        out["blk_first_sd." + k] = v.numpy()
    torch.manual_seed(6)
    blk2 = ref_model.FNblock(input_size=64, hidden_size=64, is_online=False, is_first=False).eval()
    with torch.no_grad():
        y2, fb2, nbs2 = blk2(y, fb_skip=fb, nb_skip=nbs)
    out["blk_next_y"], out["blk_next_fb"], out["blk_next_nb"] = y2.numpy(), fb2.numpy(), nbs2.numpy()
    for k, v in blk2.state_dict().items():
        out["blk_next_sd." + k] = v.numpy()
    # ---- the same two blocks at the paper's hidden size 256 (the shapes the tensor-core engine is built for): weights are
    # NOT stored (3 MB) -- `torch.manual_seed(s); FNblock(...)` reproduces them (asserted here against the oracle's seeded
    # generator, which follows nn.LSTM's registration order)
    xb = _randn((1, 6, 8, 4), 17)
    torch.manual_seed(15)
    blk = ref_model.FNblock(input_size=4, hidden_size=256, is_online=True, is_first=True).eval()
    torch.manual_seed(15)
    sd = {}
    orc._lstm_init(sd, "fullLstm.", 4, 128, True)
    orc._lstm_init(sd, "narrLstm.", 260, 256, False)
    assert all(torch.equal(sd[k], v) for k, v in blk.state_dict().items()) and len(sd) == len(blk.state_dict())
    with torch.no_grad():
        y, fb, nbs = blk(xb)
    out["blk256_first_y"], out["blk256_first_fb"], out["blk256_first_nb"] = y.numpy(), fb.numpy(), nbs.numpy()
    torch.manual_seed(16)
    blk2 = ref_model.FNblock(input_size=256, hidden_size=256, is_online=False, is_first=False).eval()
    with torch.no_grad():
        y2, fb2, nbs2 = blk2(y, fb_skip=fb, nb_skip=nbs)
    out["blk256_next_y"], out["blk256_next_fb"], out["blk256_next_nb"] = y2.numpy(), fb2.numpy(), nbs2.numpy()
    # ---- "next" row: DPIPD templates + SourceDetectLocalize (IDL) + PredDOA.predgt2DOA
    mic3 = np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0), (0.0, 0.05, 0.01)))
    for mode in ("M", "MM"):
        d = ref_module.DPIPD(ndoa_candidate=[7, 13], mic_location=mic3, nf=33, fre_max=8000, ch_mode=mode, speed=340)
        tpl, _, cand = d()
        out[f"dec_template_{mode}_re"], out[f"dec_template_{mode}_im"] = tpl.real.astype(np.float32), tpl.imag.astype(np.float32)
        _, ipd_gt, _ = d(source_doa=np.array([[[[1.2, 0.7], [0.4, -2.0]]]]))      # (nb=1, nt=1, 2, nsource=2)
        out[f"dec_gt_{mode}_re"], out[f"dec_gt_{mode}_im"] = ipd_gt.real.astype(np.float32), ipd_gt.imag.astype(np.float32)
    d2 = ref_module.DPIPD(ndoa_candidate=[37, 73], mic_location=np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), nf=257,
                          fre_max=8000, ch_mode="MM", speed=340)
    tpl2, _, _ = d2()
    t2 = np.concatenate((tpl2.real[:, :, range(1, 257), :], tpl2.imag[:, :, range(1, 257), :]), axis=2).astype(np.float32)
    t2 = torch.from_numpy(t2[18:19, 36:73])                                       # horizontal plane, azimuth 0..pi (:708)
    cand2 = [np.linspace(np.pi / 2, np.pi / 2, 1), np.linspace(0, np.pi, 37)]
    pred = 0.7 * t2[0, 5][None, None] + 0.4 * t2[0, 30][None, None] + 0.05 * _randn((2, 6, 512, 1), 15)
    out["dec_pred_ipd"] = pred.numpy()
    for snm in ("kNum", "unkNum", "KNum"):
        sdl = ref_module.SourceDetectLocalize(max_num_sources=2, source_num_mode=snm, meth_mode="IDL")
        doa, vad, ss = sdl(pred.clone(), t2, cand2)
        out[f"dec_doa_{snm}"], out[f"dec_vad_{snm}"] = doa.numpy(), vad.numpy()
    out["dec_ss"] = ss.numpy()
    pd = ref_module.PredDOA(device="cpu")
    netout = _randn((3, 4, 512), 16).tanh()                                       # a fake (nb*P, nt//12, 2nf) network output
    pb, _ = pd.predgt2DOA(pred_batch=netout)
    out["dec_preddoa_doa"], out["dec_preddoa_vad"], out["dec_preddoa_ss"] = pb["doa"].numpy(), pb["vad_sources"].numpy(), pb["spatial_spectrum"].numpy()
    # FN_lightning wrapper key names (FN-SSL/Model.py:92-99, loaded with expandtabs because of the TabError at :61)
    src = open(os.path.join(REF, "FN-SSL", "Model.py")).read().expandtabs(8)
    ns = {"__name__": "ref_fnssl_model"}
    exec(compile(src, "FN-SSL/Model.py", "exec"), ns)
    torch.manual_seed(0)
    out["lightning_keys"] = np.array(list(ns["FN_lightning"]().state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "fnssl_golden.npz"), **out)
    print("wrote fnssl_golden.npz:", {k: v.shape for k, v in out.items() if not k.startswith('blk_') or 'sd.' not in k})


def gen_ipdnet():
    _shim()
    sys.path.insert(0, os.path.join(REF, "IPDnet"))
    import FixedAarryIPDnet as ref_ipd
    import Module as ref_module
    import utils_ as ref_utils
    from oracle import fnssl_oracle as orc

    out = {}
    # front end (runIPDnetOn.py:240-254, runIPDnetOff.py:248-251); IPDnet/Module.py STFT allocates on CPU
    sig = _randn((2, 512 + 256 * 20 + 5, 4), 21)
    stft = ref_module.STFT(win_len=512, win_shift_ratio=0.5, nfft=512)(sig)
    st = stft.permute(0, 3, 1, 2)
    mag = torch.abs(st)
    mu = ref_utils.forgetting_norm(mag, sample_length=280)
    feat_on = torch.cat((torch.real(st) / (mu + 1e-6), torch.imag(st) / (mu + 1e-6)), dim=1)[:, :, range(1, 257), :]
    mo = torch.mean(mag.reshape(mag.shape[0], -1), dim=1)
    mo = mo[:, np.newaxis, np.newaxis, np.newaxis].expand(mag.shape)
    feat_off = torch.cat((torch.real(st) / (mo + 1e-6), torch.imag(st) / (mo + 1e-6)), dim=1)[:, :, range(1, 257), :]
    out["fe_feat_on"], out["fe_feat_off"] = feat_on.numpy(), feat_off.numpy()
    # networks
    cfgs = {
        "d2": dict(input_size=4, hidden_size=128, max_track=2, is_online=True),      # default 2-mic
        "m4": dict(input_size=8, hidden_size=256, max_track=2, is_online=True),      # cfg3 4-mic
        "off": dict(input_size=4, hidden_size=128, max_track=2, is_online=False),
    }
    for tag, kw in cfgs.items():
        torch.manual_seed(4)
        net = ref_ipd.IPDnet(**kw).eval()
        sd = orc.seeded_ipdnet_state_dict(4, **kw)
        ref_sd = net.state_dict()
        assert list(sd.keys()) == list(ref_sd.keys()), (tag, list(sd.keys()), list(ref_sd.keys()))
        for k in sd:
            assert torch.equal(sd[k], ref_sd[k]), (tag, k)
        x = _randn((2, kw["input_size"], 64, 26), 22)
        with torch.no_grad():
            out[f"net_{tag}"] = net(x).numpy()
            if tag == "off":
                net.n = 12        # n_seg: 26 frames -> 3 zero-padded chunks of 12
                out["net_off_chunked"] = net(x, offline_inference=True).numpy()
    # CausCnnBlock alone
    torch.manual_seed(9)
    cnn = ref_ipd.CausCnnBlock(inp_dim=20, out_dim=4, cnn_hidden_dim=128).eval()
    xc = _randn((2, 20, 12, 37), 23)
    with torch.no_grad():
        out["cnn_y"] = cnn(xc).numpy()
    for k, v in cnn.state_dict().items():
        out["cnn_sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "ipdnet_golden.npz"), **out)
    print("wrote ipdnet_golden.npz:", {k: v.shape for k, v in out.items() if 'sd.' not in k})


def gen_train():
    """Training-side forward pieces: DP-IPD targets and losses.  The Lightning modules cannot be
    imported (pytorch_lightning / torchmetrics absent), so the reference's own DPIPD / RemoveChFromBatch classes are run and
    the few tensor lines of data_preprocess / cal_loss around them are restated here verbatim."""
    _shim()
    sys.path.insert(0, os.path.join(REF, "FN-SSL", "Lightning"))
    import Module as ref_module        # FN-SSL/Lightning/Module.py
    from copy import deepcopy
    from oracle import training_oracle as tro

    out = {}
    rng = np.random.default_rng(7)
    fre_range_used = range(1, 257)
    for tag, mic, ch_mode in (("2mic", np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), "MM"),
                              ("3mic", np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0), (0.0, 0.05, 0.01))), "MM"),
                              ("3micM", np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0), (0.0, 0.05, 0.01))), "M")):
        gd = ref_module.DPIPD(ndoa_candidate=[37, 73], mic_location=mic, nf=257, fre_max=8000, ch_mode=ch_mode, speed=340)
        nb, nt, ns = 2, 5, 2
        doa = np.stack((rng.uniform(0, np.pi, (nb, nt, ns)), rng.uniform(-np.pi, np.pi, (nb, nt, ns))), axis=2).astype(np.float32)
        vad = rng.uniform(0, 1, (nb, nt, ns)).astype(np.float32)
        vad[rng.uniform(size=vad.shape) < 0.3] = 0.0
        # ---- FN-SSL/Lightning/main.py:227-259 (gt branch of data_preprocess), tar_useVAD = True
        _, ipd_batch, _ = gd(source_doa=doa)
        ipd_batch = np.concatenate((ipd_batch.real[:, :, fre_range_used, :, :], ipd_batch.imag[:, :, fre_range_used, :, :]),
                                   axis=2).astype(np.float32)
        ipd_batch = torch.from_numpy(ipd_batch)
        vad_batch = torch.from_numpy(vad)
        nb_, nt_, nf_, nmic_, num_source = ipd_batch.shape
        th = 0
        vad_batch_copy = deepcopy(vad_batch)
        vad_batch_copy[vad_batch_copy <= th] = th
        vad_batch_copy[vad_batch_copy > 0] = 1
        vad_batch_expand = vad_batch_copy[:, :, np.newaxis, np.newaxis, :].expand(nb_, nt_, nf_, nmic_, num_source)
        per_source = ipd_batch.clone()
        ipd_sum = torch.sum(ipd_batch * vad_batch_expand, dim=-1)
        out[f"{tag}_doa"], out[f"{tag}_vad"], out[f"{tag}_mic"] = doa, vad, mic
        out[f"{tag}_ipd_gt"] = ipd_sum.numpy()
        out[f"{tag}_ipd_per_source"] = per_source.numpy()
        assert float((tro.fnssl_targets(doa, vad, mic, ch_mode) - ipd_sum).abs().max()) <= 1e-6
        # ---- cal_loss, main.py:191-198 with the reference's RemoveChFromBatch
        P = nmic_
        pred = _randn((nb * P, nt, 2 * 256), 50 + P).tanh()
        reb = ref_module.RemoveChFromBatch(ch_mode=ch_mode)(pred, nb).permute(0, 2, 3, 1)
        loss = torch.nn.functional.mse_loss(reb.contiguous(), ipd_sum.contiguous())
        out[f"{tag}_pred"], out[f"{tag}_loss"] = pred.numpy(), np.float32(loss.item())
        assert abs(float(tro.fnssl_loss(pred, ipd_sum)) - float(loss)) <= 1e-7
    # ---- IPDnet: per-source targets with the non-source (Bessel) target, IPDnet/runIPDnetOn.py:209-222,256-283.  IPDnet's own
    # DPIPD (IPDnet/Module.py) differs from FN-SSL's only in the candidate grid and in skipping the unused mic pairs of mode 'M'
    # (same arithmetic for the pairs it keeps), so the FN-SSL class imported above produces its values.
    from scipy.special import jn
    mic4 = np.array(((0.0, 0.0, 0.0), (0.03, 0.0, 0.0), (0.0, 0.04, 0.0), (-0.035, -0.02, 0.01)))
    gd = ref_module.DPIPD(ndoa_candidate=[1, 37], mic_location=mic4, nf=257, fre_max=8000, ch_mode="M", speed=340)
    nb, nt, ns = 2, 4, 2
    doa = np.stack((np.full((nb, nt, ns), np.pi / 2), rng.uniform(0, np.pi, (nb, nt, ns))), axis=2).astype(np.float32)
    dp_vad = rng.uniform(0, 0.5, (nb, nt, ns)).astype(np.float32)
    dp_vad[rng.uniform(size=dp_vad.shape) < 0.4] = 0.0005          # below the 0.001 threshold -> silent
    fre_use = list(fre_range_used)
    # euclidean_distances_to_bessel (:209-222)
    distances = np.sqrt(np.sum((mic4[1:] - mic4[0, :]) ** 2, axis=1))
    frequencies = (2 * np.pi * np.linspace(0, 8000, 257) / 340)[fre_use]
    non_source_tar = np.array([np.concatenate((jn(0, frequencies * d), np.zeros(256))) for d in distances]).T
    non_source_tar_t = torch.from_numpy(non_source_tar)
    _, ipd_batch, _ = gd(source_doa=doa)
    ipd_batch = np.concatenate((ipd_batch.real[:, :, fre_use, :, :], ipd_batch.imag[:, :, fre_use, :, :]), axis=2).astype(np.float32)
    ipd_batch = torch.from_numpy(ipd_batch)
    nb_, nt_, nf_, nmic_, nsrc = ipd_batch.shape
    vad_batch_copy = deepcopy(torch.from_numpy(dp_vad))
    th = 0.001
    vad_batch_copy[vad_batch_copy <= th] = 0
    vad_batch_copy[vad_batch_copy > th] = 1
    ipd_batch = ipd_batch * vad_batch_copy[:, :, np.newaxis, np.newaxis, :].expand(nb_, nt_, nf_, nmic_, nsrc)
    for i in range(nb_):
        for j in range(nt_):
            for k in range(nsrc):
                if (ipd_batch[i, j, :, :, k] == 0).all():
                    ipd_batch[i, j, :, :, k] = non_source_tar_t.to(ipd_batch)
    out["ipdnet_doa"], out["ipdnet_vad"], out["ipdnet_mic"] = doa, dp_vad, mic4
    out["ipdnet_non_source"] = non_source_tar
    out["ipdnet_ipd_gt"] = ipd_batch.numpy()
    assert float((tro.ipdnet_targets(doa, dp_vad, mic4, non_source_tar) - ipd_batch).abs().max()) <= 1e-6
    # frame-level PIT loss (:196-206); torchmetrics' permutation search restated by the oracle (exhaustive, first minimum)
    pred = (ipd_batch[:, :, :, :, [1, 0]] + 0.3 * _randn(tuple(ipd_batch.shape), 61))
    pred[0, 1] = ipd_batch[0, 1] + 0.3 * _randn(tuple(ipd_batch.shape[2:]), 62)        # one frame in identity order
    loss, perm = tro.ipdnet_pit_loss(pred, ipd_batch.view(nb_ * nt_, nf_, nmic_, nsrc))
    out["ipdnet_pred"], out["ipdnet_pit_loss"], out["ipdnet_pit_perm"] = pred.numpy(), np.float32(loss.item()), perm.numpy()
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)
    print("wrote train_golden.npz:", {k: np.asarray(v).shape for k, v in out.items()})


def gen_grad():
    """Training step of the reference network (row f4): the UNMODIFIED reference FN_SSL / FNblock in TRAIN mode -- with the dropout
    probability of its nn.Dropout modules set to 0 so that the step is deterministic -- forward, MSE loss against a seeded target
    (cal_loss's F.mse_loss, main.py:191-198) and loss.backward().  Stored: output, loss, per-parameter gradient norms / sums, and a
    few whole gradient tensors.  The oracle's autograd (same functional forward, seeded weights) is asserted against it here."""
    _shim()
    sys.path.insert(0, os.path.join(REF, "FN-SSL", "Lightning"))
    import Model as ref_model          # FN-SSL/Lightning/Model.py
    from oracle import fnssl_oracle as orc

    out = {}
    x = _randn((1, 4, 256, 24), 21)
    tgt = _randn((1, 2, 512), 22).tanh()
    for tag, kw in (("off", dict(is_online=False)), ("on", dict(is_online=True))):
        torch.manual_seed(3)
        net = ref_model.FN_SSL(**kw).train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        y = net(x)
        loss = torch.nn.functional.mse_loss(y, tgt)
        loss.backward()
        out[f"{tag}_out"], out[f"{tag}_loss"] = y.detach().numpy(), np.float32(loss.item())
        names = [n for n, _ in net.named_parameters()]
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_grad_norm"] = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
        out[f"{tag}_grad_sum"] = np.array([float(p.grad.double().sum()) for _, p in net.named_parameters()])
        for n in ("emb2ipd.weight", "emb2ipd.bias", "block_1.fullLstm.weight_ih_l0", "block_1.fullLstm.bias_hh_l0_reverse",
                  "block_1.narrLstm.bias_ih_l0", "block_2.narrLstm.bias_hh_l0", "block_3.fullLstm.bias_ih_l0"):
            out[f"{tag}_grad_{n}"] = dict(net.named_parameters())[n].grad.numpy()
        out[f"{tag}_grad_block_2.fullLstm.weight_hh_l0[:, :4]"] = dict(net.named_parameters())["block_2.fullLstm.weight_hh_l0"].grad[:, :4].numpy()
        # the oracle's autograd reproduces it (same seeded weights, functional forward)
        sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_fnssl_state_dict(3, **kw).items()}
        yo = orc.fnssl_forward(x, sd, fast=True)
        torch.nn.functional.mse_loss(yo, tgt).backward()
        for n, p_ in net.named_parameters():
            err = float((sd[n].grad - p_.grad).abs().max()) / max(float(p_.grad.abs().max()), 1e-30)
            assert err <= 1e-4, (tag, n, err)
    # one FNblock (hidden 64) through its module API, first and non-first, gradient w.r.t. the input too
    torch.manual_seed(5)
    blk = ref_model.FNblock(input_size=4, hidden_size=64, is_online=True, is_first=True).train()
    blk2 = ref_model.FNblock(input_size=64, hidden_size=64, is_online=False, is_first=False).train()
    for m in list(blk.modules()) + list(blk2.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    xb = _randn((1, 10, 16, 4), 23).requires_grad_(True)
    y, fb, nbs = blk(xb)
    y2, fb2, nbs2 = blk2(y, fb_skip=fb, nb_skip=nbs)
    wy, wf = _randn(tuple(y2.shape), 24), _randn(tuple(fb2.shape), 25)
    ((y2 * wy).sum() + (fb2 * wf).sum()).backward()
    out["blk_y2"], out["blk_fb2"], out["blk_dx"] = y2.detach().numpy(), fb2.detach().numpy(), xb.grad.numpy()
    for n, p_ in blk.named_parameters():
        out[f"blk1_grad_{n}"] = p_.grad.numpy()
    for n, p_ in blk2.named_parameters():
        out[f"blk2_grad_{n}"] = p_.grad.numpy()
    out["blk1_sd_names"] = np.array([n for n, _ in blk.named_parameters()])
    for n, p_ in blk.state_dict().items():
        out[f"blk1_w_{n}"] = p_.numpy()
    for n, p_ in blk2.state_dict().items():
        out[f"blk2_w_{n}"] = p_.numpy()
    np.savez_compressed(os.path.join(HERE, "grad_golden.npz"), **out)
    print("wrote grad_golden.npz:", {k: np.asarray(v).shape for k, v in out.items() if "grad_norm" in k or k.endswith("_loss")})


def gen_grad_ipdnet():
    """Training step of the reference IPDnet (row f4): the UNMODIFIED reference IPDnet (online 4-mic-style hidden 128 / 2-mic default,
    and offline) and a CausCnnBlock alone in TRAIN mode with dropout probability 0: forward, MSE loss against a seeded target,
    loss.backward().  (The reference's own loss is the frame-level PIT loss through torchmetrics, absent here: its gradient is
    checked against the oracle's restatement in tests/test_training_backward.py.)"""
    _shim()
    sys.path.insert(0, os.path.join(REF, "IPDnet"))
    import FixedAarryIPDnet as ref_ipd
    from oracle import fnssl_oracle as orc

    out = {}
    cfgs = {"d2": dict(input_size=4, hidden_size=128, max_track=2, is_online=True),
            "off6": dict(input_size=6, hidden_size=64, max_track=2, is_online=False)}
    for tag, kw in cfgs.items():
        torch.manual_seed(4)
        net = ref_ipd.IPDnet(**kw).train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        x = _randn((1, kw["input_size"], 40, 25), 31)                  # 25 frames -> 2 output frames (the 25th is dropped by pooling)
        y = net(x)
        tgt = _randn(tuple(y.shape), 32).tanh()
        loss = torch.nn.functional.mse_loss(y, tgt)
        loss.backward()
        out[f"{tag}_out"], out[f"{tag}_loss"] = y.detach().numpy(), np.float32(loss.item())
        out[f"{tag}_names"] = np.array([n for n, _ in net.named_parameters()])
        out[f"{tag}_grad_norm"] = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
        out[f"{tag}_grad_sum"] = np.array([float(p.grad.double().sum()) for _, p in net.named_parameters()])
        for n in ("conv.conv3.weight", "conv.conv2.weight", "block_1.fullLstm.weight_ih_l0", "block_2.fullLstm.bias_hh_l0_reverse",
                  "block_2.narrLstm.bias_ih_l0"):
            out[f"{tag}_grad_{n}"] = dict(net.named_parameters())[n].grad.numpy()
        out[f"{tag}_grad_conv.conv1.weight[:, :4]"] = dict(net.named_parameters())["conv.conv1.weight"].grad[:, :4].numpy()
        sd = {k: v.clone().requires_grad_(True) for k, v in orc.seeded_ipdnet_state_dict(4, **kw).items()}
        yo = orc.ipdnet_forward(x, sd, is_online=kw["is_online"], fast=True)
        torch.nn.functional.mse_loss(yo, tgt).backward()
        for n, p_ in net.named_parameters():
            err = float((sd[n].grad - p_.grad).abs().max()) / max(float(p_.grad.abs().max()), 1e-30)
            assert err <= 1e-4, (tag, n, err)
    # CausCnnBlock alone, gradient w.r.t. the input too (odd sizes: 13 bins, 38 frames -> 3 output frames)
    torch.manual_seed(9)
    cnn = ref_ipd.CausCnnBlock(inp_dim=20, out_dim=4, cnn_hidden_dim=128).train()
    xc = _randn((2, 20, 13, 38), 33).requires_grad_(True)
    yc = cnn(xc)
    wc = _randn(tuple(yc.shape), 34)
    (yc * wc).sum().backward()
    out["cnn_y"], out["cnn_dx"] = yc.detach().numpy(), xc.grad.numpy()
    for k, v in cnn.state_dict().items():
        out["cnn_sd." + k] = v.numpy()
    for k, p_ in cnn.named_parameters():
        out["cnn_grad." + k] = p_.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "grad_ipdnet_golden.npz"), **out)
    print("wrote grad_ipdnet_golden.npz:", {k: np.asarray(v).shape for k, v in out.items() if k.endswith("_out") or k.startswith("cnn_d")})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        {"fnssl": gen_fnssl, "ipdnet": gen_ipdnet, "train": gen_train, "grad": gen_grad, "grad_ipdnet": gen_grad_ipdnet}[sys.argv[1]]()
    else:
        for which in ("fnssl", "ipdnet", "train", "grad", "grad_ipdnet"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), which])
