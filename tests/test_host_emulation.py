"""Index arithmetic and launch geometry of the CUDA-core TRAINING kernels, checked without a GPU.

tools/host_emu compiles fn_ssl_b200/csrc/{lstm_simt,lstm_train,conv_train}.cu -- kernels and their host launch code, unmodified
apart from the launch syntax -- with g++ against a shim that runs one OS thread per CUDA thread, one block at a time
(`__shared__` = static storage, `__syncthreads()` = a barrier).  The emulated entry points are called through ctypes on numpy
buffers and compared with torch's autograd on tiny shapes.  This is test infrastructure: it is not a CPU path of the product (the
package never loads it), it only lets a kernel edit be checked here before GPU time is spent on it; the parity tests proper are
the `-m gpu` tests of tests/test_training_backward.py."""
import ctypes as C
import os
import shutil
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "host_emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no host C++ compiler")


@pytest.fixture(scope="module")
def emu():
    import build as emu_build
    from fn_ssl_b200 import _lib
    lib = C.CDLL(emu_build.build())
    vp, i, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.fnssl_lstm_train_saved_bytes.restype = i64
    lib.fnssl_lstm_train_saved_bytes.argtypes = [i] * 5
    lib.fnssl_lstm_forward_train.argtypes = [C.POINTER(_lib.LstmArgs), vp, i64, vp]
    lib.fnssl_lstm_backward.argtypes = [C.POINTER(_lib.LstmArgs), vp, i64, vp, vp, i, vp, i, vp, i, vp, vp]
    lib.fnssl_conv3x3_forward.argtypes = [vp, i, i, vp, i, i, i, i, i, vp, i, vp, vp, i, vp]
    lib.fnssl_conv3x3_backward_data.argtypes = [vp, i, i, i, i, i, vp, i, i, vp, vp, i, vp, i, vp]
    lib.fnssl_conv3x3_backward_weight.argtypes = [vp, i, i, vp, i, i, vp, i, i, i, i, i, vp, vp, vp]
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_launch_log.restype = C.c_char_p
    return lib


def _randn(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data


def _np(t):
    return np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b.detach() if torch.is_tensor(b) else b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _sequences(grid, axis):
    nb, nt, nf, Cc = grid.shape
    return grid.reshape(nb * nt, nf, Cc) if axis == 0 else grid.permute(0, 2, 1, 3).reshape(nb * nf, nt, Cc)


def _grid(seq, axis, nb, nt, nf):
    return seq.reshape(nb, nt, nf, -1) if axis == 0 else seq.reshape(nb, nf, nt, -1).permute(0, 2, 1, 3)


@pytest.mark.parametrize("rpt", ["8", "16"])
@pytest.mark.parametrize("axis", [0, 1])
def test_emulated_lstm_rows_per_thread_variants(emu, monkeypatch, rpt, axis):
    """The 8- and 16-rows-per-thread instantiations of the training forward and the BPTT kernel (picked by grid coverage at
    batch >= 8 / 16) forced onto a small grid with a partial row tile."""
    monkeypatch.setenv("FNSSL_TRAIN_RPT", rpt)
    test_emulated_lstm_layer_backward(emu, monkeypatch, "2", axis, 64, True, 8, 8, 4, 4, (2, 9, 11))
    rows = 2 * 9 if axis == 0 else 2 * 11
    blocks = -(-rows // (int(rpt) * 4))                                   # H = 64: 4 row groups per CTA
    assert f"lstm_bwd_seq_kernel<H, RPT>[{blocks},2,1]" in emu.emu_launch_log().decode(), emu.emu_launch_log().decode()


@pytest.mark.parametrize("dw_version", ["1", "2"])
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("H,bidir,c0,ld0,c1,ld1,geom", [(32, True, 4, 4, 0, 0, (2, 5, 7)), (32, False, 8, 12, 4, 4, (2, 5, 7)),
                                                       (64, True, 5, 6, 3, 3, (2, 5, 7)),
                                                       (32, False, 8, 12, 4, 4, (1, 30, 41))])   # 1230 positions: the reduction is split
def test_emulated_lstm_layer_backward(emu, monkeypatch, dw_version, axis, H, bidir, c0, ld0, c1, ld1, geom):
    from fn_ssl_b200 import _lib
    from fn_ssl_b200.packing import pack_lstm_simt, pack_lstm_whh_t, unpack_lstm_simt_grad
    monkeypatch.setenv("FNSSL_TRAIN_DW", dw_version)      # (5, 6, 3, 3) is not a multiple-of-4 shape: it takes kernel 1 either way
    nb, nt, nf = geom
    dirs_n = 2 if bidir else 1
    torch.manual_seed(11)
    ref = torch.nn.LSTM(c0 + c1, H, batch_first=True, bidirectional=bidir)
    s0, s1 = _randn((nb, nt, nf, ld0), 1), (_randn((nb, nt, nf, ld1), 2) if c1 else None)
    dout = _randn((nb, nt, nf, H * dirs_n), 3)
    r0 = s0.clone().requires_grad_(True)
    r1 = s1.clone().requires_grad_(True) if c1 else None
    xin = r0[..., :c0] if not c1 else torch.cat((r0[..., :c0], r1[..., :c1]), dim=-1)
    href = _grid(ref(_sequences(xin, axis))[0], axis, nb, nt, nf)
    (href * dout).sum().backward()

    ps = [p.detach() for p in ref.parameters()]
    dirs = [tuple(ps[4 * d:4 * d + 4]) for d in range(dirs_n)]
    w, whh_t = _np(pack_lstm_simt(dirs)), _np(pack_lstm_whh_t(dirs))
    a0, a1, dy = _np(s0), (_np(s1) if c1 else None), _np(dout)
    out = np.zeros((nb, nt, nf, H * dirs_n), np.float32)
    nbytes = emu.fnssl_lstm_train_saved_bytes(nb, nt, nf, H, dirs_n)
    saved = np.zeros(nbytes // 4 + 4, np.float32)
    saved_p = (saved.ctypes.data + 15) & ~15
    a = _lib.LstmArgs()
    a.engine, a.axis, a.nb, a.nt, a.nf, a.hidden, a.num_dirs, a.dtype = 0, axis, nb, nt, nf, H, dirs_n, 0
    a.src0, a.c0, a.ld0 = a0.ctypes.data, c0, ld0
    a.src1, a.c1, a.ld1 = (a1.ctypes.data, c1, ld1) if c1 else (None, 0, 0)
    a.weights, a.weights_bytes = w.ctypes.data, w.nbytes
    a.out0, a.out0_ld, a.out0_off = out.ctypes.data, H * dirs_n, 0
    assert emu.fnssl_lstm_forward_train(C.byref(a), saved_p, nbytes, None) == 0, emu.emu_last_error()
    assert _rel(out, href) <= 2e-5
    d0, d1 = np.zeros_like(a0), (np.zeros_like(a1) if c1 else None)
    dw = np.full_like(w, 7.0)                              # the call zeroes it
    emu.emu_launch_log_reset()
    assert emu.fnssl_lstm_backward(C.byref(a), saved_p, nbytes, whh_t.ctypes.data, dy.ctypes.data, dy.shape[-1], d0.ctypes.data, ld0,
                                   _ptr(d1), ld1, dw.ctypes.data, None) == 0, emu.emu_last_error()
    log = emu.emu_launch_log().decode()
    vec = c0 % 4 == 0 and c1 % 4 == 0 and ld0 % 4 == 0 and (c1 == 0 or ld1 % 4 == 0)
    assert ("lstm_bwd_dw2_kernel" in log) == (dw_version == "2" and vec) and "lstm_bwd_seq_kernel" in log and "lstm_bwd_dx_kernel" in log, log
    assert _rel(d0, r0.grad) <= 1e-4
    if c1:
        assert _rel(d1, r1.grad) <= 1e-4
    for d, g4 in enumerate(unpack_lstm_simt_grad(torch.from_numpy(dw), dirs_n, c0 + c1, H)):
        for k in range(4):
            assert _rel(g4[k], list(ref.parameters())[4 * d + k].grad) <= 1e-4, (d, k)


@pytest.mark.parametrize("c0,ld0,c1,ld1,O,geom", [(5, 5, 0, 0, 7, (2, 6, 5)), (6, 8, 3, 4, 130, (2, 6, 5)), (130, 132, 2, 2, 4, (2, 6, 5)),
                                                  (6, 8, 3, 4, 130, (2, 30, 40))])               # 2400 positions: several splits
def test_emulated_conv3x3_products(emu, c0, ld0, c1, ld1, O, geom):
    nb, nt, nf = geom
    Cc = c0 + c1
    s0, s1 = _randn((nb, nt, nf, ld0), 50), (_randn((nb, nt, nf, ld1), 51) if c1 else None)
    w, dy = 0.2 * _randn((O, Cc, 3, 3), 52), _randn((nb, nt, nf, O), 53)
    r0, rw = s0.clone().requires_grad_(True), w.clone().requires_grad_(True)
    r1 = s1.clone().requires_grad_(True) if c1 else None
    xin = r0[..., :c0] if not c1 else torch.cat((r0[..., :c0], r1[..., :c1]), dim=-1)
    yref = torch.nn.functional.conv2d(xin.permute(0, 3, 2, 1), rw, padding=(1, 2))[:, :, :, :-2].permute(0, 3, 2, 1)
    (yref * dy).sum().backward()
    a0, a1, wn, dyn = _np(s0), (_np(s1) if c1 else None), _np(w), _np(dy)
    work = np.zeros(9 * Cc * O, np.float32)
    y = np.zeros((nb, nt, nf, O), np.float32)
    assert emu.fnssl_conv3x3_forward(a0.ctypes.data, c0, ld0, _ptr(a1), c1, ld1, nb, nt, nf, wn.ctypes.data, O, work.ctypes.data,
                                     y.ctypes.data, O, None) == 0, emu.emu_last_error()
    assert _rel(y, yref) <= 2e-5
    d0, d1 = np.zeros_like(a0), (np.zeros_like(a1) if c1 else None)
    assert emu.fnssl_conv3x3_backward_data(dyn.ctypes.data, O, O, nb, nt, nf, wn.ctypes.data, c0, c1, work.ctypes.data, d0.ctypes.data,
                                           ld0, _ptr(d1), ld1, None) == 0, emu.emu_last_error()
    assert _rel(d0, r0.grad) <= 1e-4
    if c1:
        assert _rel(d1, r1.grad) <= 1e-4
    dw = np.full_like(wn, 3.0)
    assert emu.fnssl_conv3x3_backward_weight(a0.ctypes.data, c0, ld0, _ptr(a1), c1, ld1, dyn.ctypes.data, O, O, nb, nt, nf,
                                             work.ctypes.data, dw.ctypes.data, None) == 0, emu.emu_last_error()
    assert _rel(dw, rw.grad) <= 1e-4


# ---- train.cu: DP-IPD targets and the losses against the reference goldens ----------------------------------------------------

@pytest.fixture(scope="module")
def tg():
    return np.load(os.path.join(ROOT, "tests", "golden", "train_golden.npz"))


@pytest.mark.parametrize("tag,mode", [("2mic", "MM"), ("3mic", "MM"), ("3micM", "M")])
def test_emulated_targets_and_mse_loss(emu, tg, tag, mode):
    from fn_ssl_b200.training import mic_pairs
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    emu.fnssl_dpipd_targets.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, f, f, i, i, f, i, vp, vp, vp]
    emu.fnssl_ipd_mse_loss.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    doa, vad, mic = np.ascontiguousarray(tg[f"{tag}_doa"], np.float32), np.ascontiguousarray(tg[f"{tag}_vad"], np.float32), tg[f"{tag}_mic"]
    nb, nt, _, ns = doa.shape
    pairs = np.ascontiguousarray(mic_pairs(mic.shape[0], mode))
    micf = np.ascontiguousarray(mic, np.float32)
    P = pairs.shape[0]
    out = np.zeros((nb, nt, 512, P), np.float32)
    assert emu.fnssl_dpipd_targets(doa.ctypes.data, vad.ctypes.data, micf.ctypes.data, pairs.ctypes.data, nb, nt, ns, mic.shape[0], P, 257,
                                   8000.0, 340.0, 1, 256, 0.0, 0, None, out.ctypes.data, None) == 0, emu.emu_last_error()
    assert float(np.abs(out - tg[f"{tag}_ipd_gt"]).max()) <= 2e-6
    per = np.zeros((nb, nt, 512, P, ns), np.float32)
    assert emu.fnssl_dpipd_targets(doa.ctypes.data, None, micf.ctypes.data, pairs.ctypes.data, nb, nt, ns, mic.shape[0], P, 257,
                                   8000.0, 340.0, 1, 256, 0.0, 1, None, per.ctypes.data, None) == 0, emu.emu_last_error()
    assert float(np.abs(per - tg[f"{tag}_ipd_per_source"]).max()) <= 2e-6
    pred, gt = np.ascontiguousarray(tg[f"{tag}_pred"], np.float32), np.ascontiguousarray(tg[f"{tag}_ipd_gt"], np.float32)
    ws, loss = np.zeros(nb * nt, np.float32), np.zeros(1, np.float32)
    assert emu.fnssl_ipd_mse_loss(pred.ctypes.data, gt.ctypes.data, nb, P, nt, 512, ws.ctypes.data, loss.ctypes.data, None) == 0
    assert abs(float(loss[0]) - float(tg[f"{tag}_loss"])) <= 2e-6 * float(tg[f"{tag}_loss"])


def test_emulated_pit_loss(emu, tg):
    vp, i = C.c_void_p, C.c_int
    emu.fnssl_ipd_pit_mse_loss.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp]
    pred = np.ascontiguousarray(tg["ipdnet_pred"], np.float32)
    gt = np.ascontiguousarray(tg["ipdnet_ipd_gt"], np.float32)
    nb, nt, _, _, ns = pred.shape
    rows = nb * nt
    p, g = pred.reshape(rows, -1, ns), gt.reshape(rows, -1, ns)
    ws, loss, perm = np.zeros(rows, np.float32), np.zeros(1, np.float32), np.zeros((rows, ns), np.int32)
    assert emu.fnssl_ipd_pit_mse_loss(p.ctypes.data, g.ctypes.data, rows, p.shape[1], ns, ws.ctypes.data, loss.ctypes.data,
                                      perm.ctypes.data, None) == 0, emu.emu_last_error()
    assert abs(float(loss[0]) - float(tg["ipdnet_pit_loss"])) <= 2e-6 * float(tg["ipdnet_pit_loss"])
    assert np.array_equal(perm, tg["ipdnet_pit_perm"])


# ---- head.cu: DP-IPD head forward / backward and the DOA linear (warp shuffles emulated with per-warp barriers) ------------------

@pytest.mark.parametrize("nt,Cc,ld", [(24, 64, 64), (31, 40, 44)])
def test_emulated_ipd_head_forward_backward_and_linear(emu, nt, Cc, ld):
    vp, i = C.c_void_p, C.c_int
    emu.fnssl_ipd_head_forward.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, vp]
    emu.fnssl_ipd_head_backward.argtypes = [vp, i, i, i, i, i, vp, vp, vp, vp, i, vp, vp, vp]
    emu.fnssl_linear_forward.argtypes = [vp, vp, vp, i, i, i, vp, vp]
    nb, nf = 2, 7
    x, w, b = _randn((nb, nt, nf, ld), 9), 0.1 * _randn((2, Cc), 10), _randn((2,), 11)
    dy = _randn((nb, nt // 12, 2 * nf), 12)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    seq = xr[..., :Cc].permute(0, 2, 1, 3).reshape(nb * nf, nt, Cc)                 # Model.py:79-87
    ipd = torch.tanh(torch.nn.functional.linear(torch.nn.AvgPool2d(kernel_size=(12, 1))(seq), wr, br))
    ipd = ipd.view(nb, nf, nt // 12, 2).permute(0, 2, 1, 3)
    yref = torch.cat((ipd[..., 0], ipd[..., 1]), dim=2)
    (yref * dy).sum().backward()
    xn, wn, bn, dyn = _np(x), _np(w), _np(b), _np(dy)
    y = np.zeros((nb, nt // 12, 2 * nf), np.float32)
    assert emu.fnssl_ipd_head_forward(xn.ctypes.data, 0, ld, nb, nt, nf, Cc, wn.ctypes.data, bn.ctypes.data, y.ctypes.data, None) == 0
    assert _rel(y, yref) <= 1e-5
    dx, dw, db = np.zeros_like(xn), np.full_like(wn, 5.0), np.full(2, 5.0, np.float32)
    assert emu.fnssl_ipd_head_backward(xn.ctypes.data, ld, nb, nt, nf, Cc, wn.ctypes.data, y.ctypes.data, dyn.ctypes.data, dx.ctypes.data, ld,
                                       dw.ctypes.data, db.ctypes.data, None) == 0, emu.emu_last_error()
    assert _rel(dx, xr.grad) <= 1e-4 and _rel(dw, wr.grad) <= 1e-4 and _rel(db, br.grad) <= 1e-4
    # DOA classifier Linear(2 nf -> 9) on the head's output rows
    lw, lb = _randn((9, 2 * nf), 13), _randn((9,), 14)
    rows = y.reshape(-1, 2 * nf)
    z = np.zeros((rows.shape[0], 9), np.float32)
    assert emu.fnssl_linear_forward(rows.ctypes.data, _np(lw).ctypes.data, _np(lb).ctypes.data, rows.shape[0], 2 * nf, 9, z.ctypes.data, None) == 0
    assert _rel(z, torch.nn.functional.linear(torch.from_numpy(rows), lw, lb)) <= 1e-5


def test_emulated_linear_backward(emu):
    vp, i = C.c_void_p, C.c_int
    emu.fnssl_linear_backward.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp]
    rows, inf, outf = 5, 300, 19
    x, w, dy = _randn((rows, inf), 90), _randn((outf, inf), 91), _randn((rows, outf), 92)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), torch.zeros(outf, requires_grad=True)
    (torch.nn.functional.linear(xr, wr, br) * dy).sum().backward()
    xn, wn, dyn = _np(x), _np(w), _np(dy)
    dx, dw, db = np.zeros_like(xn), np.zeros_like(wn), np.zeros(outf, np.float32)
    assert emu.fnssl_linear_backward(xn.ctypes.data, wn.ctypes.data, dyn.ctypes.data, rows, inf, outf, dx.ctypes.data, dw.ctypes.data,
                                     db.ctypes.data, None) == 0, emu.emu_last_error()
    assert _rel(dx, xr.grad) <= 1e-5 and _rel(dw, wr.grad) <= 1e-5 and _rel(db, br.grad) <= 1e-5


# ---- `-m gpu` test bodies against the emulated library (a patched child process; tools/host_emu/run_tests_on_emulator.py) --------

@pytest.mark.parametrize("args", [
    ["--module", "test_training_backward", "-k", "ipd_head_train", "-k", "pit_loss_gradient", "-k", "rows_per_thread", "-k", "ipdnet_fnblock_train"],
    ["--module", "test_gpu_ipdnet2", "-k", "single_layer"],      # IPDnet2: SpatialNetLayer (frequency stage + both Mamba blocks) vs the oracle
])
def test_gpu_test_bodies_on_the_emulator(args):
    import subprocess
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_emu", "run_tests_on_emulator.py")] + args,
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("test_")]
    assert res.returncode == 0 and lines and all(": ok" in ln for ln in lines), res.stdout[-3000:] + res.stderr[-2000:]
