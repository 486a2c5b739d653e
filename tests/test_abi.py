"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/fnssl_b200.h declares;
host-side logic (state_dict surface, seeded init, error behaviour) matches the reference contract."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "fnssl_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fnssl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from fn_ssl_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fnssl_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.fnssl_abi_version() == 5


def test_host_only_entry_points():
    from fn_ssl_b200 import _lib
    lib = _lib.load()
    assert lib.fnssl_stft_num_frames(64000, 512, 256) == 249          # 4 s @ 16 kHz
    assert lib.fnssl_stft_num_frames(76640, 512, 256) == 298          # FN-SSL training clip (utils_.py:9)
    assert lib.fnssl_stft_num_frames(511, 512, 256) == 0
    assert lib.fnssl_feature_rows(3, 4, _lib.PAIRS_M) == 9
    assert lib.fnssl_feature_rows(3, 4, _lib.PAIRS_MM) == 18
    assert lib.fnssl_feature_rows(3, 4, _lib.PAIRS_ALL) == 3
    assert lib.fnssl_feature_channels(4, _lib.PAIRS_ALL) == 8
    assert lib.fnssl_feature_channels(4, _lib.PAIRS_MM) == 4
    # argument validation happens before any CUDA call
    assert lib.fnssl_stft_forward(None, 1, 64000, 2, 400, 160, 400, None, None, None) != 0
    assert b"512" in lib.fnssl_last_error()
    args = _lib.LstmArgs()
    args.axis = 7
    import ctypes
    assert lib.fnssl_lstm_forward(ctypes.byref(args), None) != 0
    assert b"axis" in lib.fnssl_last_error()


def test_ipdnet2_entry_points_validate_before_any_cuda_call():
    import ctypes
    from fn_ssl_b200 import _lib
    lib = _lib.load()
    assert lib.fnssl_reflect_pad(None, 1, 100, 2, 256, None, None) != 0
    fa = _lib.SnFreqArgs()
    assert lib.fnssl_sn_freq_forward(ctypes.byref(fa), None) != 0 and b"null" in lib.fnssl_last_error()
    fa.x, fa.out = 16, 16                                   # non-null dummies: the shape checks come first
    fa.hidden, fa.squeeze, fa.groups, fa.fkernel = 192, 8, 8, 5
    assert lib.fnssl_sn_freq_forward(ctypes.byref(fa), None) != 0 and b"dim_hidden 96" in lib.fnssl_last_error()
    ta = _lib.SnTimeArgs()
    ta.x, ta.out, ta.work = 16, 16, 16
    ta.hidden, ta.d_inner, ta.d_state, ta.dt_rank, ta.d_conv = 96, 192, 16, 6, 3
    assert lib.fnssl_sn_time_forward(ctypes.byref(ta), None) != 0 and b"Mamba" in lib.fnssl_last_error()
    assert lib.fnssl_sn_head_forward(16, 1, 1, 16, 96, 16, 16, 16, 16, 30, 16, 2, 16, None) != 0    # 30 is not 2*n_src*pairs
    assert b"dim_output" in lib.fnssl_last_error()


def test_state_dict_surface_matches_reference(golden_fnssl):
    import fn_ssl_b200 as F
    from oracle import fnssl_oracle as orc
    for kw in (dict(is_online=True), dict(is_online=False), dict(is_online=True, is_doa=True)):
        torch.manual_seed(3)
        net = F.FN_SSL(**kw)
        sd = orc.seeded_fnssl_state_dict(3, **kw)              # == reference's seeded state_dict (make_golden.py)
        assert list(net.state_dict().keys()) == list(sd.keys())
        for k, v in net.state_dict().items():
            assert torch.equal(v, sd[k]), k
        net.load_state_dict(sd, strict=True)
    assert list(F.FN_lightning().state_dict().keys()) == list(golden_fnssl["lightning_keys"])
    for kw in (dict(), dict(input_size=8, hidden_size=256), dict(is_online=False)):
        torch.manual_seed(4)
        net = F.IPDnet(**kw)
        sd = orc.seeded_ipdnet_state_dict(4, kw.get("input_size", 4), kw.get("hidden_size", 128), 2, kw.get("is_online", True))
        assert list(net.state_dict().keys()) == list(sd.keys())
        for k, v in net.state_dict().items():
            assert torch.equal(v, sd[k]), k


def test_aliases_and_errors():
    import fn_ssl_b200 as F
    assert F.FullNarrowBlock is F.FNblock and F.FixedArrayIPDnet is F.IPDnet and F.CausalConv1dBlock is F.CausCnnBlock
    net = F.FN_SSL()
    with pytest.raises(RuntimeError, match="CUDA"):          # train mode = the differentiable fp32 CUDA path: no CPU fallback either
        net(torch.zeros(1, 4, 8, 24))
    with pytest.raises(RuntimeError, match="CUDA"):
        F.IPDnet()(torch.zeros(1, 4, 256, 24))
    with pytest.raises(RuntimeError, match="eval"):          # IPDnet2 has no backward kernels: inference only
        F.IPDnet2_lightning()(torch.zeros(1, 10, 256, 10))
    with pytest.raises(RuntimeError, match="CUDA"):
        net.eval()(torch.zeros(1, 4, 8, 24))                # CPU tensor: no fallback, loud failure
    with pytest.raises(RuntimeError, match="CUDA"):
        F.STFT(512, 0.5, 512)(torch.zeros(1, 4096, 2))
    with pytest.raises(Exception):
        F.AddChToBatch("XX")(torch.zeros(1, 3, 4, 5))


def test_simt_weight_packing_layout():
    from fn_ssl_b200.packing import LSTMParams, pack_lstm_simt
    torch.manual_seed(0)
    p = LSTMParams(5, 32, bidirectional=True)
    buf = pack_lstm_simt([tuple(t.detach() for t in d) for d in p.directions()])
    H, I, Kp = 32, 5, 40
    assert buf.numel() == 2 * Kp * H * 4 + 2 * H * 4
    w4 = buf[: 2 * Kp * H * 4].reshape(2, Kp, H, 4)
    b4 = buf[2 * Kp * H * 4:].reshape(2, H, 4)
    wi, wh, bi, bh = [t.detach() for t in p.directions()[1]]
    assert w4[1, 3, 7, 2] == wi[2 * H + 7, 3]                 # gate g, unit 7, input 3
    assert w4[1, I + 9, 7, 3] == wh[3 * H + 7, 9]             # gate o, unit 7, hidden 9
    assert torch.all(w4[:, I + H:] == 0)
    assert torch.allclose(b4[1, 7, 1], bi[H + 7] + bh[H + 7])


def test_addchtobatch_cpu_matches_oracle():
    # pure indexing: also valid on CPU tensors
    import fn_ssl_b200 as F
    from oracle import fnssl_oracle as orc
    x = torch.randn(2, 4, 3, 5, dtype=torch.complex64)
    for mode in ("M", "MM"):
        assert torch.equal(F.AddChToBatch(mode)(x), orc.add_ch_to_batch(x, mode))


def test_ipdnet2_state_dict_surface_matches_reference_checkpoint():
    """Key set and shapes equal the reference's shipped checkpoint (IPDnet2/checkpoints/ipdnet2_small.ckpt; its key /
    shape table is oracle.ipdnet2_param_shapes, asserted against the checkpoint by tests/golden/make_golden_ipdnet2.py)."""
    import os
    import fn_ssl_b200 as F
    from oracle import ipdnet2_oracle as orc2
    net = F.IPDnet2_lightning()
    shapes = orc2.ipdnet2_param_shapes(dim_input=10, dim_output=16, num_layers=8)
    sd = net.state_dict()
    assert set(sd) == {"arch." + k for k in shapes} and len(sd) == 326
    for k, shp in shapes.items():
        assert tuple(sd["arch." + k].shape) == tuple(shp), k
    net.arch.load_state_dict(orc2.seeded_ipdnet2_state_dict(1), strict=True)
    ck = "/root/reference/IPDnet2/checkpoints/ipdnet2_small.ckpt"      # only present in the build container
    if os.path.exists(ck):
        net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=False)["state_dict"], strict=True)
    with pytest.raises(Exception, match="configuration"):
        F.OnlineSpatialNet(dim_input=10, dim_output=16, num_layers=2, dim_squeeze=8, num_freqs=256, dim_hidden=192,
                           attention='mamba(16,4)')
    with pytest.raises(RuntimeError, match="CUDA"):
        net.eval()(torch.zeros(1, 10, 256, 10))


def test_ctypes_structures_match_the_c_header(tmp_path):
    """Compile a C program against include/fnssl_b200.h (gcc, host only) and compare sizeof / offsetof of every argument
    struct with the ctypes mirror in fn_ssl_b200/_lib.py -- a silent layout mismatch would corrupt kernel arguments."""
    import ctypes
    import shutil
    import subprocess
    from fn_ssl_b200 import _lib
    gcc = shutil.which("gcc") or shutil.which("cc")
    if not gcc:
        pytest.skip("no C compiler")
    structs = {"fnssl_lstm_args": _lib.LstmArgs, "fnssl_sn_fconv_weights": _lib.SnFconvWeights,
               "fnssl_sn_freq_args": _lib.SnFreqArgs, "fnssl_mamba_weights": _lib.MambaWeights,
               "fnssl_sn_time_args": _lib.SnTimeArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fnssl_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_library_links_from_plain_c(tmp_path):
    """The boundary is a C ABI: a C99 program links against libfnssl_b200.so with nothing but the header (no C++ name
    mangling, no torch / Python dependency) and calls the host-only entry points."""
    import shutil
    import subprocess
    from fn_ssl_b200 import _lib
    _lib.load()
    gcc = shutil.which("gcc") or shutil.which("cc")
    if not gcc:
        pytest.skip("no C compiler")
    src = tmp_path / "host.c"
    src.write_text('#include <stdio.h>\n#include "fnssl_b200.h"\nint main(void) {\n'
                   '  printf("%d %d %d\\n", fnssl_abi_version(), fnssl_stft_num_frames(64000, 512, 256), fnssl_feature_rows(3, 4, FNSSL_PAIRS_MM));\n'
                   '  if (fnssl_stft_forward(NULL, 1, 64000, 2, 400, 160, 400, NULL, NULL, NULL) == 0) return 1;\n'
                   '  printf("%s\\n", fnssl_last_error());\n  return 0;\n}\n')
    exe = tmp_path / "host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-l:libfnssl_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    first, second = res.stdout.splitlines()[:2]
    assert first.split() == ["5", "249", "18"] and "512" in second
