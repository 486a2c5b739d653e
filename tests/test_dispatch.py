"""Which tensor-core LSTM kernel serves which layer (fnssl_lstm_tc_kernel_for: the dispatcher's decision without a launch, so it
runs on a machine without a GPU).  Shapes are the layers of BASELINE.json's configurations; DESIGN.md section 4.2 states the
policy: lstm_tc5.cu for H = 128 layers with at least ~one wave of 512-row cluster tiles, lstm_tc6.cu for H = 256 layers by wave
count and for mid-size two-source H = 128 layers, lstm_tc4.cu for small grids and carried state."""
import ctypes as C

import pytest

from fn_ssl_b200 import _lib

F16, ALONG_FREQ, ALONG_TIME, TCGEN05 = 1, 0, 1, 1
NT, NF = 249, 256            # 4 s at 16 kHz: 249 STFT frames x 256 bins


def _args(axis, nb, hidden, dirs, c0, c1=0, residual=False, state=False):
    a = _lib.LstmArgs()
    a.engine, a.axis, a.nb, a.nt, a.nf = TCGEN05, axis, nb, NT, NF
    a.hidden, a.num_dirs, a.dtype = hidden, dirs, F16
    a.src0, a.c0, a.ld0 = 0x10000, c0, c0             # (pointers are never dereferenced by the query)
    if c1:
        a.src1, a.c1, a.ld1 = 0x20000, c1, c1
    a.weights, a.weights_bytes = 0x30000, 0
    a.out0, a.out0_ld, a.out0_off = 0x40000, hidden * dirs, 0
    if residual:                                       # the in-place residual output of the FN blocks (out1 == addend)
        a.addend, a.addend_ld, a.out1, a.out1_ld = 0x50000, hidden * dirs, 0x50000, hidden * dirs
    if state:
        a.h_state, a.c_state, a.state_flags = 0x60000, 0x70000, 3
    return a


def _kernel(**kw):
    lib = _lib.load()
    assert lib.fnssl_abi_version() == _lib.ABI_VERSION
    return lib.fnssl_lstm_tc_kernel_for(C.byref(_args(**kw)))


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ("FNSSL_TC_PAIR", "FNSSL_TC_PAIR_MIN", "FNSSL_TC_PAIR256", "FNSSL_TC_PAIR256_MIN", "FNSSL_TC_PAIR128_MIN"):
        monkeypatch.delenv(k, raising=False)


@pytest.mark.parametrize("nb", [256, 128, 64, 32])
def test_cfg4_shards_run_the_h128_pair_kernel(nb):
    """BASELINE configs[3], global batch 256 on 1 / 2 / 4 / 8 GPUs: every BLSTM(2x128) layer of a shard has >= 30 cluster tiles."""
    assert _kernel(axis=ALONG_FREQ, nb=nb, hidden=128, dirs=2, c0=16) == 5                        # block 1, full band
    assert _kernel(axis=ALONG_TIME, nb=nb, hidden=128, dirs=2, c0=256, c1=16, residual=True) == 5  # block 1, narrow band (+ raw features)
    assert _kernel(axis=ALONG_FREQ, nb=nb, hidden=128, dirs=2, c0=256, residual=True) == 5
    assert _kernel(axis=ALONG_TIME, nb=nb, hidden=128, dirs=2, c0=256, residual=True) == 5


def test_cfg2_layers():
    """configs[1], 16 utterances: 32 cluster tiles of 256 rows -- below lstm_tc5's threshold; the two-source layer goes to
    lstm_tc6<128> (lstm_tc4's x ring is a stage short there), the others stay on lstm_tc4."""
    assert _kernel(axis=ALONG_FREQ, nb=16, hidden=128, dirs=2, c0=16) == 4
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=128, dirs=2, c0=256, c1=16, residual=True) == 6
    assert _kernel(axis=ALONG_FREQ, nb=16, hidden=128, dirs=2, c0=256, residual=True) == 4
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=128, dirs=2, c0=256, residual=True) == 4
    # online variant: uni-LSTM(256) narrow band, 16 clusters of 8 = 2 waves against lstm_tc4's 3
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=256, dirs=1, c0=256, c1=16, residual=True) == 6
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=256, dirs=1, c0=256, residual=True) == 6


def test_small_grids_and_carried_state_stay_on_the_cluster_kernel():
    """configs[0] (one utterance), the streaming API (carried (h, c)), a wide second source."""
    for nb in (1, 2, 4):
        assert _kernel(axis=ALONG_FREQ, nb=nb, hidden=128, dirs=2, c0=16) == 4
        assert _kernel(axis=ALONG_TIME, nb=nb, hidden=128, dirs=2, c0=256, c1=16, residual=True) == 4
        assert _kernel(axis=ALONG_TIME, nb=nb, hidden=256, dirs=1, c0=256, c1=16, residual=True) == 4
    assert _kernel(axis=ALONG_TIME, nb=64, hidden=256, dirs=1, c0=256, c1=16, residual=True, state=True) == 4
    assert _kernel(axis=ALONG_TIME, nb=64, hidden=128, dirs=1, c0=256, state=True) == 4
    assert _kernel(axis=ALONG_TIME, nb=64, hidden=256, dirs=1, c0=256, c1=32) == 4      # lstm_tc6 takes <= 16 extra channels


def test_h256_wave_count_policy():
    """15 clusters of 8 CTAs are co-resident: lstm_tc6 (256 rows per cluster) against lstm_tc4 (128 rows per cluster)."""
    assert _kernel(axis=ALONG_TIME, nb=7, hidden=256, dirs=1, c0=256) == 4      # 7 clusters = 1 wave either way: lstm_tc4's is shorter
    assert _kernel(axis=ALONG_TIME, nb=8, hidden=256, dirs=1, c0=256) == 6      # 1 wave against 2
    assert _kernel(axis=ALONG_TIME, nb=15, hidden=256, dirs=1, c0=256) == 6
    assert _kernel(axis=ALONG_TIME, nb=32, hidden=256, dirs=1, c0=256, c1=16) == 6      # IPDnet cfg3: 3 waves against 5
    assert _kernel(axis=ALONG_TIME, nb=256, hidden=256, dirs=1, c0=256) == 6


def test_switches_and_unsupported_shapes(monkeypatch):
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=96, dirs=1, c0=64) == 0       # not built: the fp32 kernel serves it
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=128, dirs=2, c0=20) == 0      # channel counts must be padded to 16
    monkeypatch.setenv("FNSSL_TC_PAIR", "0")
    monkeypatch.setenv("FNSSL_TC_PAIR256", "0")
    assert _kernel(axis=ALONG_FREQ, nb=256, hidden=128, dirs=2, c0=16) == 4
    assert _kernel(axis=ALONG_TIME, nb=16, hidden=128, dirs=2, c0=256, c1=16, residual=True) == 4
    assert _kernel(axis=ALONG_TIME, nb=256, hidden=256, dirs=1, c0=256) == 4
    monkeypatch.setenv("FNSSL_TC_PAIR", "1")
    monkeypatch.setenv("FNSSL_TC_PAIR_MIN", "1")
    assert _kernel(axis=ALONG_FREQ, nb=1, hidden=128, dirs=2, c0=16) == 5       # (how the GPU tests force the pair kernel on small layers)
