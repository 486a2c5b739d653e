cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_q4.so
FNSSL_TC_WAIT_TIMEOUT=1 FNSSL_B200_LIB=$V timeout 300 python - <<'PY' 2>&1 | tail -8 | tee $O/r2_q4_diag_44.log
import torch, sys
sys.path.insert(0, ".")
from fn_ssl_b200 import _lib
from fn_ssl_b200.packing import LSTMParams, run_lstm
lib = _lib.load()
for nb in (16, 64, 256):
    torch.manual_seed(0)
    p = LSTMParams(16, 128, bidirectional=True).cuda()
    g0 = torch.randn(nb, 249, 256, 16, device="cuda").half()
    try:
        h, _ = run_lstm(p, "tcgen05", 0, g0, 16, None, 0)
        torch.cuda.synchronize()
        print("nb", nb, "ok", float(h.float().abs().mean()), "kernel", "pair" if nb >= 30 else "tc4")
    except Exception as e:
        print("nb", nb, "FAILED", str(e)[:100], "site", lib.fnssl_lstm_tc_error_site())
        break
PY
