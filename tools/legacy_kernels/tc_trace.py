"""Print the in-kernel timeline of the cluster LSTM kernel (FNSSL_TC_TRACE=1): SM-clock deltas per step."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FNSSL_TC_TRACE"] = "1"
import torch  # noqa: E402
from fn_ssl_b200 import _lib, ops  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

cfgs = [("full  in256 H128 x2 add", 0, 16, 249, 256, 256, 0, 128, True, True),
        ("full  in16  H128 x2    ", 0, 16, 249, 256, 16, 0, 128, True, False),
        ("narrow in256+16 H128 x2 add", 1, 16, 249, 256, 256, 16, 128, True, True),
        ("narrow in256 H256 x1 add", 1, 16, 249, 256, 256, 0, 256, False, True)]
names = ["mma:step start", "mma:wait h_full", "mma:h_full passed", "mma:x-part(t+1) issued", "mma:acc_full committed",
         "x-part: cycles waiting X_FULL", "x-part: cycles issuing MMA+commit", "",
         "epi:iter start", "epi:acc_full passed", "epi:tmem loaded", "epi:math done", "epi:st.async issued", "epi:iter end"]
for name, axis, nb, nt, nf, c0, c1, H, bidir, add in cfgs:
    torch.manual_seed(0)
    p = LSTMParams(c0 + c1, H, bidirectional=bidir).cuda()
    g0 = torch.randn(nb, nt, nf, c0, device="cuda").half()
    g1 = torch.randn(nb, nt, nf, c1, device="cuda").half() if c1 else None
    oc = H * (2 if bidir else 1)
    ga = torch.randn(nb, nt, nf, oc, device="cuda").half() if add else None
    for _ in range(2):
        run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 160)()
    if not _lib.load().fnssl_lstm_tc_trace(buf):
        print("no trace"); continue
    tr = [[buf[s * 16 + k] for k in range(16)] for s in range(8)]
    print(f"== {name}: period (epi iter start to next) = {[tr[s+1][8]-tr[s][8] for s in range(7)]}")
    s = 4
    t0 = tr[s][9]
    print("   x-part of steps 8..15: waiting on X_FULL", [tr[i][5] for i in range(8)], " issuing", [tr[i][6] for i in range(8)])
    print("   per-warp st.async issue (step 12, rel.):", [buf[128 + w] - t0 for w in range(2, 18)])   # reference: epilogue passes ACC_FULL of step s
    for k in (0, 1, 2, 4, 3, 8, 9, 10, 11, 12, 13):
        print(f"   {names[k]:28s} step{s + 8}: {tr[s][k] - t0:7d}    step{s + 9}: {tr[s + 1][k] - t0:7d}")
