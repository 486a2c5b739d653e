"""In-kernel timeline of the CTA-pair LSTM kernel (lstm_tc5.cu, FNSSL_TC_TRACE=1): SM-clock stamps of cluster 0 / rank 0 for
half-slots 32..47 (n = 4 t + 2 chain + unit half)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FNSSL_TC_TRACE"] = "1"
os.environ["FNSSL_TC_PAIR"] = "1"
os.environ["FNSSL_TC_PAIR_MIN"] = "1"
import torch  # noqa: E402
from fn_ssl_b200 import _lib  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

NB = int(os.environ.get("TRACE_NB", "64"))
cfgs = [("full in16", 0, NB, 249, 256, 16, 0, True, False), ("full in256 add", 0, NB, 249, 256, 256, 0, True, True)]
names = {0: "h-iss: top", 1: "h-iss: XP_DONE passed", 2: "h-iss: H_FULL passed", 3: "h-iss: H_MATE passed", 4: "h-iss: ACC_FULL committed",
         5: "epi: top", 6: "epi: ACC_FULL passed", 7: "epi: math done", 8: "epi: handed to publisher",
         9: "x-iss: pass top", 10: "x-iss: buffers free", 11: "x-iss: pass issued", 12: "pub: top", 13: "pub: pushes issued", 14: "pub: stores issued", 15: "epi: TMEM loads done"}
for name, axis, nb, nt, nf, c0, c1, bidir, add in cfgs:
    torch.manual_seed(0)
    p = LSTMParams(c0 + c1, 128, bidirectional=bidir).cuda()
    g0 = torch.randn(nb, nt, nf, c0, device="cuda").half()
    oc = 256
    for _ in range(2):
        ga = torch.randn(nb, nt, nf, oc, device="cuda").half() if add else None
        run_lstm(p, "tcgen05", axis, g0, c0, None, 0, addend=ga, inplace_addend=add)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 256)()
    if not _lib.load().fnssl_lstm_tc4_trace(buf):
        print("no trace"); continue
    tr = [[buf[s * 16 + k] for k in range(16)] for s in range(16)]
    t0 = tr[0][5]
    print(f"== {name}: epilogue half-slot period = {[tr[s + 1][5] - tr[s][5] for s in range(15)]}")
    for k in (9, 10, 11, 0, 1, 2, 3, 4, 5, 6, 15, 7, 8, 12, 13, 14):
        print(f"   {names[k]:28s}" + "".join(f" n{s + 32}:{(tr[s][k] - t0) if tr[s][k] else 0:6d}" for s in range(0, 9)))
