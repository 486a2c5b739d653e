"""Print the in-kernel timeline of the two-chain cluster LSTM kernel (lstm_tc4.cu, FNSSL_TC_TRACE=1): SM-clock stamps of
CTA (0,0) for slots 16..31 (slot n = 2 t + sub)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FNSSL_TC_TRACE"] = "1"
os.environ.setdefault("FNSSL_TC_WAIT_TIMEOUT", "1")
import torch  # noqa: E402
from fn_ssl_b200 import _lib  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

cfgs = [("full  in256 H128 x2 add", 0, 16, 249, 256, 256, 0, 128, True, True),
        ("full  in16  H128 x2    ", 0, 16, 249, 256, 16, 0, 128, True, False),
        ("narrow in256 H256 x1 add", 1, 16, 249, 256, 256, 0, 256, False, True),
        ("narrow in256+16 H256 x1 add", 1, 16, 249, 256, 256, 16, 256, False, True),
        ("narrow in256+16 H128 x2 add", 1, 16, 249, 256, 256, 16, 128, True, True)]
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if any(a in c[0] for a in sys.argv[1:])]
names = {0: "h-mma: slot start", 1: "h-mma: h_full passed", 2: "h-mma: acc_full committed", 8: "x-mma: slot start",
         9: "x-mma: slot issued", 4: "epi: slot start", 5: "epi: acc_full passed", 6: "epi: math done", 7: "epi: published",
         10: "epi: wait H_FREE (cyc)", 11: "x-mma: wait ACC_EMPTY (cyc)", 12: "x-mma: wait X_FULL (cyc)"}
REL = 10   # entries below this index are absolute stamps, the rest are cycle counts
print("FNSSL_TC_DEBUG =", os.environ.get("FNSSL_TC_DEBUG", "0"))
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
    buf = (C.c_longlong * 256)()
    if not _lib.load().fnssl_lstm_tc4_trace(buf):
        print("no trace"); continue
    tr = [[buf[s * 16 + k] for k in range(16)] for s in range(16)]
    print(f"== {name}: slot period (epi slot start to next) = {[tr[s + 1][4] - tr[s][4] for s in range(15)]}")
    t0 = tr[4][4]
    for k in (8, 9, 11, 12, 0, 1, 2, 4, 5, 6, 7, 10):
        print(f"   {names[k]:30s}" + "".join(f" n{s + 16}:{tr[s][k] - (t0 if k < REL else 0):6d}" for s in range(4, 9)))
