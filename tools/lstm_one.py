"""Run one cfg2 LSTM layer a few times (for ncu captures): python tools/lstm_one.py [kernel=4] [reps=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FNSSL_TC_KERNEL"] = sys.argv[1] if len(sys.argv) > 1 else "4"
import torch  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

nb, nt, nf, c0, H = 16, 249, 256, 256, 128
torch.manual_seed(0)
p = LSTMParams(c0, H, bidirectional=True).cuda()
g0 = torch.randn(nb, nt, nf, c0, device="cuda").half()
ga = torch.randn(nb, nt, nf, 2 * H, device="cuda").half()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    run_lstm(p, "tcgen05", 0, g0, c0, None, 0, addend=ga)
torch.cuda.synchronize()
print("done")
