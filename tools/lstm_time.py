"""Time single LSTM layers of the BASELINE configs on each tcgen05 kernel generation (CUDA events, median of N).

    python tools/lstm_time.py [kernels=4,2] [debug flags ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

CFGS = [("full   in16      H128 x2    ", 0, 16, 249, 256, 16, 0, 128, True, False),
        ("narrow in256+16  H128 x2 add", 1, 16, 249, 256, 256, 16, 128, True, True),
        ("full   in256     H128 x2 add", 0, 16, 249, 256, 256, 0, 128, True, True),
        ("narrow in256     H128 x2 add", 1, 16, 249, 256, 256, 0, 128, True, True),
        ("narrow in256+16  H256 x1 add", 1, 16, 249, 256, 256, 16, 256, False, True),
        ("narrow in256     H256 x1 add", 1, 16, 249, 256, 256, 0, 256, False, True),
        ("full   in256     H128 x2 add B=4", 0, 4, 249, 256, 256, 0, 128, True, True),
        ("full   in256     H128 x2 add B=64", 0, 64, 249, 256, 256, 0, 128, True, True)]


def main():
    kernels = (sys.argv[1] if len(sys.argv) > 1 else "4,2").split(",")
    debugs = sys.argv[2:] or ["0"]
    for name, axis, nb, nt, nf, c0, c1, H, bidir, add in CFGS:
        torch.manual_seed(0)
        p = LSTMParams(c0 + c1, H, bidirectional=bidir).cuda()
        g0 = torch.randn(nb, nt, nf, c0, device="cuda").half()
        g1 = torch.randn(nb, nt, nf, c1, device="cuda").half() if c1 else None
        oc = H * (2 if bidir else 1)
        ga = torch.randn(nb, nt, nf, oc, device="cuda").half() if add else None
        flop = 2.0 * nb * nt * nf * (2 if bidir else 1) * 4 * H * (c0 + c1 + H)
        line = f"{name}:"
        ref = None
        for k in kernels:
            for dbg in debugs:
                os.environ["FNSSL_TC_KERNEL"] = k
                os.environ["FNSSL_TC_DEBUG"] = dbg
                try:
                    for _ in range(2):
                        h, hs = run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga)
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(7):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        h, hs = run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga)
                        e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ms = sorted(ts)[len(ts) // 2]
                    extra = ""
                    if dbg == "0":
                        if ref is None:
                            ref = h.float()
                        else:
                            extra = f" (maxdiff vs first {float((h.float() - ref).abs().max()):.1e})"
                    elif ref is not None:
                        extra = f" (maxdiff vs first {float((h.float() - ref).abs().max()):.1e})"
                    line += f"  k{k}/d{dbg}: {ms:.3f} ms {flop / ms / 1e9:.0f} TF{extra}"
                except Exception as e:  # noqa: BLE001
                    line += f"  k{k}/d{dbg}: FAILED {str(e)[:80]}"
                    print(line, flush=True)
                    raise
        print(line, flush=True)


if __name__ == "__main__":
    main()
