"""Time single LSTM layers of the BASELINE configs on the tcgen05 kernel (CUDA events, median of N) and print a checksum of
the outputs (variants of the kernel must agree bit for bit).

    python tools/lstm_time.py [substring filters ...]           # FNSSL_B200_LIB selects a variant library
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FNSSL_TC_WAIT_TIMEOUT", "1")
import torch  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402

CFGS = [("full_in16_H128x2", 0, 16, 249, 256, 16, 0, 128, True, False),
        ("narrow_in256+16_H128x2_add", 1, 16, 249, 256, 256, 16, 128, True, True),
        ("full_in256_H128x2_add", 0, 16, 249, 256, 256, 0, 128, True, True),
        ("narrow_in256_H128x2_add", 1, 16, 249, 256, 256, 0, 128, True, True),
        ("narrow_in256+16_H256x1_add", 1, 16, 249, 256, 256, 16, 256, False, True),
        ("narrow_in256_H256x1_add", 1, 16, 249, 256, 256, 0, 256, False, True),
        ("full_in256+16_H128x2", 0, 32, 249, 256, 256, 16, 128, True, False),
        ("full_in256_H128x2_add_B2", 0, 2, 249, 256, 256, 0, 128, True, True),
        ("full_in16_H128x2_B1", 0, 1, 249, 256, 16, 0, 128, True, False),
        ("narrow_in256+16_H128x2_B1", 1, 1, 249, 256, 256, 16, 128, True, False),
        ("narrow_in256+16_H256x1_B1", 1, 1, 249, 256, 256, 16, 256, False, False),
        ("full_in256_H128x2_add_B64", 0, 64, 249, 256, 256, 0, 128, True, True),
        ("full_in16_H128x2_b256", 0, 256, 249, 256, 16, 0, 128, True, False),
        ("narrow_in256+16_H128x2_add_b256", 1, 256, 249, 256, 256, 16, 128, True, True),
        ("full_in256_H128x2_add_b256", 0, 256, 249, 256, 256, 0, 128, True, True),
        ("narrow_in256_H128x2_add_b256", 1, 256, 249, 256, 256, 0, 128, True, True),
        ("narrow_in256+16_H256x1_add_b15", 1, 15, 249, 256, 256, 16, 256, False, True),
        ("narrow_in256+16_H256x1_add_b60", 1, 60, 249, 256, 256, 16, 256, False, True),
        ("narrow_in256_H256x1_add_b15", 1, 15, 249, 256, 256, 0, 256, False, True),
        ("narrow_in256_H256x1_add_b60", 1, 60, 249, 256, 256, 0, 256, False, True)]


def main():
    filt = sys.argv[1:]
    tag = os.path.basename(os.environ.get("FNSSL_B200_LIB", "default"))
    for name, axis, nb, nt, nf, c0, c1, H, bidir, add in CFGS:
        if filt and not any(f in name for f in filt):
            continue
        torch.manual_seed(0)
        p = LSTMParams(c0 + c1, H, bidirectional=bidir).cuda()
        g0 = torch.randn(nb, nt, nf, c0, device="cuda").half()
        g1 = torch.randn(nb, nt, nf, c1, device="cuda").half() if c1 else None
        oc = H * (2 if bidir else 1)
        ga0 = torch.randn(nb, nt, nf, oc, device="cuda").half() if add else None
        flop = 2.0 * nb * nt * nf * (2 if bidir else 1) * 4 * H * (c0 + c1 + H)
        ts = []
        for i in range(9):
            ga = ga0.clone() if add else None            # the residual sum is accumulated in place
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h, hs = run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga, inplace_addend=add)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        if os.environ.get("LSTM_TIME_BURST"):      # K launches back to back between one pair of events: host launch latency amortised
            K = 10
            gas = [ga0.clone() for _ in range(K)] if add else [None] * K
            torch.cuda.synchronize()
            import time
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = time.perf_counter()
            e0.record()
            for k in range(K):
                run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=gas[k], inplace_addend=add)
            e1.record()
            h1 = time.perf_counter()
            torch.cuda.synchronize()
            print(f"    burst of {K}: {e0.elapsed_time(e1) / K:.3f} ms per launch on the device, host issue {1e3 * (h1 - h0) / K:.3f} ms per launch", flush=True)
        ck = int(h.view(torch.int16).to(torch.int64).sum()) ^ (int(hs.view(torch.int16).to(torch.int64).sum()) if add else 0)
        print(f"[{tag}] {name}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s  checksum {ck & 0xffffffff:08x}  finite={bool(torch.isfinite(h).all())}", flush=True)


if __name__ == "__main__":
    main()
