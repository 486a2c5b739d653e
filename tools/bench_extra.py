"""Secondary measurements (not the contract bench): IPDnet cfg3 (4-mic, hidden 256, online, batch 32x4s) and FN-SSL at
other batch sizes, on one GPU.  Prints one JSON object per workload."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fn_ssl_b200 as F  # noqa: E402
from fn_ssl_b200 import ops  # noqa: E402


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ops.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    rec = ops.profile_stop()
    by = {}
    for label, flops, nbytes, a, b in rec:
        d = by.setdefault(label, [0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += 1
    return e0.elapsed_time(e1) / steps, {k: round(v[0] / v[1], 3) for k, v in by.items()}


def main():
    dev = "cuda"
    which = sys.argv[1:] or ["ipdnet", "fnssl_b64"]
    if "ipdnet" in which:
        kw = dict(input_size=8, hidden_size=256, max_track=2, is_online=True)
        torch.manual_seed(0)
        net = F.IPDnet(**kw).eval()
        pipe = F.IPDnetPipeline(net.to(dev))
        sig = torch.randn(32, 64000, 4, device=dev)
        ms, layers = timed(lambda: pipe(sig))
        print(json.dumps({"workload": "IPDnet 4-mic hidden 256 online, batch 32x4s (cfg3)", "ms_per_step": round(ms, 3),
                          "frames_per_s": round(32 * 249 / ms * 1e3, 1), "lstm_ms": layers}))
    if "decode" in which:
        import numpy as np
        gd = F.DPIPD(ndoa_candidate=[37, 73], mic_location=np.array(((-0.04, 0.0, 0.0), (0.04, 0.0, 0.0))), nf=257,
                     fre_max=8000, ch_mode="MM", speed=340)
        sdl = F.SourceDetectLocalize(max_num_sources=1, source_num_mode="kNum", meth_mode="IDL")
        netout = torch.randn(16, 20, 512, device=dev).tanh()
        ms, _ = timed(lambda: F.pred_ipd_to_doa(netout, gd, sdl, ch_mode="MM"))
        print(json.dumps({"workload": "IPD->DOA decode (IDL, 37 candidates), 16 utterances x 20 frames", "ms_per_step": round(ms, 4)}))
    for tag, B, online in (("fnssl_b64", 64, False), ("fnssl_b64_online", 64, True), ("fnssl_b4", 4, False)):
        if tag not in which:
            continue
        torch.manual_seed(0)
        net = F.FN_SSL(is_online=online).eval()
        pipe = F.FNSSLPipeline(net.to(dev))
        sig = torch.randn(B, 64000, 2, device=dev)
        ms, layers = timed(lambda: pipe(sig))
        print(json.dumps({"workload": f"FN-SSL {'online' if online else 'offline'} batch {B}x4s", "ms_per_step": round(ms, 3),
                          "frames_per_s": round(B * 249 / ms * 1e3, 1), "lstm_ms": layers}))


def stream_bench(which):
    """Streaming: per-chunk latency of FNSSLStream.push (12 new frames = 192 ms of audio per stream) vs number of streams."""
    dev = "cuda"
    torch.manual_seed(0)
    net = F.FN_SSL(is_online=True).eval().to(dev)
    for nb in (1, 16, 64):
        if f"stream_b{nb}" not in which and "stream" not in which:
            continue
        st = F.FNSSLStream(net, nb=nb, nch=2)
        first = torch.randn(nb, 512 + 256 * 11, 2, device=dev)
        chunk = torch.randn(nb, 256 * 12, 2, device=dev)
        st.push(first)
        ms, layers = timed(lambda: st.push(chunk), steps=20, warmup=3)
        print(json.dumps({"workload": f"FN-SSL online streaming, {nb} streams, 12-frame (192 ms) chunks", "ms_per_chunk": round(ms, 3),
                          "realtime_factor": round(192.0 / ms, 1), "frames_per_s": round(nb * 12 / ms * 1e3, 1), "lstm_ms": layers}))


if __name__ == "__main__":
    stream_bench(sys.argv[1:])
    main()
