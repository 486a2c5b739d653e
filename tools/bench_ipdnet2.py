"""IPDnet2 (OnlineSpatialNet + Mamba) on one GPU: BASELINE.json configs[4] shape per GPU -- 8 mics, 6 s @ 16 kHz,
batch 64 (S = 2 sources: dim_output 28; see DESIGN.md for the 3-source note) -- and the reference's own 5-mic
configuration.  Prints one JSON object per workload with per-launch CUDA-event times; `--cpu` adds the oracle's time on
a bounded sample (1 utterance) on the host cores."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fn_ssl_b200 as F  # noqa: E402
from fn_ssl_b200 import ops  # noqa: E402


def timed(fn, steps=5, warmup=3):
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
        d = by.setdefault(label, [0.0, 0, flops, nbytes])
        d[0] += a.elapsed_time(b); d[1] += 1
    ker = {k: {"launches_per_step": v[1] // steps, "avg_ms": round(v[0] / v[1], 4), "tflops_fp32": round(v[2] / (v[0] / v[1]) / 1e9, 2),
               "hbm_gbs": round(v[3] / (v[0] / v[1]) / 1e6, 1)} for k, v in by.items()}
    return e0.elapsed_time(e1) / steps, ker


def main():
    dev = "cuda"
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg5", "default"]
    for tag, B, M, secs in (("cfg5", 64, 8, 6), ("default", 16, 5, 4), ("tiny", 2, 2, 1)):
        if tag not in which:
            continue
        torch.manual_seed(0)
        net = F.OnlineSpatialNet(dim_input=2 * M, dim_output=4 * (M - 1), num_layers=8, dim_hidden=96, num_heads=4,
                                 dim_squeeze=8, num_freqs=256, attention='mamba(16,4)').eval().to(dev)
        pipe = F.IPDnet2Pipeline(net)
        n = secs * 16000
        sig = torch.randn(B, n, M, device=dev)
        nt = n // 320 + 1
        ms, ker = timed(lambda: pipe(sig))
        rec = {"workload": f"IPDnet2 OnlineSpatialNet {M}-mic 2-source, 8 layers, batch {B}x{secs}s (hop 320, center=True)",
               "ms_per_step": round(ms, 3), "frames_per_s": round(B * nt / ms * 1e3, 1), "frames_per_step": B * nt,
               "dtype": "f32", "kernels": ker}
        if "--cpu" in sys.argv:
            from oracle import ipdnet2_oracle as orc2
            sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            s1 = sig[:1].cpu()
            torch.set_num_threads(min(16, os.cpu_count() or 1))
            with torch.no_grad():
                orc2.ipdnet2_forward(orc2.preprocess_ipdnet2(s1[:, :16000]), sd)
                t0 = time.perf_counter()
                orc2.ipdnet2_forward(orc2.preprocess_ipdnet2(s1), sd)
                dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": round(nt / dt, 1), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"1 utterance of {secs} s, 1 timed pass after warm-up (oracle; the Mamba scan is a Python loop over frames)"}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
