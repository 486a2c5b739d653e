#!/usr/bin/env python
"""Benchmark of the FN-SSL forward hot path (BASELINE.json: TF-frames/s, 2ch, 4 s @ 16 kHz, 512/256 STFT).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

Headline workload (default, `--workload cfg4`): BASELINE.json configs[3] / north_star's scaling target -- FN-SSL with the
DOA-classification head (`FN_SSL(is_doa=True)`, FN-SSL/Lightning/Model.py:71,88-89), 2 mics, a GLOBAL batch of 256 x 4 s
split over the ranks by `distributed.shard_range` => STRONG scaling (total work fixed as N grows; reference: Lightning DDP,
FN-SSL/Lightning/main.py:286-288).  The same global-256 workload runs at --gpus 1, so the 1-GPU line and the N = 1 point of
the scaling run are the same measurement.  At N = 1 the line also carries `extra`: configs[1] (cfg2, 16 x 4 s, DP-IPD head,
both narrow-band variants), configs[0] (cfg1, one FN block, one utterance: latency) and `gpu_torch_baseline` -- the
reference's own PyTorch path (`torch.stft` + cuDNN `nn.LSTM`) timed on the same GPU (SURVEY.md section 0's stated bar).

A "step" = one pass of the whole path (STFT -> features -> 3 FN blocks -> DP-IPD head [-> DOA head]) over one batch of
synthetic white-noise utterances.  1 TF-frame = one STFT frame of one network batch row through the whole path.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSAMPLE, NCH, NT = 64000, 2, 249          # 4 s @ 16 kHz, 2 mics -> 249 frames (512/256, center=False)

WORKLOADS = {
    # name: (global batch [strong] or per-GPU batch [weak], scaling, is_doa, FN blocks)
    "cfg4": dict(batch=256, scaling="strong", is_doa=True, blocks=3,
                 text="FN-SSL 2-mic, 3 FN blocks, DOA-classification head (is_doa), GLOBAL batch {gb}x4s@16kHz sharded over {w} GPU(s) "
                      "(BASELINE configs[3]), 512/256 STFT"),
    "cfg2": dict(batch=16, scaling="weak", is_doa=False, blocks=3,
                 text="FN-SSL 2-mic, 3 FN blocks, DP-IPD head, batch {pb}x4s@16kHz per GPU (BASELINE configs[1]), 512/256 STFT"),
    "cfg1": dict(batch=1, scaling="weak", is_doa=False, blocks=1,
                 text="FN-SSL 2-mic, 1 FN block, DP-IPD head, {pb} utterance(s) of 4s@16kHz per GPU (BASELINE configs[0]), 512/256 STFT"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="offline", choices=["offline", "online"],
                    help="narrow-band layer: offline = BLSTM (north_star wording), online = uni-LSTM 256 (the code default)")
    ap.add_argument("--engine", default=None, help="auto / tcgen05 / simt")
    ap.add_argument("--batch", type=int, default=None, help="override: global batch (strong workloads) / per-GPU batch (weak)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads / GPU-PyTorch comparator at N = 1")
    return ap.parse_args()


def mac_per_bin(online: bool, blocks: int) -> int:
    """MAC per TF-bin of the LSTM stack."""
    first = 135168 + (528384 if online else 397312)
    later = 393216 + (524288 if online else 393216)
    return first + (blocks - 1) * later


def flops_per_frame(online: bool, blocks: int = 3) -> float:
    return 2.0 * mac_per_bin(online, blocks) * 256


def workload_config(args, world: int):
    """The `config` object of the JSON line -- identical in both arms (ours / reference)."""
    w = WORKLOADS[args.workload]
    batch = args.batch or w["batch"]
    gb = batch if w["scaling"] == "strong" else batch * world
    online = args.variant == "online"
    return {
        "workload": w["text"].format(gb=gb, w=world, pb=batch) + ", narrow-band=" + ("uni-LSTM(256)" if online else "BLSTM(2x128)"),
        "name": args.workload, "variant": args.variant, "global_batch": gb, "scaling": w["scaling"],
        "head": "doa180" if w["is_doa"] else "dp-ipd", "fn_blocks": w["blocks"],
        "flop_per_frame": flops_per_frame(online, w["blocks"]),
        "l2": "per-step activations (0.13 GB per utterance) exceed the 126 MB L2; no explicit flush",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) -- the ONLY place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------------------------

def cpu_reference_run(online: bool, is_doa: bool, blocks: int, steps: int, warmup: int, nb: int = 1):
    """Reference algorithm on the host cores: oracle port (torch CPU kernels = what the reference's nn.LSTM /
    torch.stft dispatch to), `nb` 4-s utterance(s) per step (a bounded sample of the workload).  The intra-op thread
    count is auto-tuned over {8 (the reference's own OMP_NUM_THREADS, main.py:25), 16, 32}: more threads than the
    small per-step GEMMs can use make oneDNN's LSTM slower (measured on the B200 host: 8: 1.21 s, 16: 1.02 s,
    32: 1.20 s, 64: 2.48 s, 128: 43 s per utterance), so larger counts are not probed."""
    import torch
    from oracle import fnssl_oracle as orc
    ncpu = os.cpu_count() or 1
    sd = orc.seeded_fnssl_state_dict(0, is_online=online, is_doa=is_doa)
    sig = orc.white_noise(nb, NSAMPLE, NCH)

    def one():
        t0 = time.perf_counter()
        feat = orc.preprocess_fnssl(sig)
        if blocks == 3:
            out = orc.fnssl_forward(feat, sd, fast=True)
            assert out.shape == (nb, NT // 12, 180 if is_doa else 512)
        else:   # cfg1: one FN block + DP-IPD head (FNblock.forward, FN-SSL/Lightning/Model.py:31-50)
            x, _, _ = orc.fnssl_block(feat.permute(0, 3, 2, 1), sd, "block_1.", True, fast=True)
            assert x.shape[-1] == 256
        return time.perf_counter() - t0

    cands = sorted({c for c in (8, 16, 32) if c <= ncpu}) or [ncpu]
    probe = {}
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            one()                                   # warm-up at this thread count
            probe[c] = one()
            if probe[c] > 1.15 * min(probe.values()):  # past the sweet spot
                break
        best = min(probe, key=probe.get)
        torch.set_num_threads(best)
        for _ in range(warmup):
            one()
        times = [one() for _ in range(steps)]
    total = sum(times)
    return {"value": nb * NT * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": best,
            "sample": f"{nb} utterance(s) of 4 s per step (a bounded sample of the workload's batch), {len(times)} timed steps after "
                      f"{warmup} warm-up; threads auto-tuned over {list(probe)} of {ncpu} host CPUs (s/step: "
                      + ", ".join(f"{k}:{v:.2f}" for k, v in probe.items()) + ")"}


# ------------------------------------------------------------------------------------------------------------------
# GPU-PyTorch comparator: the reference's own library path (torch.stft -> cuFFT, nn.LSTM -> cuDNN) on the same GPU.
# A restatement of FN-SSL/Lightning/Model.py:31-50,72-90 with stock torch modules; none of this repo's kernels.
# ------------------------------------------------------------------------------------------------------------------

def gpu_torch_baseline(dev, B: int, online: bool, is_doa: bool, steps: int = 3, warmup: int = 2):
    import torch
    import torch.nn as nn

    class Block(nn.Module):
        def __init__(self, inp, first):
            super().__init__()
            self.first = first
            self.full = nn.LSTM(inp, 128, batch_first=True, bidirectional=True)
            nin = 256 + (inp if first else 0)
            self.narr = nn.LSTM(nin, 256 if online else 128, batch_first=True, bidirectional=not online)

        def forward(self, x, fb_skip=None):
            nb, nt, nf, _ = x.shape
            nb_skip = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
            x = x.reshape(nb * nt, nf, -1)
            if not self.first:
                x = x + fb_skip
            x, _ = self.full(x)
            fb = x
            x = x.view(nb, nt, nf, -1).permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
            x = torch.cat((x, nb_skip), dim=-1) if self.first else x + nb_skip
            x, _ = self.narr(x)
            return x.view(nb, nf, nt, -1).permute(0, 2, 1, 3), fb

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.b1, self.b2, self.b3 = Block(4, True), Block(256, False), Block(256, False)
            self.emb = nn.Linear(256, 2)
            self.doa = nn.Linear(512, 180) if is_doa else None

        def forward(self, x):
            x = x.permute(0, 3, 2, 1)
            nb, nt, nf, _ = x.shape
            x, fb = self.b1(x)
            x, fb = self.b2(x, fb)
            x, fb = self.b3(x, fb)
            x = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
            ipd = torch.tanh(self.emb(nn.functional.avg_pool2d(x, kernel_size=(12, 1))))
            ipd = ipd.view(nb, nf, ipd.shape[1], -1).permute(0, 2, 1, 3)
            r = torch.cat((ipd[..., 0], ipd[..., 1]), dim=2)
            return self.doa(r) if self.doa is not None else r

    def preprocess(sig):   # FN-SSL/Lightning/main.py:206-225 + Module.py:48-68 + utils_.py:9-55 (2 mics: one pair)
        win = torch.hann_window(512, device=sig.device)
        spec = torch.stack([torch.stft(sig[:, :, c], n_fft=512, hop_length=256, win_length=512, window=win, center=False,
                                       normalized=False, return_complex=True) for c in range(sig.shape[2])], dim=1)
        mag = spec.abs().reshape(sig.shape[0], -1, spec.shape[-1])
        alpha, mu, mus = (298 - 1) / (298 + 1), 0, []
        for t in range(spec.shape[-1]):        # the reference's Python loop over frames
            a = min((t - 1) / (t + 1), alpha)
            mu = a * mu + (1 - a) * mag[:, :, t].mean(dim=1, keepdim=True)
            mus.append(mu)
        mu = torch.stack(mus, dim=-1).reshape(sig.shape[0], 1, 1, -1) + 1e-6
        return torch.cat((spec.real / mu, spec.imag / mu), dim=1)[:, :, 1:257, :].contiguous()

    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234)
    sig = torch.randn(B, NSAMPLE, NCH, generator=gen).to(dev)
    out = {"batch": B, "unit": "frames/s", "what": "torch.stft + nn.LSTM (cuDNN) restatement of the reference forward on this GPU; "
           "network-only = features resident, e2e = + reference-style preprocessing (per-frame Python loop of forgetting_norm)"}

    def timeit(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    with torch.no_grad():
        net = Net().to(dev).eval()
        feat = preprocess(sig)
        for tag, tf32, half in (("fp32", False, False), ("tf32", True, False), ("fp16", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                m, f = (net.half(), feat.half()) if half else (net.float(), feat.float())
                ms = timeit(lambda: m(f))
                out["network_" + tag] = {"ms_per_step": round(ms, 3), "value": round(B * NT / ms * 1e3, 1)}
            except Exception as exc:   # e.g. out of memory: report, do not fail the bench
                out["network_" + tag] = {"error": str(exc)[:200]}
        net.float()
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            ms = timeit(lambda: net(preprocess(sig)))
            out["e2e_tf32"] = {"ms_per_step": round(ms, 3), "value": round(B * NT / ms * 1e3, 1)}
        except Exception as exc:
            out["e2e_tf32"] = {"error": str(exc)[:200]}
    del net, feat, sig
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------

def main():
    args = parse()
    online = args.variant == "online"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(online, wl["is_doa"], wl["blocks"], max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": "TF-frames/sec", "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, max(1, args.gpus)),
                "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # rank 0 prints ONE JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION, printed to stdout at
    # communicator creation) out of it; an explicit INFO / TRACE setting is left alone (its output then goes to stderr)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]
    elif os.environ.get("NCCL_DEBUG"):
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import fn_ssl_b200 as F
    from fn_ssl_b200 import config, ops
    from fn_ssl_b200 import distributed as D
    # (oracle/ is imported only inside cpu_reference_run -- the cpu_baseline / reference-arm legs)

    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus and rank == 0 and world > 1:
        print(f"[bench] note: WORLD_SIZE={world} differs from --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class OneBlockNet(nn.Module):
        """cfg1: FNblock(input_size=4, is_first=True) + DP-IPD head (FN-SSL/Lightning/Model.py:6-50,79-87)."""

        def __init__(self, is_online):
            super().__init__()
            self.block_1 = F.FNblock(input_size=4, is_online=is_online, is_first=True)
            self.emb2ipd = nn.Linear(256, 2)
            self.engine = None

        def _engine(self):
            return config.resolve(self.engine, (self.block_1.full_hidden_size, self.block_1.narr_hidden_size))

        def forward_grid(self, g0, eng=None, states=None):
            eng = eng or self._engine()
            n1, _, _ = self.block_1._run(eng, g0, 4, g0, None, False, False)
            return ops.ipd_head(n1, n1.shape[-1], self.emb2ipd.weight, self.emb2ipd.bias)

    def run_workload(name: str, is_online: bool, batch: int, K: int, W: int, want_profile: bool):
        """Measure one workload on this process group: device-resident value, end-to-end value, per-kernel CUDA-event times."""
        w = WORKLOADS[name]
        # weights: PyTorch default init under torch.manual_seed(0) -- the module reproduces the reference's init stream, so
        # this equals `torch.manual_seed(0); FN_SSL(...)` of the reference; rank 0's copy is broadcast
        torch.manual_seed(0 if rank == 0 else 1000 + rank)
        net = (F.FN_SSL(is_online=is_online, is_doa=w["is_doa"]) if w["blocks"] == 3 else OneBlockNet(is_online)).eval()
        net.to(dev)
        wbytes = D.broadcast_weights(net, src=0)               # thin weight broadcast (NCCL)
        net.engine = args.engine
        eng = net._engine()
        pipe = F.FNSSLPipeline(net)

        if w["scaling"] == "strong":                           # the global batch is split over the ranks
            gb = batch
            spans = [D.shard_range(gb, r, world) for r in range(world)]
        else:
            gb = batch * world
            spans = [(r * batch, (r + 1) * batch) for r in range(world)]
        counts = [hi - lo for lo, hi in spans]
        lo, hi = spans[rank]
        nloc = hi - lo
        gen = torch.Generator().manual_seed(1234 + rank)        # white noise, sigma = 1
        sig_host = torch.randn(max(nloc, 1), NSAMPLE, NCH, generator=gen, dtype=torch.float32)[:nloc].pin_memory()
        sig_dev = sig_host.to(dev)
        tail = (NT // 12, 180 if w["is_doa"] else 512)
        out_host = torch.empty((gb,) + tail, dtype=torch.float32).pin_memory() if rank == 0 else None

        def step_resident():
            return D.all_gather_outputs(pipe(sig_dev), counts)      # per-utterance outputs gathered on every rank

        def step_e2e():
            # the public end-to-end call: pinned host input -> (H2D on a side stream, double-buffered, so step i+1's copy
            # overlaps step i's kernels) -> forward -> all-gather -> D2H of the gathered result on rank 0
            out = D.all_gather_outputs(pipe.run_host(sig_host), counts)
            if out_host is not None:
                out_host.copy_(out, non_blocking=True)
            return out

        def timed(fn, steps, warmup, profile=False):
            for _ in range(warmup):
                fn()
            barrier()
            l0 = ops.LAUNCHES
            if profile:
                ops.profile_start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            rec = ops.profile_stop() if profile else []
            ms = e0.elapsed_time(e1)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), ops.LAUNCHES - l0, rec

        ms_total, launches, rec = timed(step_resident, K, W, profile=want_profile)
        ms_e2e, _, _ = timed(step_e2e, K, max(1, min(W, 2)))
        frames = gb * NT
        by = {}
        for label, flops, nbytes, a, b in rec:
            d = by.setdefault(label, {"ms": 0.0, "n": 0, "flops": flops, "bytes": nbytes})
            d["ms"] += a.elapsed_time(b); d["n"] += 1
        kernels = []
        for label, d in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
            avg = d["ms"] / d["n"]
            kernels.append({"kernel": label, "launches": d["n"], "avg_ms": round(avg, 4), "share_of_step": round(d["ms"] / ms_total, 4),
                            "tflops": round(d["flops"] / (avg * 1e-3) / 1e12, 2), "hbm_gbs": round(d["bytes"] / (avg * 1e-3) / 1e9, 1),
                            "flop_per_launch": d["flops"], "bytes_per_launch": d["bytes"]})
        res = {"value": frames * K / (ms_total * 1e-3), "ms_per_step": ms_total / K, "ms_total": ms_total,
               "e2e_value": frames * K / (ms_e2e * 1e-3), "e2e_ms_per_step": ms_e2e / K,
               "h2d": gb * NSAMPLE * NCH * 4, "d2h": gb * tail[0] * tail[1] * 4, "launches": launches, "kernels": kernels,
               "engine": eng, "global_batch": gb, "per_gpu_batch": max(counts), "weights_broadcast_bytes": wbytes}
        del pipe, net, sig_dev
        torch.cuda.empty_cache()
        return res

    W, K = max(3, args.warmup), max(1, args.steps)
    batch = args.batch or wl["batch"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main_res = run_workload(args.workload, online, batch, K, W, True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary workloads (N = 1 only; every rank would have to take part in the collectives otherwise)
    extra = None
    if world == 1 and not args.no_extra and args.workload == "cfg4":
        extra = {}
        for tag, name, onl, b, k in (("cfg2_offline", "cfg2", False, 16, 20), ("cfg2_online", "cfg2", True, 16, 10),
                                     ("cfg1_offline", "cfg1", False, 1, 20), ("cfg1_online", "cfg1", True, 1, 20)):
            r = run_workload(name, onl, b, k, 3, True)
            e = {"workload": WORKLOADS[name]["text"].format(gb=b, w=1, pb=b) + (", online" if onl else ", offline"),
                 "value": round(r["value"], 1), "unit": "frames/s", "ms_per_step": round(r["ms_per_step"], 4),
                 "e2e": {"value": round(r["e2e_value"], 1), "ms_per_step": round(r["e2e_ms_per_step"], 4)}, "steps": k,
                 "kernels": [{kk: v for kk, v in kr.items() if kk in ("kernel", "avg_ms", "tflops")} for kr in r["kernels"][:6]]}
            if name == "cfg1":
                e["latency_ms"] = round(r["ms_per_step"], 4)
            extra[tag] = e

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA-event durations recorded inside the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_burst = float(peaks.get("bf16_tflops", 1650.0))
    tf_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    kernels, eng = main_res["kernels"], main_res["engine"]
    roof = None
    if kernels:
        top = kernels[0]
        tensor_bound = eng == "tcgen05"
        # SIMT engine: fp32 CUDA cores; its roof is the FP32 FMA pipe (148 SMs x 128 lanes x 2 x clock), not the tensor pipe
        fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        # burst peak: the per-launch events time each kernel on its own (0.9 - 15 ms); the sustained figure (a 4-s back-to-back
        # matmul at power-capped clocks) is printed beside it
        peak = tf_burst if tensor_bound else fp32_peak
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this layer
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
            ent = tr.get(f"{top['kernel']}@B{main_res['per_gpu_batch']}")
            if ent:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": top["tflops"], "peak": round(peak, 1), "unit": "TFLOP/s",
                "frac": round(top["tflops"] / peak, 4), "traffic": traffic,
                "peak_source": (peak_src + ", bf16/fp16 dense BURST (kernel timed alone per launch)") if tensor_bound
                else "FP32 FMA pipe, 148 SMs x 128 x 2 x max clock",
                "frac_vs_sustained_peak": round(top["tflops"] / tf_sust, 4) if tensor_bound else None,
                "sustained_peak": tf_sust if tensor_bound else None,
                "algorithmic_flop_per_launch": top["flop_per_launch"], "algorithmic_bytes_per_launch": top["bytes_per_launch"],
                "avg_launch_ms": top["avg_ms"],
                "hbm_achieved_gbs": top["hbm_gbs"], "hbm_peak_gbs": hbm_peak, "hbm_frac": round(top["hbm_gbs"] / hbm_peak, 4),
                "note": "LSTM layers are tensor-pipe bound (AI 260-512 FLOP/B); hbm_frac is the figure BASELINE.json names"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(online, wl["is_doa"], wl["blocks"], 3, 1)
        cpu = {"value": round(r["value"], 1), "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    torch_base = None
    if world == 1 and not args.no_extra:
        try:
            torch_base = gpu_torch_baseline(dev, 32 if args.workload == "cfg4" else 16, online, wl["is_doa"])
        except Exception as exc:
            torch_base = {"error": str(exc)[:300]}

    # training step at batch 8 x 4 s, forward + MSE loss + backward, next to cuDNN + autograd on the same GPU: run
    # in a child process (tools/bench_train.py) so that nothing it does can disturb this line; informational, never fatal.
    train_step = None
    if world == 1 and not args.no_extra and args.workload == "cfg4":
        try:
            torch.cuda.empty_cache()
            res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_train.py"), "8"], capture_output=True, text=True,
                                 timeout=90)
            last = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            train_step = json.loads(last[-1]) if last else {"error": (res.stderr or "no output")[-300:]}
        except Exception as exc:
            train_step = {"error": str(exc)[:300]}
        if extra is not None:
            extra["training_step_b8"] = train_step

    cfg = workload_config(args, world)
    run = {"engine": eng, "per_gpu_batch": main_res["per_gpu_batch"], "weights_broadcast_bytes": main_res["weights_broadcast_bytes"]}
    fpf = flops_per_frame(online, wl["blocks"])
    line = {
        "metric": "TF-frames/sec", "value": round(main_res["value"], 1), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(main_res["ms_per_step"], 4), "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate+state (tcgen05)" if eng == "tcgen05" else "f32",
        "data": "synthetic", "config": cfg, "run": run,    # `config` is key-for-key the reference arm's
        "e2e": {"value": round(main_res["e2e_value"], 1), "unit": "frames/s", "h2d_bytes_per_step": main_res["h2d"],
                "d2h_bytes_per_step": main_res["d2h"], "ms_per_step": round(main_res["e2e_ms_per_step"], 4),
                "note": "pinned host signal -> H2D (side stream, double-buffered) -> forward -> all-gather -> D2H of the gathered output on rank 0"},
        "gpu_launches": main_res["launches"], "clocks": clocks, "roofline": roof,
        "kernels": [{k: v for k, v in kr.items() if k not in ("flop_per_launch", "bytes_per_launch")} for kr in kernels[:8]],
        "cpu_baseline": cpu, "gpu_torch_baseline": torch_base, "extra": extra,
        "model_tflops": round(main_res["value"] * fpf / 1e12, 2),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
