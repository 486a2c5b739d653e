"""Bring-up diagnostics for the CTA-pair LSTM kernel (lstm_tc5.cu): small layers in subprocesses (a trapped launch kills only
its own CUDA context), error against the CPU oracle, the timeout site code on failure."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, axis, nb, nt, nf, c0, c1, bidir, addend (in place)   [H = 128 unless the name starts with h256]
    ("h256_time_c64_uni_L1", 1, 2, 1, 256, 64, 0, False, False),
    ("h256_time_c64_uni_L3", 1, 2, 3, 256, 64, 0, False, False),
    ("h256_time_c256_uni_add", 1, 3, 7, 256, 256, 0, False, True),
    ("h256_freq_c256_uni_ragged", 0, 3, 100, 5, 256, 0, False, False),
    ("h256_time_c16_bi_nf300", 1, 2, 4, 300, 16, 0, True, False),
    ("h256_time_c64+16_uni_L1", 1, 2, 1, 256, 64, 16, False, False),
    ("h256_time_c256+16_uni_add", 1, 3, 7, 256, 256, 16, False, True),
    ("h256_freq_c256+16_uni_ragged", 0, 3, 100, 5, 256, 16, False, False),
    ("m128_time_c64_uni_L1", 1, 2, 1, 256, 64, 0, False, False),
    ("m128_time_c64_uni_L3", 1, 2, 3, 256, 64, 0, False, False),
    ("m128_time_c256+16_bi_add", 1, 3, 7, 256, 256, 16, True, True),
    ("m128_freq_c16_bi_ragged", 0, 3, 100, 5, 16, 0, True, False),
    ("time_c64_uni_L1", 1, 2, 1, 256, 64, 0, False, False),
    ("time_c64_uni_L2", 1, 2, 2, 256, 64, 0, False, False),
    ("time_c64_uni_L5", 1, 2, 5, 256, 64, 0, False, False),
    ("freq_c16_bi", 0, 1, 300, 6, 16, 0, True, False),
    ("time_c256_bi_add", 1, 3, 7, 256, 256, 0, True, True),
    ("time_c256+16_bi_add", 1, 2, 9, 256, 256, 16, True, True),
    ("freq_c256_bi_add_ragged", 0, 3, 211, 5, 256, 0, True, True),
    ("time_c64_bi_ragged_nf40", 1, 3, 6, 40, 64, 0, True, False),
]


def run_case(idx):
    import torch
    from fn_ssl_b200 import _lib, ops
    from fn_ssl_b200.packing import LSTMParams, run_lstm
    from oracle import fnssl_oracle as orc
    name, axis, nb, nt, nf, c0, c1, bidir, use_add = CASES[idx]
    H, dev = (256 if name.startswith("h256") else 128), "cuda"
    torch.manual_seed(idx)
    p = LSTMParams(c0 + c1, H, bidirectional=bidir).to(dev)
    g = torch.Generator().manual_seed(100 + idx)
    x0 = torch.randn(nb, nt, nf, c0, generator=g)
    x1 = torch.randn(nb, nt, nf, c1, generator=g) if c1 else None
    oc = H * (2 if bidir else 1)
    add = torch.randn(nb, nt, nf, oc, generator=g) if use_add else None
    g0 = ops.grid_copy(x0.to(dev), c0, torch.float16)
    g1 = ops.grid_copy(x1.to(dev), c1, torch.float16) if c1 else None
    ga = ops.grid_copy(add.to(dev), oc, torch.float16) if use_add else None
    ga_ref = ga.float().cpu() if use_add else None
    try:
        h, hs = run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga, inplace_addend=use_add)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] LAUNCH FAILED: {str(e)[:200]}; timeout site = {_lib.load().fnssl_lstm_tc_error_site()}")
        return
    x = torch.cat([t for t in (g0[..., :c0].float().cpu(), g1[..., :c1].float().cpu() if c1 else None) if t is not None], -1)
    sd = {"l." + k: v.detach().cpu() for k, v in p.state_dict().items()}
    if axis == 0:
        ref = orc.lstm(x.reshape(nb * nt, nf, -1), sd, "l.").reshape(nb, nt, nf, oc)
    else:
        ref = orc.lstm(x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1), sd, "l.").reshape(nb, nf, nt, oc).permute(0, 2, 1, 3)
    hc = h.float().cpu()
    err = (hc - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print(f"[{name}] rel-to-max err = {rel:.3e}  finite={bool(torch.isfinite(hc).all())}")
    if rel > 1e-3:
        step_axis = 2 if axis == 0 else 1
        per_step = err.amax(dim=[d for d in range(4) if d != step_axis])
        print("   max err per step :", [f"{v:.2e}" for v in per_step[:8].tolist()])
        per_unit = err.amax(dim=(0, 1, 2)).reshape(-1, 8).amax(1)
        print("   max err per 8 units:", [f"{v:.1e}" for v in per_unit.tolist()])
        rows = err.amax(dim=3)
        rows = rows.reshape(nb * nt, nf).amax(1) if axis == 0 else rows.permute(0, 2, 1).reshape(nb * nf, nt).amax(1)
        print("   max err per 32 rows:", [f"{v:.1e}" for v in rows[: (rows.numel() // 32) * 32].reshape(-1, 32).amax(1)[:24].tolist()])
        print("   got[0,0,0,:8] =", hc[0, 0, 0, :8].tolist())
        print("   ref[0,0,0,:8] =", ref[0, 0, 0, :8].tolist())
    if use_add:
        rel1 = float((hs.float().cpu() - (ref + ga_ref)).abs().max() / (ref + ga_ref).abs().max())
        print(f"[{name}] h+addend rel err = {rel1:.3e}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        only = os.environ.get("PAIR_DEBUG_ONLY", "")
        for i in range(len(CASES)):
            if only and only not in CASES[i][0]:
                continue
            env = dict(os.environ, FNSSL_TC_PAIR="1", FNSSL_TC_PAIR_MIN="1", FNSSL_TC_PAIR256_MIN="1", FNSSL_TC_WAIT_TIMEOUT="1")
            if CASES[i][0].startswith("m128"):      # lstm_tc6.cu's H = 128 instantiation instead of lstm_tc5.cu
                env.update(FNSSL_TC_PAIR_MIN="1000000", FNSSL_TC_PAIR128_MIN="1")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=300, env=env)
            print(r.stdout.strip())
            if r.returncode != 0:
                print(f"[{CASES[i][0]}] exit code {r.returncode}: {r.stderr.strip()[-500:]}")
