"""Small single-launch cases for compute-sanitizer (racecheck / synccheck / memcheck) of the tcgen05 kernels.

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py lstm128
Cases: lstm128 (H=128, cluster of 4, 128-row sub-tiles, residual reduce-add), lstm256 (H=256, cluster of 8, 64-row sub-tiles),
lstm256n (H=256 with the narrow second-source ring), pair128 / pair128f / pair256 (the CTA-pair kernels), conv (CausCnnBlock tcgen05 implicit GEMM).  Each prints its error against
the CPU oracle so a sanitizer-clean run is also a correct run.  Pipeline waits are unbounded here (the tool slows kernels ~100x).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.pop("FNSSL_TC_WAIT_TIMEOUT", None)
import torch  # noqa: E402
from fn_ssl_b200 import ops  # noqa: E402
from fn_ssl_b200.packing import LSTMParams, run_lstm  # noqa: E402
from oracle import fnssl_oracle as orc  # noqa: E402

LSTM = {  # axis, nb, nt, nf, c0, c1, H, bidir, addend(in place)
    "lstm128": (1, 1, 5, 300, 256, 0, 128, True, True),
    "lstm128f": (0, 1, 140, 6, 64, 16, 128, True, False),
    "lstm256": (1, 1, 5, 128, 256, 0, 256, False, True),
    "lstm256n": (1, 1, 5, 128, 256, 16, 256, False, False),
    # the CTA-pair kernels (cta_group::2), forced on small layers: lstm_tc5.cu (H = 128) and lstm_tc6.cu (H = 256)
    "pair128": (1, 1, 5, 300, 256, 0, 128, True, True),
    "pair128f": (0, 1, 300, 6, 64, 16, 128, True, False),
    "pair256": (1, 1, 5, 300, 256, 0, 256, False, True),
    "pair256n": (1, 1, 5, 300, 256, 16, 256, False, True),
    "pair128n6": (1, 1, 5, 300, 256, 16, 128, True, True),      # lstm_tc6.cu's H = 128 instantiation (two-source layer)
}


def lstm_case(name):
    axis, nb, nt, nf, c0, c1, H, bidir, add = LSTM[name]
    dev = "cuda"
    os.environ["FNSSL_TC_PAIR"] = "1" if name.startswith("pair") else "0"
    os.environ["FNSSL_TC_PAIR_MIN"] = "1000000" if name.endswith("6") else "1"
    os.environ["FNSSL_TC_PAIR128_MIN"] = "1"
    os.environ["FNSSL_TC_PAIR256_MIN"] = "1" if name.startswith("pair") else "1000000"
    torch.manual_seed(1)
    p = LSTMParams(c0 + c1, H, bidirectional=bidir).to(dev)
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(nb, nt, nf, c0, generator=g)
    x1 = torch.randn(nb, nt, nf, c1, generator=g) if c1 else None
    oc = H * (2 if bidir else 1)
    addend = torch.randn(nb, nt, nf, oc, generator=g) if add else None
    g0 = ops.grid_copy(x0.to(dev), c0, torch.float16)
    g1 = ops.grid_copy(x1.to(dev), c1, torch.float16) if c1 else None
    ga = ops.grid_copy(addend.to(dev), oc, torch.float16) if add else None
    ga_ref = ga.float().cpu() if add else None
    h, hs = run_lstm(p, "tcgen05", axis, g0, c0, g1, c1, addend=ga, inplace_addend=add)
    torch.cuda.synchronize()
    x = torch.cat([t for t in (g0[..., :c0].float().cpu(), g1[..., :c1].float().cpu() if c1 else None) if t is not None], -1)
    sd = {"l." + k: v.detach().cpu() for k, v in p.state_dict().items()}
    if axis == 0:
        ref = orc.lstm(x.reshape(nb * nt, nf, -1), sd, "l.").reshape(nb, nt, nf, oc)
    else:
        ref = orc.lstm(x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1), sd, "l.").reshape(nb, nf, nt, oc).permute(0, 2, 1, 3)
    err = float((h.float().cpu() - ref).abs().max() / ref.abs().max())
    msg = f"[{name}] rel-to-max err vs oracle {err:.3e}"
    if add:
        msg += f", h+addend {float((hs.float().cpu() - (ref + ga_ref)).abs().max() / ref.abs().max()):.3e}"
    print(msg)
    assert err <= 1e-3


def conv_case():
    import fn_ssl_b200 as F
    dev = "cuda"
    torch.manual_seed(21)
    cin0, cin1 = 64, 8
    cnn = F.CausCnnBlock(inp_dim=cin0 + cin1, out_dim=6, cnn_hidden_dim=128).eval().to(dev)
    g = torch.Generator().manual_seed(24)
    x = torch.randn(1, cin0 + cin1, 150, 26, generator=g)
    g0 = ops.cfirst_to_grid(x[:, :cin0].to(dev), torch.float16)
    g1 = ops.cfirst_to_grid(x[:, cin0:].to(dev), torch.float16)
    xq = torch.cat([g0[..., :cin0].float().cpu(), g1[..., :cin1].float().cpu()], -1).permute(0, 3, 2, 1)
    ref = orc.causcnn(xq, {"conv." + k: v.detach().cpu() for k, v in cnn.state_dict().items()})
    y = cnn.forward_grid(g0, cin0, g1, cin1)
    torch.cuda.synchronize()
    err = float((y.cpu() - ref).abs().max() / ref.abs().max())
    print(f"[conv] rel-to-max err vs oracle {err:.3e}")
    assert err <= 1e-3


if __name__ == "__main__":
    for name in sys.argv[1:] or ["lstm128"]:
        conv_case() if name == "conv" else lstm_case(name)
