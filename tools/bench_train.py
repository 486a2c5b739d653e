"""Training-step timing (not the contract bench): FN_SSL (offline, 3 blocks) forward + MSE loss + backward on one GPU at a few
batch sizes -- the fp32 training path of fn_ssl_b200 (CUDA kernels with hand-written backward passes) next to the reference's own
library path restated with stock torch modules (nn.LSTM -> cuDNN, autograd), fp32 and TF32, same shapes, same GPU.
Prints one JSON object per batch size.      python tools/bench_train.py [B ...] [--once] [--ours-only] [--dw-compare] [--ipdnet]
(--once: a single step, for ncu; --ours-only: skip the cuDNN comparator; --dw-compare: time the step with both weight-gradient
kernels, FNSSL_TRAIN_DW=1 / 2; --ipdnet: also time IPDnet's training step, 4-mic hidden 256 online, batch 4)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import fn_ssl_b200 as F  # noqa: E402
from fn_ssl_b200 import training as T  # noqa: E402

NT, NF = 249, 256          # 4 s @ 16 kHz, 512/256 STFT


class TorchBlock(nn.Module):   # FN-SSL/Lightning/Model.py:6-50 with stock modules (dropout 0: deterministic comparison)
    def __init__(self, inp, first):
        super().__init__()
        self.first = first
        self.full = nn.LSTM(inp, 128, batch_first=True, bidirectional=True)
        self.narr = nn.LSTM(256 + (inp if first else 0), 128, batch_first=True, bidirectional=True)

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


class TorchNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.b1, self.b2, self.b3 = TorchBlock(4, True), TorchBlock(256, False), TorchBlock(256, False)
        self.emb = nn.Linear(256, 2)

    def forward(self, x):
        x = x.permute(0, 3, 2, 1)
        nb, nt, nf, _ = x.shape
        x, fb = self.b1(x)
        x, fb = self.b2(x, fb)
        x, fb = self.b3(x, fb)
        x = x.permute(0, 2, 1, 3).reshape(nb * nf, nt, -1)
        ipd = torch.tanh(self.emb(nn.functional.avg_pool2d(x, kernel_size=(12, 1))))
        ipd = ipd.view(nb, nf, ipd.shape[1], -1).permute(0, 2, 1, 3)
        return torch.cat((ipd[..., 0], ipd[..., 1]), dim=2)


def timed(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd = bwd = 0.0
    for _ in range(steps):
        ev[0].record()
        loss = step(fwd_only=True)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        fwd += ev[0].elapsed_time(ev[1])
        bwd += ev[1].elapsed_time(ev[2])
    return fwd / steps, bwd / steps


def main():
    once = "--once" in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 8]
    dev = "cuda"
    for B in sizes:
        torch.manual_seed(0)
        x = torch.randn(B, 4, NF, NT, device=dev)
        tgt = torch.randn(B, NT // 12, 2 * NF, device=dev).tanh()
        tgt4 = tgt.unsqueeze(-1).contiguous()                        # (nb, nt2, 2nf, P = 1): cal_loss layout
        net = F.FN_SSL(is_online=False).to(dev).train()
        for m in net.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0

        def ours(fwd_only=False):
            net.zero_grad(set_to_none=True)
            loss = T.ipd_mse_loss(net(x), tgt4)
            if not fwd_only:
                loss.backward()
            return loss

        if once:
            ours()
            torch.cuda.synchronize()
            print(json.dumps({"batch": B, "once": True}))
            continue
        res = {"workload": f"FN-SSL offline (3 blocks, BLSTM 2x128), training step, batch {B} x 4 s, fp32", "batch": B,
               "frames": B * NT}
        variants = [("fn_ssl_b200_fp32", None)]
        if "--dw-compare" in sys.argv:
            variants = [("fn_ssl_b200_fp32_dw1", "1"), ("fn_ssl_b200_fp32_dw2", "2")]
        for key, dw in variants:
            if dw is not None:
                os.environ["FNSSL_TRAIN_DW"] = dw          # read by fnssl_lstm_backward on every call
            f_ms, b_ms = timed(ours, steps=3, warmup=1)
            res[key] = {"forward_ms": round(f_ms, 2), "backward_ms": round(b_ms, 2), "step_ms": round(f_ms + b_ms, 2),
                        "frames_per_s": round(B * NT / (f_ms + b_ms) * 1e3, 1)}
        del net
        torch.cuda.empty_cache()
        if "--ours-only" in sys.argv:
            print(json.dumps(res), flush=True)
            continue
        ref = TorchNet().to(dev).train()
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def theirs(fwd_only=False):
                ref.zero_grad(set_to_none=True)
                loss = nn.functional.mse_loss(ref(x), tgt)
                if not fwd_only:
                    loss.backward()
                return loss
            try:
                f_ms, b_ms = timed(theirs, steps=3, warmup=1)
                res["torch_cudnn_" + tag] = {"forward_ms": round(f_ms, 2), "backward_ms": round(b_ms, 2), "step_ms": round(f_ms + b_ms, 2),
                                             "frames_per_s": round(B * NT / (f_ms + b_ms) * 1e3, 1)}
            except Exception as exc:
                res["torch_cudnn_" + tag] = {"error": str(exc)[:200]}
        del ref
        torch.cuda.empty_cache()
        print(json.dumps(res), flush=True)


def ipdnet_step():
    """IPDnet 4-mic, hidden 256, online (BASELINE configs[2]'s network) training step at batch 4 x 4 s: forward + frame-level PIT
    loss + backward, fp32 kernels."""
    dev = "cuda"
    B = 4
    torch.manual_seed(0)
    net = F.IPDnet(input_size=8, hidden_size=256, max_track=2, is_online=True).to(dev).train()
    x = torch.randn(B, 8, NF, NT, device=dev)
    gt = torch.randn(B, NT // 12, 2 * NF, 3, 2, device=dev).tanh()

    def step(fwd_only=False):
        net.zero_grad(set_to_none=True)
        loss, _ = T.ipd_pit_mse_loss(net(x), gt)
        if not fwd_only:
            loss.backward()
        return loss
    f_ms, b_ms = timed(step, steps=2, warmup=1)
    print(json.dumps({"workload": f"IPDnet 4-mic hidden 256 online, training step (PIT loss), batch {B} x 4 s, fp32", "batch": B,
                      "fn_ssl_b200_fp32": {"forward_ms": round(f_ms, 2), "backward_ms": round(b_ms, 2), "step_ms": round(f_ms + b_ms, 2),
                                           "frames_per_s": round(B * NT / (f_ms + b_ms) * 1e3, 1)}}), flush=True)


if __name__ == "__main__":
    main()
    if "--ipdnet" in sys.argv:
        ipdnet_step()
