"""Training-side forward pieces: DP-IPD targets and losses.

CPU: the oracle restatement against the golden vectors produced by the reference's own DPIPD / RemoveChFromBatch classes
(tests/golden/make_golden.py train).  GPU: the CUDA kernels (through the C ABI) against the same goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import training_oracle as tro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAGS = [("2mic", "MM"), ("3mic", "MM"), ("3micM", "M")]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "train_golden.npz"))


def _maxerr(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).float()
    b = torch.as_tensor(np.asarray(b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max())


# ---- oracle vs reference goldens (CPU) --------------------------------------------------------------------------------

@pytest.mark.parametrize("tag,mode", TAGS)
def test_oracle_targets_and_loss_match_reference(g, tag, mode):
    t = tro.fnssl_targets(g[f"{tag}_doa"], g[f"{tag}_vad"], g[f"{tag}_mic"], mode)
    assert _maxerr(t, g[f"{tag}_ipd_gt"]) <= 1e-6                       # unit-modulus targets: absolute tolerance
    loss = tro.fnssl_loss(torch.from_numpy(g[f"{tag}_pred"]), torch.from_numpy(g[f"{tag}_ipd_gt"]))
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-6 * float(g[f"{tag}_loss"])


def test_oracle_ipdnet_targets_and_pit(g):
    t = tro.ipdnet_targets(g["ipdnet_doa"], g["ipdnet_vad"], g["ipdnet_mic"], g["ipdnet_non_source"])
    assert _maxerr(t, g["ipdnet_ipd_gt"]) <= 1e-6
    loss, perm = tro.ipdnet_pit_loss(torch.from_numpy(g["ipdnet_pred"]), torch.from_numpy(g["ipdnet_ipd_gt"]))
    assert abs(float(loss) - float(g["ipdnet_pit_loss"])) <= 1e-6 and np.array_equal(perm.numpy(), g["ipdnet_pit_perm"])
    # the permutation really is the per-frame minimiser: no other assignment gives a smaller loss
    p = torch.from_numpy(g["ipdnet_pred"]).reshape(8, -1, 2)
    t2 = torch.from_numpy(g["ipdnet_ipd_gt"]).reshape(8, -1, 2)
    for r in range(8):
        keep = ((p[r] - t2[r]) ** 2).sum()
        swap = ((p[r][:, [1, 0]] - t2[r]) ** 2).sum()
        assert (perm[r].tolist() == [0, 1]) == bool(keep <= swap)


def test_training_abi_symbols_exported():
    from fn_ssl_b200 import _lib
    lib = _lib.load()
    for name in ("fnssl_dpipd_targets", "fnssl_ipd_mse_loss", "fnssl_ipd_pit_mse_loss"):
        assert hasattr(lib, name)


# ---- CUDA kernels (GPU) -----------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("tag,mode", TAGS)
def test_gpu_fnssl_targets_and_loss(g, tag, mode):
    from fn_ssl_b200 import training as T
    doa = torch.from_numpy(g[f"{tag}_doa"]).cuda()
    vad = torch.from_numpy(g[f"{tag}_vad"]).cuda()
    tgt = T.dpipd_targets(doa, g[f"{tag}_mic"], vad=vad, ch_mode=mode, speed=340.0)
    assert _maxerr(tgt, g[f"{tag}_ipd_gt"]) <= 2e-6                     # float32 cos/sin of a float64 phase
    per = T.dpipd_targets(doa, g[f"{tag}_mic"], vad=None, ch_mode=mode, speed=340.0, per_source=True)
    assert _maxerr(per, g[f"{tag}_ipd_per_source"]) <= 2e-6
    pred = torch.from_numpy(g[f"{tag}_pred"]).cuda().requires_grad_(True)
    loss = T.ipd_mse_loss(pred, torch.from_numpy(g[f"{tag}_ipd_gt"]).cuda())
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 2e-6 * float(g[f"{tag}_loss"])
    loss.backward()                                                     # analytic gradient == autograd of the reference formula
    pr = torch.from_numpy(g[f"{tag}_pred"]).requires_grad_(True)
    tro.fnssl_loss(pr, torch.from_numpy(g[f"{tag}_ipd_gt"])).backward()
    assert _maxerr(pred.grad, pr.grad.numpy()) <= 1e-8


@pytest.mark.gpu
def test_gpu_ipdnet_targets_and_pit_loss(g):
    from fn_ssl_b200 import training as T
    ns_t = torch.from_numpy(g["ipdnet_non_source"]).float().cuda()
    tgt = T.dpipd_targets(torch.from_numpy(g["ipdnet_doa"]).cuda(), g["ipdnet_mic"], vad=torch.from_numpy(g["ipdnet_vad"]).cuda(),
                          ch_mode="M", speed=340.0, vad_threshold=0.001, per_source=True, non_source=ns_t)
    assert _maxerr(tgt, g["ipdnet_ipd_gt"]) <= 2e-6
    np.testing.assert_allclose(T.non_source_target(g["ipdnet_mic"]), g["ipdnet_non_source"], rtol=0, atol=1e-12)
    loss, perm = T.ipd_pit_mse_loss(torch.from_numpy(g["ipdnet_pred"]).cuda(), torch.from_numpy(g["ipdnet_ipd_gt"]).cuda())
    assert abs(float(loss) - float(g["ipdnet_pit_loss"])) <= 2e-6 * float(g["ipdnet_pit_loss"])
    assert np.array_equal(perm.cpu().numpy(), g["ipdnet_pit_perm"])


@pytest.mark.gpu
def test_gpu_pit_three_sources_and_sizes():
    """3 sources (6 permutations) at a training-sized batch: loss equals the oracle's exhaustive search; deterministic."""
    from fn_ssl_b200 import training as T
    gen = torch.Generator().manual_seed(5)
    gt = torch.randn(4, 20, 512, 3, 3, generator=gen)
    order = torch.stack([torch.randperm(3, generator=gen) for _ in range(80)])            # a different shuffle per frame
    pred = torch.stack([gt.reshape(80, -1, 3)[r][:, order[r]] for r in range(80)]).reshape(gt.shape) + 0.05 * torch.randn(gt.shape, generator=gen)
    loss, perm = T.ipd_pit_mse_loss(pred.cuda(), gt.cuda())
    ref_loss, ref_perm = tro.ipdnet_pit_loss(pred, gt)
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * float(ref_loss)
    assert np.array_equal(perm.cpu().numpy(), ref_perm.numpy())
    loss2, _ = T.ipd_pit_mse_loss(pred.cuda(), gt.cuda())
    assert float(loss2) == float(loss)
