"""world_size-2 gloo test of the multi-GPU host logic (weight broadcast, utterance sharding, output all-gather)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from fn_ssl_b200 import FN_SSL
    from fn_ssl_b200 import distributed as D
    D.init_from_env(backend="gloo")
    torch.manual_seed(100 + rank)                       # different init per rank ...
    net = FN_SSL()
    nbytes = D.broadcast_weights(net, src=0)            # ... equalised by the broadcast
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(ref) for _ in range(world)]
    dist.all_gather(gathered, ref)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    n_utt = 5                                           # uneven shard: 3 + 2
    lo, hi = D.shard_range(n_utt, rank, world)
    counts = [D.shard_range(n_utt, r, world)[1] - D.shard_range(n_utt, r, world)[0] for r in range(world)]
    full = torch.arange(n_utt * 20 * 4, dtype=torch.float32).reshape(n_utt, 20, 4)
    out = D.all_gather_outputs(full[lo:hi].clone(), counts)
    # IPDnet2: a second model family through the same helpers (326 tensors, 5-D per-utterance outputs)
    from fn_ssl_b200 import OnlineSpatialNet
    torch.manual_seed(200 + rank)
    net2 = OnlineSpatialNet(dim_input=10, dim_output=16, num_layers=8, dim_squeeze=8, num_freqs=256, dim_hidden=96,
                            attention='mamba(16,4)')
    D.broadcast_weights(net2, src=0)
    ref2 = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
    g2 = [torch.empty_like(ref2) for _ in range(world)]
    dist.all_gather(g2, ref2)
    full5 = torch.arange(n_utt * 3 * 8 * 4 * 2, dtype=torch.float32).reshape(n_utt, 3, 8, 4, 2)
    out5 = D.all_gather_outputs(full5[lo:hi].clone(), counts)
    same = same and all(torch.equal(g, g2[0]) for g in g2) and torch.equal(out5, full5)
    q.put((rank, same, nbytes, torch.equal(out, full), (lo, hi)))
    dist.destroy_process_group()


def test_broadcast_shard_allgather_gloo():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [r[4] for r in res] == [(0, 3), (3, 5)]
    for rank, same, nbytes, ok, _ in res:
        assert same and ok
        assert nbytes == 2511362 * 4                     # FN_SSL online parameter count


def test_shard_range_covers_everything():
    from fn_ssl_b200.distributed import shard_range
    for n in (1, 7, 16, 256):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _grad_worker(rank, world, port, q):
    """Data-parallel training step on two ranks == the same step on the concatenated batch (what DDP guarantees): per-rank
    gradients of a mean loss over the local shard, averaged by all_reduce_gradients, equal the full-batch gradient."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from fn_ssl_b200 import distributed as D
    from fn_ssl_b200.packing import LSTMParams
    D.init_from_env(backend="gloo")
    torch.manual_seed(7)                                 # same parameters on every rank (as after broadcast_weights)
    layer = LSTMParams(6, 32, bidirectional=True)        # the package's nn.LSTM-shaped parameter holder
    head = torch.nn.Linear(64, 3)
    frozen = torch.nn.Linear(3, 3)
    for p in frozen.parameters():
        p.requires_grad_(False)
    model = torch.nn.ModuleList([layer, head, frozen])

    def forward(x):                                      # host-side stand-in for the CUDA layer: same parameters through ATen
        flat = [t for d in layer.directions() for t in d]
        h = torch.zeros(2, x.shape[0], 32)
        y = torch._VF.lstm(x, (h, h), flat, True, 1, 0.0, False, True, True)[0]
        return head(y).pow(2).mean()

    full = torch.randn(4, 5, 6, generator=torch.Generator().manual_seed(3))
    lo, hi = D.shard_range(4, rank, world)
    forward(full[lo:hi]).backward()
    nbytes = D.all_reduce_gradients(model)
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
    model.zero_grad()
    forward(full).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
    ok = torch.allclose(got, want, rtol=1e-5, atol=1e-7) and all(p.grad is None for p in frozen.parameters())
    q.put((rank, ok, nbytes, int(want.numel())))
    dist.destroy_process_group()


def test_gradient_all_reduce_equals_full_batch_gloo():
    world, port = 2, 29733
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
    for rank, ok, nbytes, n in res:
        assert ok and nbytes == 4 * n
