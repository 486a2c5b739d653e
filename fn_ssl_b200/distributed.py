"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).

The path shards over utterances with no data-path exchange (every utterance / mic pair is independent through
the whole forward); the only collectives are
    * one broadcast of the flat weight buffer at start-up (what Lightning's DDP wrapper does at wrap time,
      FN-SSL/Lightning/main.py:286-288), and
    * one all-gather of the per-utterance outputs per batch (<= 1.3 MB per rank).
Training (fn_ssl_b200.training) is data-parallel the same way; its one exchange step is the gradient all-reduce that
Lightning's DDP strategy performs after backward (main.py:286-288): `all_reduce_gradients` -- one flat fp32 bucket per call.
Works with the gloo backend on CPU tensors too (used by the world_size-2 tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

Tensor = torch.Tensor


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment; returns (rank, world, local_rank)."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def broadcast_weights(module: nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer of `module` from rank `src` as ONE flat buffer; returns bytes sent."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return 0
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
            off += n
    return flat.numel() * 4


@torch.no_grad()
def all_gather_outputs(local_out: Tensor, counts: List[int]) -> Tensor:
    """Gather per-rank output shards (counts[r] utterances each, identical trailing shape) in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_out
    world = dist.get_world_size()
    assert len(counts) == world and local_out.shape[0] == counts[dist.get_rank()]
    mx = max(counts)
    tail = tuple(local_out.shape[1:])
    out = local_out.new_empty((world * mx,) + tail)
    if all(c == mx for c in counts):                     # the usual case: equal shards, no staging copy at all
        dist.all_gather_into_tensor(out, local_out.contiguous())
        return out
    padded = local_out.new_empty((mx,) + tail)           # uneven shards: pad to the largest (the padding rows are dropped below)
    padded[: local_out.shape[0]] = local_out
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


@torch.no_grad()
def all_reduce_gradients(module: nn.Module, average: bool = True) -> int:
    """Data-parallel training step, the exchange DDP does (FN-SSL/Lightning/main.py:286-288): sum (average) the gradients of
    every parameter over the ranks as ONE flat fp32 bucket (2.5 M parameters = 10 MB: a single NVSwitch all-reduce, sized for
    launch latency rather than overlap -- the backward pass of a step is hundreds of milliseconds).  Parameters without a
    gradient on this rank contribute zeros and receive the reduced value, as under DDP.  Returns the bucket's bytes."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * 4
