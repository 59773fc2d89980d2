"""Batch-sharded data parallelism for the TT recurrent path (one process per GPU).

The path shards only along batch (sequences are independent; the few-KB TT cores are replicated):
inference needs no communication, training needs ONE collective per backward -- an all-reduce(sum) of
a single flat FP32 buffer holding every TT-core and bias gradient.  With NCCL over
NVLink/NVSwitch the message (21-306 KiB) is pure latency, so it is never split per parameter.
The reference has no distributed code at all; this module is what a maintainer adds around
`loss.backward()`.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) rows of a global batch owned by `rank`: contiguous blocks, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, world_size: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    """This rank's contiguous block of sequences of a batch-first tensor (a view, no copy)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def flatten_grads(params: Sequence[torch.Tensor]) -> torch.Tensor:
    """One contiguous FP32 buffer with every gradient (zeros where a parameter has none)."""
    parts = []
    for p in params:
        g = p.grad
        parts.append(torch.zeros(p.numel(), dtype=p.dtype, device=p.device) if g is None else g.reshape(-1))
    return torch.cat(parts) if parts else torch.zeros(0)


def unflatten_into_grads(flat: torch.Tensor, params: Sequence[torch.Tensor]) -> None:
    off = 0
    for p in params:
        n = p.numel()
        chunk = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = chunk.clone()
        else:
            p.grad.copy_(chunk)
        off += n


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> int:
    """Sum (or average) the gradients of `params` over all ranks with a single all-reduce.

    Returns the number of elements reduced.  No-op without an initialised process group."""
    plist: List[torch.Tensor] = [p for p in params if p.requires_grad]
    if not plist or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = flatten_grads(plist)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    unflatten_into_grads(flat, plist)
    return flat.numel()
