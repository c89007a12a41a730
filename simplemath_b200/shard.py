"""Multi-GPU plumbing for the hot path.

The path shards with zero exchange: every output element depends on one element
of each operand (include/math/calculate.h:96 of the reference), so rank g of G
owns the flat output range shard_range(n, g, G) and runs the same kernels on it.
Broadcast (small) operands are replicated.  The only collective is an
all-gather used to assemble shards for VERIFICATION, outside any timed region
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import shard_range  # noqa: F401  (re-export)


def shard_sizes(n: int, world: int, align: int = 1) -> list[int]:
    out = []
    for r in range(world):
        b, e = shard_range(n, r, world, align)
        out.append(e - b)
    return out


def gather_shards(local: torch.Tensor, n: int, align: int = 1, group=None) -> torch.Tensor:
    """All-gather the per-rank shards of a flat array of n elements (shards as
    shard_range cuts them) and return the assembled array on every rank."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n, world, align)
    rank = dist.get_rank(group)
    assert local.numel() == sizes[rank], (local.numel(), sizes[rank])
    width = max(sizes)
    pad = torch.zeros(width, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local.reshape(-1)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a per-rank scalar (device-measured milliseconds) over all ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_scalar_sum(value, dtype=torch.float64, device=None, group=None):
    """Sum of per-rank partial reductions (the sharded dot product: each rank reduces its flat
    range with smb_dot, then ONE scalar all-reduce -- the first place a collective sits on a
    path).  int32 partials wrap exactly like the single-GPU result when summed as
    int64 and truncated by the caller."""
    t = torch.tensor([value], dtype=dtype, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.item()
