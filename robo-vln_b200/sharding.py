"""Data-parallel sharding of the policy step over environments.

Environments are independent (each has its own LSTM state), so rank r owns the contiguous
slice [r*B/G, (r+1)*B/G) of the global batch, weights are replicated, the hidden state stays
rank-local, and the ONLY collective of the path is one all-gather of the packed per-env
outputs [logits(4) | action(2) | stop(1)] -- 28 bytes per environment.  The reference has no
counterpart (it places hi on cuda:0 and lo on cuda:1, hierarchical_trainer.py:292-296,517-521).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

PACK = 7  # 4 sub-goal logits + (v, omega) + stop logit


def shard_range(global_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition (first `rem` ranks get one extra row)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(global_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_outputs(logits: torch.Tensor, actions: torch.Tensor, stop: torch.Tensor) -> torch.Tensor:
    return torch.cat([logits, actions, stop], dim=1).contiguous()


def unpack_outputs(packed: torch.Tensor):
    return packed[:, :4], packed[:, 4:6], packed[:, 6:7]


def all_gather_outputs(local: torch.Tensor, global_rows: int, group=None) -> torch.Tensor:
    """local [b_r, 7] on every rank -> [global_rows, 7] on every rank (rank order = row order)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(global_rows, r, world) for r in range(world)]
    maxb = max(hi - lo for lo, hi in sizes)
    if all((hi - lo) == maxb for lo, hi in sizes):
        out = torch.empty((global_rows, PACK), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged tail: pad to the largest shard, gather, then drop the padding
    pad = torch.zeros((maxb, PACK), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * maxb, PACK), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = [buf[r * maxb: r * maxb + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    assert parts[rank].shape[0] == local.shape[0]
    return torch.cat(parts, dim=0)
