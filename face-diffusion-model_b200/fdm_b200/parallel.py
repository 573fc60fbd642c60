"""Multi-GPU plumbing: clips are independent, so a batch shards across ranks with no intra-step communication.

One process per GPU (torchrun), weights replicated, clip b of the global batch lives on rank b // B_local; the
in-kernel sampler noise is keyed by the GLOBAL clip index, so results do not depend on the world size. The only
collective is one all-gather of the decoded vertex sequences; the reference has no distributed
code at all."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, equal-count split (the all-gather needs equal counts): returns (first clip, clip count)."""
    if n_clips % world != 0:
        raise ValueError(f"{n_clips} clips do not split evenly over {world} ranks (pad the batch)")
    per = n_clips // world
    return rank * per, per


def gather_clips(local: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather a per-rank (B_local, ...) tensor into (world * B_local, ...) in rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.view(-1), group=group)
    return out


def sample_sharded(run_local: Callable[[int, int], torch.Tensor], n_clips: int, group=None) -> torch.Tensor:
    """run_local(first_clip, count) -> (count, T, V) vertices of this rank's shard; returns all clips on every rank."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    first, count = shard_range(n_clips, rank, world)
    return gather_clips(run_local(first, count), group)
