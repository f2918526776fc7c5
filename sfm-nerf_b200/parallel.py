"""Multi-GPU host logic (one process per GPU, torch.distributed): how rays shard and how gradients
combine.  Rays are independent through K1-K5, so
  * inference: rank r renders the contiguous ray-id slice ``shard_range(H*W, r, world)`` of the frame
    into its own slice of the output -- no collective on the data path;
  * training: every rank draws ``n_rays`` rays of the global batch (distinct counter ranges of the
    device RNG), back-propagates its LOCAL mean-MSE, and the flat fp32 gradient (661,636 floats,
    2.6 MB) is sum-allreduced over NCCL/NVLink; Adam applies it scaled by 1/world, which equals the
    gradient of the global-batch mean (train.py:460) because the shards are equal-sized.
Pure host functions: usable (and tested) with the gloo backend on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) slice of ``n_items`` ray ids owned by ``rank``; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def batch_counter(iteration: int, rank: int, world: int) -> int:
    """Index of the (iteration, rank) ray batch in the global counter-based RNG stream: batches of
    different ranks / iterations never overlap."""
    return int(iteration) * int(world) + int(rank)


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the flat gradient vector (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def mean_scale(world: int) -> float:
    """Factor applied to the summed local-mean gradients (inside the fused Adam kernel)."""
    return 1.0 / float(world)


def gather_frame(local_rgb: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Assemble the full frame from the per-rank slices (outside any timed region)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rgb
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_range(n_items, r, world) for r in range(world))]
    pad = torch.zeros((max(sizes),) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype, device=local_rgb.device)
    pad[: local_rgb.shape[0]] = local_rgb
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
