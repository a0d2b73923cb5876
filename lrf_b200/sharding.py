"""Multi-GPU plumbing: images are independent (no cross-image term anywhere in
lrf/compression/qmf.py:116-292), so a batch is cut into contiguous per-rank ranges with no collective on
the data path; the only communication is one all_gather of the per-image (bpp, PSNR) pairs at the end."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Images [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(stats: torch.Tensor, total: int, rank: int, world: int) -> torch.Tensor:
    """all_gather of ragged per-rank (n_i, k) float tensors → (total, k) in image order on every rank."""
    if world == 1:
        return stats
    k = stats.shape[1]
    longest = (total + world - 1) // world
    padded = torch.zeros((longest, k), dtype=stats.dtype, device=stats.device)
    padded[: stats.shape[0]] = stats
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts)
