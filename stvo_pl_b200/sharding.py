"""Host-side partitioning of independent frame pairs over ranks.

The path shards embarrassingly: every quantity of match + optimizePose is local to one (prev, curr) pair, so
B pairs are cut into contiguous blocks, one per rank (one process per GPU), with NO data-path collective.
torch.distributed is only plumbing: barrier, max-over-ranks of the device time, optional gather of results.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` pairs owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(results: np.ndarray, total: int, device=None) -> np.ndarray:
    """All ranks receive the PlPoseResult records of all pairs in pair order (host-side concatenation of the
    per-rank blocks; ~1 KB per pair, negligible next to the solve)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return results
    world = dist.get_world_size()
    item = results.dtype.itemsize
    counts = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
    cap = max(counts) * item
    buf = np.zeros(cap, np.uint8)
    raw = results.view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    mine = torch.from_numpy(buf).to(device) if device is not None else torch.from_numpy(buf)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    parts = [o.cpu().numpy()[:c * item].view(results.dtype) for o, c in zip(outs, counts)]
    return np.concatenate(parts)
