"""Scene-wise sharding of a batched query over the GPUs of one box (one process per GPU).

Scenes are independent: rank r owns a contiguous range of scenes, registers its own
copy of the mesh assets and runs the same kernels; there is no data-path collective. The only
exchange is the all-gather of the per-scene statistics record that the pack kernels write
([scenes, 6] int64 = {geometry pairs, BV-pair tests, element pairs, faces, surfaces, reserved}),
issued on the stream that produced it (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

STATS_COLUMNS = ("geometry_pairs", "bv_pair_tests", "element_pairs", "faces", "surfaces", "reserved")


def shard_bounds(total_scenes: int, world_size: int) -> np.ndarray:
    """[world_size + 1] boundaries of the contiguous, balanced scene ranges (strong split)."""
    if total_scenes < 0 or world_size < 1:
        raise ValueError("shard_bounds: bad arguments")
    base, rem = divmod(total_scenes, world_size)
    sizes = np.full(world_size, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def cost_balanced_bounds(scene_costs, world_size: int) -> np.ndarray:
    """[world_size + 1] boundaries of contiguous scene ranges whose largest total cost is minimal (SURVEY.md §8e:
    "balanced further by a precomputed cost" when scenes differ a lot, e.g. the clutter config). scene_costs: a
    non-negative cost per scene, typically the element_pairs column of a dry run's gathered statistics
    (gather_scene_stats). Exact: binary search on the bottleneck over the prefix sums, then a greedy fill that also
    leaves no rank without a scene while scenes remain."""
    c = np.asarray(scene_costs, dtype=np.float64)
    if c.ndim != 1 or world_size < 1 or np.any(c < 0) or not np.all(np.isfinite(c)):
        raise ValueError("cost_balanced_bounds: bad arguments")
    n = len(c)
    prefix = np.concatenate([[0.0], np.cumsum(c)])

    def parts_needed(limit: float) -> int:
        parts, start = 0, 0
        while start < n:
            # Furthest end with prefix[end] - prefix[start] <= limit (at least one scene per part).
            end = int(np.searchsorted(prefix, prefix[start] + limit, side="right")) - 1
            end = max(end, start + 1)
            parts += 1
            start = end
        return parts

    lo, hi = (float(c.max()) if n else 0.0), float(prefix[-1])
    for _ in range(100):
        if hi - lo <= 1e-12 * max(hi, 1.0):
            break
        mid = 0.5 * (lo + hi)
        if parts_needed(mid) <= world_size:
            hi = mid
        else:
            lo = mid
    bounds, start = [0], 0
    for r in range(world_size):
        remaining_ranks = world_size - r - 1
        if start >= n:
            bounds.append(n)
            continue
        end = int(np.searchsorted(prefix, prefix[start] + hi * (1 + 1e-12), side="right")) - 1
        end = max(end, start + 1)
        end = min(end, n - min(remaining_ranks, n - start - 1))  # keep a scene for each later rank while any remain
        if remaining_ranks == 0:
            end = n
        bounds.append(end)
        start = end
    return np.asarray(bounds, dtype=np.int64)


def shard_of(total_scenes: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) scene range of `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("shard_of: rank out of range")
    b = shard_bounds(total_scenes, world_size)
    return int(b[rank]), int(b[rank + 1])


def weak_shard(scenes_per_rank: int, rank: int) -> Tuple[int, int]:
    """Weak scaling: every rank owns `scenes_per_rank` scenes; rank r starts at r * scenes_per_rank."""
    return rank * scenes_per_rank, (rank + 1) * scenes_per_rank


def gather_scene_stats(stats_local, total_scenes: Optional[int] = None, group=None):
    """All-gathers the per-scene statistics of every rank into scene order.

    stats_local: torch int64 tensor [S_local, 6] on the device the process group works on. With
    equal shards this is one all_gather_into_tensor; ragged shards (strong split of a scene count that
    does not divide) are padded to the largest shard and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats_local.clone()
    world = dist.get_world_size(group)
    n_local = torch.tensor([stats_local.shape[0]], dtype=torch.int64, device=stats_local.device)
    counts = torch.zeros(world, dtype=torch.int64, device=stats_local.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.tolist()
    width = stats_local.shape[1]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0], width), dtype=stats_local.dtype, device=stats_local.device)
        dist.all_gather_into_tensor(out, stats_local.contiguous(), group=group)
    else:
        m = max(counts)
        padded = torch.zeros((m, width), dtype=stats_local.dtype, device=stats_local.device)
        padded[: stats_local.shape[0]] = stats_local
        buf = torch.empty((world * m, width), dtype=stats_local.dtype, device=stats_local.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        out = torch.cat([buf[r * m: r * m + counts[r]] for r in range(world)], dim=0)
    if total_scenes is not None and out.shape[0] != total_scenes:
        raise RuntimeError(f"gathered {out.shape[0]} scene records, expected {total_scenes}")
    return out
