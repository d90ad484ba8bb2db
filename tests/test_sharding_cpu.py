"""Multi-GPU host logic on CPU (gloo, world_size 2): scene sharding, rank-independent scene
generation, and the all-gather of per-scene statistics — the only collective of the path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drake_b200 import sharding
from drake_b200 import workloads as W


def test_shard_bounds_cover_every_scene_once():
    for total, world in [(16384, 8), (10, 4), (3, 8), (0, 2), (1025, 2)]:
        b = sharding.shard_bounds(total, world)
        assert b[0] == 0 and b[-1] == total and np.all(np.diff(b) >= 0)
        assert np.diff(b).max() - np.diff(b).min() <= 1
        assert [sharding.shard_of(total, r, world) for r in range(world)] == [(int(b[r]), int(b[r + 1])) for r in
                                                                              range(world)]
    with pytest.raises(ValueError):
        sharding.shard_of(10, 4, 4)
    assert sharding.weak_shard(16384, 3) == (49152, 65536)


@pytest.mark.parametrize("name", ["c3", "c4", "c5", "c2s"])
def test_any_rank_generates_the_same_scenes(name):
    """Scene i's poses do not depend on which rank generates it or on the shard size."""
    kw = {"c5": dict(hint=0.5)}.get(name, {})   # a coarse sphere keeps the CPU suite fast
    whole = W.CONFIGS[name](12, 0, **kw).poses
    lo = W.CONFIGS[name](5, 0, **kw).poses
    hi = W.CONFIGS[name](7, 5, **kw).poses
    np.testing.assert_array_equal(np.concatenate([lo, hi]), whole)
    # Rotations are rotations.
    R = whole.reshape(-1, 12)[:, :9].reshape(-1, 3, 3)
    np.testing.assert_allclose(R @ np.transpose(R, (0, 2, 1)), np.tile(np.eye(3), (len(R), 1, 1)), atol=1e-14)


def _oracle_scene_stats(lo, hi):
    """Per-scene statistics records for C3 scenes [lo, hi) from the CPU oracle."""
    from oracle import ho
    from tests import parity
    wl = W.config3(hi - lo, lo)
    sc, order = parity.oracle_scene_for(wl)
    out = np.zeros((hi - lo, 6), np.int64)
    for s in range(hi - lo):
        surfs, st = sc.compute(wl.poses[s][order], ho.KPOLYGON)
        out[s] = [st["geometry_pairs"], 0, st["element_pairs"], st["faces"], len(surfs), 0]
    return out


def _worker(rank, world, port, total, ragged, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_of(total, rank, world) if ragged else sharding.weak_shard(total // world, rank)
        local = torch.from_numpy(_oracle_scene_stats(lo, hi))
        got = sharding.gather_scene_stats(local, total)
        if rank == 0:
            ret.put(got.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total,ragged", [(8, False), (7, True)])
def test_stats_all_gather_world_size_2(total, ragged):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, ragged, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = _oracle_scene_stats(0, total)
    np.testing.assert_array_equal(got, want)
    assert want[:, 4].sum() > 0   # the scenes do produce surfaces


def test_cost_balanced_bounds():
    """Contiguous cost-balanced scene ranges: cover every scene once, never worse than the equal-count split, optimal
    on small cases (checked against brute force), and equal to the equal-count split for uniform costs."""
    import itertools
    from drake_b200 import sharding
    rng = np.random.default_rng(0)

    def bottleneck(costs, b):
        return max(costs[b[r]:b[r + 1]].sum() for r in range(len(b) - 1))

    for n, w in [(1, 1), (1, 4), (7, 3), (12, 4), (10, 8), (64, 8)]:
        for _ in range(5):
            costs = rng.integers(0, 50, size=n).astype(float)
            b = sharding.cost_balanced_bounds(costs, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1 and np.all(np.diff(b) >= 0)
            assert np.count_nonzero(np.diff(b)) == min(n, w)     # no idle rank while scenes remain
            equal = sharding.shard_bounds(n, w)
            if n <= 12:   # brute force over all contiguous partitions into at most w non-empty parts
                best = min(bottleneck(costs, np.array((0,) + cut + (n,)))
                           for k in range(0, min(w, n)) for cut in itertools.combinations(range(1, n), k))
                best_full = min(bottleneck(costs, np.array((0,) + cut + (n,)))
                                for cut in itertools.combinations(range(1, n), min(w, n) - 1))
                assert best <= bottleneck(costs, b) <= best_full + 1e-9
            assert bottleneck(costs, b) <= bottleneck(costs, equal) + 1e-9 or n <= w
    np.testing.assert_array_equal(sharding.cost_balanced_bounds(np.ones(16), 4), sharding.shard_bounds(16, 4))
    with pytest.raises(ValueError):
        sharding.cost_balanced_bounds([1.0, -1.0], 2)
