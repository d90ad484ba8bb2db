"""Parity at the BASELINE configs' own sizes (run on the B200 box with -m gpu): the 2 097 152-tetrahedron spheres of
config 5 and the 4096-scene clutter batch of config 4 against the oracle face for face, and the multi-chunk walk of a
large-mesh batch against its single-chunk run. (The smaller parity cases are in tests/test_gpu_parity.py.)"""
import numpy as np
import pytest

from oracle import ho
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import drake_b200 as db
    return db.Context(0)


def _digest(s):
    keys = np.sort(parity.face_keys(s.elem_M, s.elem_N))
    return (s.id_M, s.id_N, s.num_faces, len(s.vertices), keys.tobytes())


def test_one_full_size_config5_scene_against_the_oracle(ctx):
    """BASELINE config 5 at its own mesh size: two 2 097 152-tetrahedron spheres (one shared asset), millions of candidate
    tetrahedron pairs and ~27 000 contact polygons in ONE geometry pair: every face against the oracle's tree-driven
    run (MODE_TREE), emitted element pairs bit-exact, values to 1e-10; the pair-local offsets and the per-surface
    reduction over more than 4096 faces are exercised at this size."""
    from drake_b200 import workloads as W
    wl = W.config5(1)
    assert wl.meshes[0].num_elements == 2097152
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    st = batch.stats()
    osc, order = parity.oracle_scene_for(wl)
    want, ostats = osc.compute(wl.poses[0][order], ho.KPOLYGON, mode=ho.MODE_TREE)
    assert len(got[0]) == len(want) == 1
    assert want[0].num_faces > 20000
    parity.compare_surface(got[0][0], want[0])
    assert st["tet_tet_pairs"] == ostats["element_pairs"]       # candidate pair count, bit-exact
    assert st["tet_tet_pairs"] > 1_000_000


def test_config5_multi_chunk_equals_single_chunk(ctx):
    """Four full-size config-5 scenes walked in 4, 3 and 2 chunks of scenes (work lists reused, offsets carried on the
    device) give the counters and, face for face, the surfaces of the single-chunk query."""
    from drake_b200 import workloads as W
    wl = W.config5(4)
    batch = wl.register(ctx)
    poses = wl.poses.reshape(-1, 12)
    batch.set_chunks(1)
    base = batch.compute_contact_surfaces(poses, ho.KPOLYGON)
    base_stats = batch.stats()
    assert batch.num_chunks == 1
    assert sum(len(s) for s in base) >= 3
    for chunks in (4, 3, 2):
        batch.set_chunks(chunks)
        got = batch.compute_contact_surfaces(poses, ho.KPOLYGON)
        assert batch.num_chunks == chunks
        st = batch.stats()
        for k in ("bv_pair_tests", "leaf_pairs", "tet_tet_pairs", "polygons", "faces", "vertices", "surfaces"):
            assert st[k] == base_stats[k], (chunks, k)
        assert st["total_area"] == pytest.approx(base_stats["total_area"], rel=1e-12)
        for a, b in zip(got, base):
            assert [_digest(s) for s in a] == [_digest(s) for s in b]
            for x, y in zip(a, b):
                assert x.total_area == pytest.approx(y.total_area, rel=1e-12)
                assert np.sort(x.areas).sum() == pytest.approx(np.sort(y.areas).sum(), rel=1e-12)


def test_full_size_config4_batch_against_the_oracle_on_sampled_scenes(ctx):
    """BASELINE config 4 at its full size -- 4096 scenes x 37 geometries, 2.7 M potential geometry pairs, ~2 G candidate
    element pairs per query, the batch walked in chunks -- with every surface of ten scenes spread over the batch (first,
    last, both sides of chunk boundaries) compared with the scene-level oracle, and batch-level properties: surfaces
    sorted by (scene, id_M, id_N), per-scene statistics summing to the totals, no cull violation in the self-check."""
    import torch
    from drake_b200 import engine, workloads as W
    wl = W.config4(4096)
    batch = wl.register(ctx)
    poses = wl.poses.reshape(-1, 12)
    stats_dev = torch.zeros((wl.num_scenes, 6), dtype=torch.int64, device="cuda")
    batch.set_poses(poses)
    batch.run(ho.KPOLYGON, engine.STRICT, stats_dev.data_ptr())
    st = batch.stats()
    res = batch.copy_results()
    assert st["surfaces"] > 150_000 and st["tet_tet_pairs"] + st["tet_tri_pairs"] > 1_000_000_000
    key = res.surf_scene.astype(np.int64) * (1 << 40) + res.surf_id_M * (1 << 20) + res.surf_id_N
    assert np.all(np.diff(key) > 0)                      # (scene, id_M, id_N) order, no duplicates
    per_scene = stats_dev.cpu().numpy()
    assert per_scene[:, 3].sum() == st["faces"] and per_scene[:, 4].sum() == st["surfaces"]
    chunks = batch.num_chunks
    bounds = [int(4096 * c / chunks) for c in range(1, chunks)]
    chosen = sorted({0, 1, 4095, 2047, 2048, 1234, *bounds[:2], *[b - 1 for b in bounds[:2]]})[:10]
    mine = {}
    for s in res.surfaces(scenes=chosen):
        mine.setdefault(s.scene, []).append(s)
    osc, order = parity.oracle_scene_for(wl)
    compared = 0
    for sc in chosen:
        want, _ = osc.compute(wl.poses[sc][order], ho.KPOLYGON, mode=ho.MODE_TREE)
        got = mine.get(sc, [])
        assert [(x.id_M, x.id_N) for x in got] == [(x.id_M, x.id_N) for x in want], sc
        for g, w in zip(got, want):
            parity.compare_surface(g, w)
        compared += len(want)
    assert compared >= 200
    batch.set_verify(True)
    batch.run(ho.KPOLYGON, engine.STRICT)
    v = batch.stats()
    assert v["cull_violations"] == 0
    for k in ("polygons", "faces", "vertices", "surfaces"):
        assert v[k] == st[k], k
