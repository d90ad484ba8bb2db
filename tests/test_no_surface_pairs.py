"""The pairs that produce no contact surface -- rigid-rigid, a geometry with no hydroelastic representation, the
half-space combinations MaybeMakeContactSurface rejects (hydroelastic_calculator.cc:77-156) -- only matter when the
geometry-level broad phase reports them (FindCollisionCandidates, proximity_engine.cc:867-883,905-934): then the strict
query fails with the reference's result code and the WithFallback query hands them to the point-pair narrow phase
(:958-970). Far apart they do neither. Oracle (CPU) and the CUDA path (gpu) must agree on which pairs count."""
import math

import numpy as np
import pytest

from oracle import ho

R = 0.25


def _meshes():
    soft = ho.VolumeMesh.sphere(R, 0.25)
    return soft, ho.VolumeField.sphere(soft, R, 1e8), ho.TriMesh.sphere(R, 0.1)


def _scene(kinds, box=None):
    """kinds: per geometry 'soft' | 'rigid' | 'undef' | 'hs_rigid' | 'hs_soft', ids 1, 2, ..."""
    soft, fld, rigid = _meshes()
    sc = ho.Scene()
    for k, kind in enumerate(kinds):
        gid = k + 1
        if kind == "soft":
            sc.add_compliant(gid, soft, fld)
        elif kind == "rigid":
            sc.add_rigid(gid, rigid)
        elif kind == "undef":
            sc.add_undefined(gid, box)
        else:
            sc.add_half_space(gid, kind == "hs_soft", 1e6 if kind == "hs_soft" else 0.0, 0.0)
    return sc


TILT = ho.rpy_matrix(0.3, 0.0, 0.0)
CASES = [
    # kinds, poses (per geometry: rotation or None, translation), box of the undefined geometry, does the pair count?, code
    (("rigid", "rigid"), [(None, (0, 0, 0)), (None, (0, 0, 0.3))], None, True, "rigid-rigid"),
    (("rigid", "rigid"), [(None, (0, 0, 0)), (None, (0, 0, 3.0))], None, False, None),
    # Rotated: FCL boxes the bounding sphere of the local box (radius |h| = 0.25 sqrt(3) = 0.433): 0.8 apart still counts.
    (("rigid", "rigid"), [(TILT, (0, 0, 0)), (None, (0, 0, 0.65))], None, True, "rigid-rigid"),
    (("rigid", "rigid"), [(None, (0, 0, 0)), (None, (0, 0, 0.65))], None, False, None),
    (("soft", "undef"), [(None, (0, 0, 0)), (None, (0, 0, 5.0))], None, True, "without a hydroelastic"),   # no box: always
    (("soft", "undef"), [(None, (0, 0, 0)), (None, (0, 0, 5.0))], (0, 0, 0, 0.1, 0.1, 0.1), False, None),
    (("soft", "undef"), [(None, (0, 0, 0)), (None, (0, 0, 0.3))], (0, 0, 0, 0.1, 0.1, 0.1), True, "without a hydroelastic"),
    (("hs_rigid", "rigid"), [(None, (0, 0, 0)), (None, (0, 0, 1.0))], None, False, None),      # above an axis-aligned ground
    (("hs_rigid", "rigid"), [(None, (0, 0, 0)), (None, (0, 0, 0.2))], None, True, "rigid-rigid"),
    (("hs_rigid", "rigid"), [(TILT, (0, 0, 0)), (None, (0, 0, 1.0))], None, True, "rigid-rigid"),  # tilted: FCL box is everything
    (("hs_rigid", "hs_soft"), [(None, (0, 0, 0)), (TILT, (0, 0, -5.0))], None, True, "two half spaces"),
    (("hs_soft", "soft"), [(None, (0, 0, 0)), (None, (0, 0, 1.0))], None, False, None),
    (("hs_soft", "soft"), [(None, (0, 0, 0)), (None, (0, 0, 0.1))], None, True, "one of which is a half space"),
]


def _poses(spec):
    return np.stack([ho.xform(Rm, p) for Rm, p in spec])


@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_policy(case):
    kinds, spec, box, counts, msg = CASES[case]
    sc = _scene(kinds, box)
    P = _poses(spec)
    if counts:
        with pytest.raises(RuntimeError) as ei:
            sc.compute(P, ho.KPOLYGON)
        assert msg in str(ei.value)
        surfaces, _ = sc.compute(P, ho.KPOLYGON, with_fallback=True)
        assert surfaces == []
        np.testing.assert_array_equal(sc.fallback_pairs(P), [[1, 2]])
    else:
        surfaces, _ = sc.compute(P, ho.KPOLYGON)
        assert surfaces == []
        assert len(sc.fallback_pairs(P)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_gpu_policy_matches_the_oracle(case):
    import drake_b200 as db
    kinds, spec, box, counts, msg = CASES[case]
    ctx = db.Context(0)
    soft, fld, rigid = _meshes()
    assets = []
    for kind in kinds:
        if kind == "soft":
            assets.append(ctx.register_compliant_mesh(soft.verts, soft.tets, fld.values))
        elif kind == "rigid":
            assets.append(ctx.register_rigid_mesh(rigid.verts, rigid.tris))
        elif kind == "undef":
            assets.append(-1)
        else:
            assets.append(ctx.register_half_space(kind == "hs_soft", 1e6 if kind == "hs_soft" else 0.0, 0.0))
    batch = db.SceneBatch(ctx, [0, len(kinds)], assets, list(range(1, len(kinds) + 1)))
    if box is not None:
        boxes = np.full((len(kinds), 6), np.nan)
        boxes[kinds.index("undef")] = box
        batch.set_geometry_bounds(boxes)
    P = _poses(spec)
    batch.set_poses(P)
    codes = {"rigid-rigid": 6, "without a hydroelastic": 7, "two half spaces": 11, "one of which is a half space": 12}
    if counts:
        with pytest.raises(db.HydroError) as ei:
            batch.run(db.POLYGON, db.STRICT)
        assert ei.value.status == codes[msg]
    else:
        batch.run(db.POLYGON, db.STRICT)
        assert batch.sizes().num_surfaces == 0
    batch.run(db.POLYGON, db.WITH_FALLBACK)
    assert batch.sizes().num_surfaces == 0
    fp = batch.fallback_pairs()
    assert batch.stats()["fallback_pairs"] == len(fp) == (1 if counts else 0)
    if counts:
        assert list(fp[0][:3]) == [0, 0, 1]


@pytest.mark.gpu
def test_gpu_fallback_pairs_of_a_mixed_batch_match_the_oracle():
    """64 scenes of {rigid sphere, rigid sphere, soft sphere, undefined geometry with a box} at random poses: the
    surfaces of the WithFallback query and its fallback pair list (scene, ids) against the scene-level oracle."""
    import drake_b200 as db
    from tests import parity
    soft, fld, rigid = _meshes()
    ctx = db.Context(0)
    a_s, a_r = ctx.register_compliant_mesh(soft.verts, soft.tets, fld.values), ctx.register_rigid_mesh(rigid.verts, rigid.tris)
    S, ids = 64, [4, 2, 7, 9]
    batch = db.SceneBatch.replicated(ctx, S, [a_r, a_r, a_s, -1], ids)
    box = (0.01, -0.02, 0.0, 0.1, 0.15, 0.05)
    boxes = np.full((S * 4, 6), np.nan)
    boxes[3::4] = box
    batch.set_geometry_bounds(boxes)
    osc = ho.Scene()
    osc.add_rigid(4, rigid)
    osc.add_rigid(2, rigid)
    osc.add_compliant(7, soft, fld)
    osc.add_undefined(9, box)
    order = np.argsort(ids)
    rng = np.random.default_rng(5)
    P = np.stack([np.stack([parity.random_pose(rng, 0.35) if rng.uniform() < 0.6 else ho.xform(None, rng.uniform(-0.35, 0.35, 3))
                            for _ in range(4)]) for _ in range(S)])
    got = batch.compute_contact_surfaces(P.reshape(-1, 12), db.POLYGON, db.WITH_FALLBACK)
    fp = batch.fallback_pairs()
    gid = np.tile(np.asarray(ids), S)
    total = 0
    for s in range(S):
        want, _ = osc.compute(P[s][order], ho.KPOLYGON, with_fallback=True)
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want]
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w)
        mine = fp[fp[:, 0] == s]
        np.testing.assert_array_equal(np.stack([gid[mine[:, 1]], gid[mine[:, 2]]], axis=1).reshape(-1, 2),
                                      osc.fallback_pairs(P[s][order]).reshape(-1, 2))
        total += len(mine)
    assert 20 < total < 6 * S
