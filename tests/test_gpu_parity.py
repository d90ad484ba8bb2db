"""GPU-vs-oracle parity through the C-ABI (run on the B200 box with -m gpu)."""
import math

import numpy as np
import pytest

from oracle import ho
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import drake_b200 as db
    return db.Context(0)


@pytest.fixture(scope="module", autouse=True, params=["per_level", "on_chip"])
def mid_phase(request):
    """Every case of this module runs twice: through the per-level pipeline (k_traverse + k_filter, small-pair
    path off) and with every eligible geometry pair descended and culled on chip by one warp (k_small_pairs;
    by default only batches large enough to fill the device take it). Candidate lists are kept in both."""
    import drake_b200 as db
    db.SceneBatch.test_defaults = (0 if request.param == "per_level" else 1, True)
    yield request.param
    db.SceneBatch.test_defaults = None


def sphere_kat_scene():
    R = 0.25
    X_WA = ho.xform(ho.rpy_matrix(math.pi, 0, 0))
    X_WB = ho.xform(None, [0, 0, -(R + R) + 0.125])
    soft = ho.VolumeMesh.sphere(R, 0.25)
    return R, X_WA, X_WB, soft, ho.VolumeField.sphere(soft, R, 1e8)


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("ids", [(1, 2), (2, 1)])
def test_calculator_kat_soft_rigid(ctx, rep, ids):
    """hydroelastic_calculator_test.cc:320-355: 100 triangles / 28 polygons, both id orders."""
    import drake_b200 as db
    R, X_WA, X_WB, soft, fld = sphere_kat_scene()
    rigid = ho.TriMesh.sphere(R, 0.1)
    a_s, a_r = parity.register_compliant(ctx, soft, fld), parity.register_rigid(ctx, rigid)
    batch = db.SceneBatch(ctx, [0, 2], [a_s, a_r], list(ids))
    got = batch.compute_contact_surfaces(np.stack([X_WA, X_WB]), rep)[0]
    assert len(got) == 1
    assert got[0].num_faces == (100 if rep == ho.KTRIANGLE else 28)
    want = ho.soft_rigid(soft, fld, rigid, X_WA, X_WB, ids[0], ids[1], rep)
    parity.compare_surface(got[0], want)


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_calculator_kat_soft_soft(ctx, rep):
    """hydroelastic_calculator_test.cc:467-504: 12 triangles / 4 polygons."""
    import drake_b200 as db
    R, X_WA, X_WB, soft, fld = sphere_kat_scene()
    a = parity.register_compliant(ctx, soft, fld)
    batch = db.SceneBatch(ctx, [0, 2], [a, a], [1, 2])
    got = batch.compute_contact_surfaces(np.stack([X_WA, X_WB]), rep)[0]
    assert len(got) == 1
    assert got[0].num_faces == (12 if rep == ho.KTRIANGLE else 4)
    want = ho.soft_soft(soft, fld, soft, fld, X_WA, X_WB, 1, 2, rep)
    parity.compare_surface(got[0], want)


def test_octagon_and_heptagon_literals(ctx):
    """field_intersection_test.cc:196-259 and mesh_intersection_test.cc:660-709 through the full GPU path."""
    import drake_b200 as db
    m0 = ho.VolumeMesh.from_arrays([[2, 0, 2], [-2, 0, 2], [0, 2, -2], [0, -2, -2]], [[0, 1, 2, 3]])
    f0 = ho.VolumeField.from_values(m0, [8e4, 4e4, 4e4, 0])
    m1 = ho.VolumeMesh.from_arrays([[1.5, 1.5, 2.5], [-1.5, -1.5, 2.5], [-1.5, 1.5, -2.5], [1.5, -1.5, -2.5]],
                                   [[0, 1, 2, 3]])
    f1 = ho.VolumeField.from_values(m1, [7e4, 1e4, 4e4, 4e4])
    a0, a1 = parity.register_compliant(ctx, m0, f0), parity.register_compliant(ctx, m1, f1)
    batch = db.SceneBatch(ctx, [0, 2], [a0, a1], [1, 2])
    s = batch.compute_contact_surfaces(np.stack([ho.IDENTITY, ho.IDENTITY]), ho.KPOLYGON)[0][0]
    assert s.num_faces == 1 and s.face_data[0] == 8
    expect = np.array([[-0.5, -1, 0], [0.5, -1, 0], [1, -0.5, 0], [1, 0.5, 0], [0.5, 1, 0], [-0.5, 1, 0], [-1, 0.5, 0],
                       [-1, -0.5, 0]])
    np.testing.assert_allclose(s.vertices[s.polygons()[0]], expect, rtol=0, atol=1e-14)
    assert abs(s.total_area - 3.5) < 1e-14
    # Heptagon: a rigid triangle through a tet whose pressure grows with z... use a field with
    # gradient along +z so the face normal (+z) passes the filter.
    ft = ho.VolumeField.from_values(m0, [2.0, 2.0, -2.0, -2.0])  # p = z
    tri = ho.TriMesh.from_arrays([[1.5, 1.5, 0], [-1.5, 0, 0], [0, -1.5, 0]], [[0, 1, 2]])
    at, ar = parity.register_compliant(ctx, m0, ft), parity.register_rigid(ctx, tri)
    batch = db.SceneBatch(ctx, [0, 2], [at, ar], [1, 2])
    s = batch.compute_contact_surfaces(np.stack([ho.IDENTITY, ho.IDENTITY]), ho.KPOLYGON)[0][0]
    assert s.num_faces == 1 and s.face_data[0] == 7
    hept = np.array([[1, 1, 0], [0.5, 1, 0], [-1, 0.25, 0], [-1, -0.5, 0], [-0.5, -1, 0], [0.25, -1, 0], [1, 0.5, 0]])
    got = s.vertices[s.polygons()[0]]
    k = int(np.argmin(np.linalg.norm(got - hept[0], axis=1)))
    np.testing.assert_allclose(np.roll(got, -k, axis=0), hept, rtol=0, atol=1e-14)


def test_asset_derived_quantities_bitwise(ctx):
    """Per-tet gradient / value-at-origin / inward normals and per-triangle normals are bit-identical
    to the oracle's (mesh_field_linear.h:440-509, volume_mesh.cc:91-138)."""
    mS = ho.VolumeMesh.ellipsoid(3.01, 3.5, 4.0, 2.0, dense=True)
    fS = ho.VolumeField.ellipsoid(mS, 3.01, 3.5, 4.0, 1e5)
    a = parity.register_compliant(ctx, mS, fS)
    g, mo, nn = ctx.asset_derived(a)
    _, og, omo, _ = fS.arrays()
    np.testing.assert_array_equal(g, og)
    np.testing.assert_array_equal(mo, omo)
    np.testing.assert_array_equal(nn, mS.arrays()[2])
    mR = ho.TriMesh.sphere(3.0, 2.0)
    np.testing.assert_array_equal(ctx.asset_derived(parity.register_rigid(ctx, mR)), mR.arrays()[2])


BENCH_ARGS = [(0, 4, 0), (1, 4, 0), (2, 4, 0), (3, 4, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (2, 3, 0), (2, 4, 1),
              (2, 4, 2), (2, 4, 3), (2, 3, 1), (2, 2, 2)]


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_config1_mesh_intersection_benchmark(ctx, rep):
    """BASELINE config 1 (mesh_intersection_benchmark.cc:153-271): all 13 argument triples in one
    batch of 13 scenes; candidate sets bit-exact against the leaf predicate; surfaces to 1e-10."""
    import drake_b200 as db
    hints = [4.0, 3.0, 2.0, 1.0]
    trans = [(7, 7, 7), (4, 4, 4), (3.5, 3.5, 3.5), (1.2, 1.2, 1.2), (0, 0, 0)]
    assets, meshes = {}, {}
    for res in range(4):
        mS = ho.VolumeMesh.ellipsoid(3.01, 3.5, 4.0, hints[res], dense=True)
        fS = ho.VolumeField.ellipsoid(mS, 3.01, 3.5, 4.0, 1e5)
        mR = ho.TriMesh.sphere(3.0, hints[res])
        meshes[res] = (mS, fS, mR)
        assets[res] = (parity.register_compliant(ctx, mS, fS), parity.register_rigid(ctx, mR))
    geom_asset, poses = [], []
    for res, ov, rot in BENCH_ARGS:
        geom_asset += list(assets[res])
        poses += [ho.IDENTITY, ho.xform(ho.angle_axis_matrix(rot / 3.0 * math.pi / 4, [1, 1, 1]), trans[ov])]
    n = len(BENCH_ARGS)
    batch = db.SceneBatch(ctx, np.arange(n + 1) * 2, geom_asset, [1, 2] * n)
    got = batch.compute_contact_surfaces(np.stack(poses), rep)
    cand = batch.candidates(0)
    for k, (res, ov, rot) in enumerate(BENCH_ARGS):
        mS, fS, mR = meshes[res]
        want = parity.memo(("c1", rep, k), lambda: ho.soft_rigid(mS, fS, mR, poses[2 * k], poses[2 * k + 1], 1, 2, rep))
        if want is None:
            assert got[k] == []
        else:
            assert len(got[k]) == 1
            parity.compare_surface(got[k][0], want)
        want_c, _ = parity.memo(("c1c", k), lambda: ho.candidates(mS, mR, poses[2 * k + 1], ho.MODE_BRUTE))
        mine = cand[cand[:, 0] == k][:, 1:]
        mine = mine[np.lexsort((mine[:, 1], mine[:, 0]))]
        np.testing.assert_array_equal(mine, want_c)  # bit-exact candidate element-pair set


def test_config1_without_the_normal_filter_reproduces_the_benchmark_output(ctx):
    """mesh_intersection_benchmark.cc runs SampleVolumeFieldOnSurface with filter_face_normal_along_field_gradient = false
    (its rigid sphere sits INSIDE the compliant ellipsoid, where the filter rejects every face). With the filter switched
    off (dbh_batch_set_normal_filter) the four full-overlap cases give the areas the reference prints
    (mesh_intersection_benchmark.cc:90-93: 93.76, 93.76, 107.59, 111.67 m^2) and match the oracle face for face; with it on
    (ComputeContactSurfaces) they are empty."""
    import drake_b200 as db
    from drake_b200 import workloads as W
    wl = W.config1(args=[(0, 4, 0), (1, 4, 0), (2, 4, 0), (3, 4, 0)])
    poses = wl.poses.reshape(-1, 12)
    batch = wl.register(ctx)
    assert all(s == [] for s in batch.compute_contact_surfaces(poses, ho.KTRIANGLE))
    wl.normal_filter = False
    batch = wl.register(ctx, assets=wl.assets)
    got = batch.compute_contact_surfaces(poses, ho.KTRIANGLE)
    hints = [4.0, 3.0, 2.0, 1.0]
    for k, printed in enumerate([93.76, 93.76, 107.59, 111.67]):
        mS = ho.VolumeMesh.ellipsoid(3.01, 3.5, 4.0, hints[k], dense=True)
        fS = ho.VolumeField.ellipsoid(mS, 3.01, 3.5, 4.0, 1e5)
        mR = ho.TriMesh.sphere(3.0, hints[k])
        want = parity.memo(("c1nf", k), lambda: ho.soft_rigid(mS, fS, mR, wl.poses[k][0], wl.poses[k][1], 1, 2, ho.KTRIANGLE,
                                                              filter_face_normal=False))
        assert len(got[k]) == 1
        assert round(float(got[k][0].total_area), 2) == printed
        parity.compare_surface(got[k][0], want)


def test_config2_compliant_compliant(ctx):
    """BASELINE config 2 (compliant_mesh_intersection_benchmark.cc:112-202): ellipsoid x sphere,
    dense, resolutions 0..2 x overlaps 0..3, plus two dense spheres at random poses."""
    import drake_b200 as db
    hints = [4.0, 3.0, 2.0]
    trans = [(7, 7, 7), (4, 4, 4), (3.5, 3.5, 3.5), (1.2, 1.2, 1.2)]
    geom_asset, poses, meshes = [], [], []
    for h in hints:
        mS = ho.VolumeMesh.ellipsoid(3.01, 3.5, 4.0, h, dense=True)
        fS = ho.VolumeField.ellipsoid(mS, 3.01, 3.5, 4.0, 1e5)
        mR = ho.VolumeMesh.sphere(3.0, h, dense=True)
        fR = ho.VolumeField.sphere(mR, 3.0, 1e5)
        aS, aR = parity.register_compliant(ctx, mS, fS), parity.register_compliant(ctx, mR, fR)
        for t in trans:
            geom_asset += [aS, aR]
            poses += [ho.IDENTITY, ho.xform(None, t)]
            meshes.append((mS, fS, mR, fR))
    rng = np.random.default_rng(11)
    mA = ho.VolumeMesh.sphere(3.0, 1.0, dense=True)
    fA = ho.VolumeField.sphere(mA, 3.0, 1e5)
    aA = parity.register_compliant(ctx, mA, fA)
    for _ in range(3):
        geom_asset += [aA, aA]
        d = rng.normal(size=3)
        poses += [parity.random_pose(rng, 0.5), parity.random_pose(rng, 0.2, center=4.5 * d / np.linalg.norm(d))]
        meshes.append((mA, fA, mA, fA))
    n = len(meshes)
    batch = db.SceneBatch(ctx, np.arange(n + 1) * 2, geom_asset, [1, 2] * n)
    got = batch.compute_contact_surfaces(np.stack(poses), ho.KPOLYGON)
    cand = batch.candidates(1)
    nonnull = 0
    for k, (mS, fS, mR, fR) in enumerate(meshes):
        want = parity.memo(("c2", k), lambda: ho.soft_soft(mS, fS, mR, fR, poses[2 * k], poses[2 * k + 1], 1, 2,
                                                           ho.KPOLYGON, mode=ho.MODE_TREE))
        if want is None:
            assert got[k] == []
        else:
            nonnull += 1
            parity.compare_surface(got[k][0], want)
        X_MN = ho.invert_and_compose(poses[2 * k], poses[2 * k + 1])
        want_c, _ = parity.memo(("c2c", k), lambda: ho.candidates(mS, mR, X_MN, ho.MODE_TREE))
        mine = cand[cand[:, 0] == k][:, 1:]
        mine = mine[np.lexsort((mine[:, 1], mine[:, 0]))]
        np.testing.assert_array_equal(mine, want_c)
    assert nonnull >= 6


def test_scene_engine_contract(ctx):
    """geometry/test/proximity_engine_test.cc:1956-2249: ordering, filter, anchored x anchored,
    rigid-rigid error in the strict query and not in WithFallback."""
    import drake_b200 as db
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.25)
    fld = ho.VolumeField.sphere(soft, R, 1e8)
    rigid = ho.TriMesh.sphere(R, 0.1)
    a_s, a_r = parity.register_compliant(ctx, soft, fld), parity.register_rigid(ctx, rigid)
    ids = [30, 10, 20]
    poses = np.stack([ho.xform(None, [0, 0.2, 0.2]), ho.xform(None, [0, 0, 0]), ho.xform(None, [0, 0, 0.375])])
    batch = db.SceneBatch(ctx, [0, 3], [a_s, a_r, a_s], ids)
    got = batch.compute_contact_surfaces(poses, db.POLYGON)[0]
    assert [(s.id_M, s.id_N) for s in got] == [(10, 20), (10, 30), (20, 30)]
    osc = ho.Scene()
    osc.add_compliant(30, soft, fld)
    osc.add_rigid(10, rigid)
    osc.add_compliant(20, soft, fld)
    want, _ = osc.compute(poses[[1, 2, 0]], ho.KPOLYGON)
    for g, w in zip(got, want):
        parity.compare_surface(g, w)
    assert got[0].grad_eM is None and got[0].grad_eN is not None and got[0].swapped
    batch = db.SceneBatch(ctx, [0, 3], [a_s, a_r, a_s], ids, excluded=[[0, 1]])
    got = batch.compute_contact_surfaces(poses, db.POLYGON)[0]
    assert [(s.id_M, s.id_N) for s in got] == [(10, 20), (20, 30)]
    # Rigid-rigid.
    near = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(None, [0, 0, 0.3]), ho.xform(None, [0, 0, 10.0])])
    batch = db.SceneBatch(ctx, [0, 3], [a_r, a_r, a_s], [1, 2, 3])
    batch.set_poses(near)
    with pytest.raises(db.HydroError) as ei:
        batch.run(db.POLYGON, db.STRICT)
    assert ei.value.status == 6 and "rigid-rigid" in str(ei.value)
    batch.run(db.POLYGON, db.WITH_FALLBACK)
    assert batch.sizes().num_surfaces == 0
    apart = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(None, [0, 0, 3.0]), ho.xform(None, [0, 0, 0.3])])
    got = batch.compute_contact_surfaces(apart, db.POLYGON)[0]
    assert [(s.id_M, s.id_N) for s in got] == [(1, 3)]
    batch = db.SceneBatch(ctx, [0, 3], [a_r, a_r, a_s], [1, 2, 3], geom_anchored=[1, 1, 0])
    both = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(None, [0, 0, 0.6]), ho.xform(None, [0, 0, 0.3])])
    got = batch.compute_contact_surfaces(both, db.POLYGON)[0]
    assert [(s.id_M, s.id_N) for s in got] == [(1, 3), (2, 3)]
    # Undefined geometry: strict query fails, fallback ignores it.
    batch = db.SceneBatch(ctx, [0, 2], [a_s, -1], [1, 2])
    batch.set_poses(near[:2])
    with pytest.raises(db.HydroError) as ei:
        batch.run(db.POLYGON, db.STRICT)
    assert ei.value.status == 7
    batch.run(db.POLYGON, db.WITH_FALLBACK)


def test_empty_and_degenerate_inputs(ctx):
    import drake_b200 as db
    R, X_WA, X_WB, soft, fld = sphere_kat_scene()
    a = parity.register_compliant(ctx, soft, fld)
    # A scene with a single geometry has no pairs.
    batch = db.SceneBatch(ctx, [0, 1], [a], [1])
    assert batch.compute_contact_surfaces(np.stack([X_WA]), db.POLYGON) == [[]]
    # Ragged batch: scenes with 1, 2 and 0 geometries.
    batch = db.SceneBatch(ctx, [0, 1, 3, 3], [a, a, a], [1, 1, 2])
    got = batch.compute_contact_surfaces(np.stack([X_WA, X_WA, X_WB]), db.POLYGON)
    assert [len(g) for g in got] == [0, 1, 0] and got[1][0].num_faces == 4
    # Degenerate tetrahedron is rejected at registration (MeshGradientMode::kOkOrThrow).
    with pytest.raises(db.HydroError) as ei:
        ctx.register_compliant_mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], [[0, 1, 2, 3]], [0, 0, 0, 1])
    assert ei.value.status == 4
    # A tetrahedron with zero pressure gradient passes both filters.
    flat = ho.VolumeField.from_values(soft, np.zeros(soft.num_vertices))
    rigid = ho.TriMesh.sphere(R, 0.1)
    af, ar = parity.register_compliant(ctx, soft, flat), parity.register_rigid(ctx, rigid)
    batch = db.SceneBatch(ctx, [0, 2], [af, ar], [1, 2])
    got = batch.compute_contact_surfaces(np.stack([X_WA, X_WB]), db.POLYGON)[0]
    want = ho.soft_rigid(soft, flat, rigid, X_WA, X_WB, 1, 2, ho.KPOLYGON)
    parity.compare_surface(got[0], want)


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_config3_ball_plate_scenes(ctx, rep):
    """BASELINE config 3: 24 of the batched ball / plate / floor scenes against the scene-level oracle
    (ids, ordering, every surface); soft-rigid and soft-soft pairs in one batch, anchored floor."""
    from drake_b200 import workloads as W
    wl = W.config3(24)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), rep)
    osc, order = parity.oracle_scene_for(wl)
    total = 0
    for s in range(wl.num_scenes):
        want = parity.memo(("c3", rep, s), lambda: osc.compute(wl.poses[s][order], rep)[0])
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want]
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w)
        total += len(want)
    assert total >= 40


def test_config4_clutter_scenes_including_coplanar_clips(ctx):
    """BASELINE config 4: 3 scenes of 32 soft boxes / ellipsoids in a rigid bin. The grid box meshes
    have boundary tetrahedra with zero pressure gradient, whose equilibrium plane against another
    body is coplanar with that body's boundary faces: intermediate polygons outgrow 8 vertices and
    take the exceptional path; results must still match the oracle face for face."""
    from drake_b200 import workloads as W
    wl = W.config4(3)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    osc, order = parity.oracle_scene_for(wl)
    for s in range(wl.num_scenes):
        want = parity.memo(("c4", s), lambda: osc.compute(wl.poses[s][order], ho.KPOLYGON)[0])
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want]
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w)
        assert len(want) >= 20


@pytest.mark.parametrize("config,scenes", [("c3", 64), ("c4", 4), ("c2s", 4)])
def test_conservative_culls_never_reject_an_emitting_candidate(ctx, config, scenes):
    """Self-check mode (dbh_batch_set_verify): the reference's exact clipping sequence runs for every
    candidate element pair; no candidate the fp32 cull stage rejects may emit a polygon, and the
    polygon / face / vertex totals must equal those of the normal (culled) pipeline."""
    from drake_b200 import engine, workloads as W
    wl = W.CONFIGS[config](scenes)
    batch = wl.register(ctx)
    batch.set_poses(wl.poses.reshape(-1, 12))
    batch.run(ho.KPOLYGON, engine.STRICT)
    normal = batch.stats()
    batch.set_verify(True)
    batch.run(ho.KPOLYGON, engine.STRICT)
    checked = batch.stats()
    assert checked["cull_violations"] == 0
    assert checked["tet_tri_full_clips"] == checked["tet_tri_pairs"]
    assert checked["tet_tet_full_clips"] == checked["tet_tet_pairs"]
    for k in ("polygons", "faces", "vertices", "surfaces", "tet_tri_pairs", "tet_tet_pairs"):
        assert normal[k] == checked[k], k
    assert normal["polygons"] > 0
    # The cull stage must actually cull (otherwise this test checks nothing).
    assert normal["tet_tri_full_clips"] + normal["tet_tet_full_clips"] < normal["tet_tri_pairs"] + normal["tet_tet_pairs"]


def _surface_digest(surfaces):
    """Order-independent digest of a scene's surfaces (faces inside a surface come in atomic order)."""
    out = []
    for s in surfaces:
        keys = np.sort(parity.face_keys(s.elem_M, s.elem_N))
        out.append((s.id_M, s.id_N, s.num_faces, len(s.vertices), keys.tobytes(), float(s.total_area),
                    float(np.sort(s.areas).sum()), float(np.sort(s.pressures).sum())))
    return out


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_chunked_query_equals_single_chunk(ctx, rep):
    """A query that walks the batch in several chunks of scenes (work lists reused) returns the same
    surfaces, face for face, as the single-chunk query; explicit chunk counts incl. more chunks than scenes."""
    from drake_b200 import workloads as W
    wl = W.config3(37)
    batch = wl.register(ctx)
    poses = wl.poses.reshape(-1, 12)
    want = [_surface_digest(s) for s in batch.compute_contact_surfaces(poses, rep)]
    base = batch.stats()
    for chunks in (2, 5, 37, 1000):
        batch.set_chunks(chunks)
        assert batch.num_chunks == min(chunks, 37)
        got = [_surface_digest(s) for s in batch.compute_contact_surfaces(poses, rep)]
        st = batch.stats()
        for k in ("bv_pair_tests", "tet_tri_pairs", "tet_tet_pairs", "polygons", "faces", "vertices", "surfaces", "leaf_pairs"):
            assert st[k] == base[k], (chunks, k)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert len(g) == len(w)
            for a, b in zip(g, w):
                assert a[:5] == b[:5]
                assert a[5] == pytest.approx(b[5], rel=1e-12)
                assert a[6] == pytest.approx(b[6], rel=1e-12)
                assert a[7] == pytest.approx(b[7], rel=1e-12)


def test_work_lists_stay_within_the_memory_budget_by_chunking(ctx, monkeypatch):
    """With a tiny work-list budget the batch splits itself into chunks of scenes instead of growing
    the lists; results still match the oracle."""
    from drake_b200 import workloads as W
    monkeypatch.setenv("DBH_WORK_BUDGET_MB", "16")
    wl = W.config3(512)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    assert batch.num_chunks > 1
    osc, order = parity.oracle_scene_for(wl)
    for s in (0, 100, 255, 256, 511):
        want, _ = osc.compute(wl.poses[s][order], ho.KPOLYGON)
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want]
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w)
    # Asking for a single chunk by hand: the lists overflow, and since the budget still holds the batch splits
    # itself again (run() handles the capacity errors); nothing changes in the results.
    batch.set_chunks(1)
    again = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    assert batch.num_chunks > 1
    for s in (0, 255, 511):
        for g, w in zip(_surface_digest(again[s]), _surface_digest(got[s])):
            assert g[:5] == w[:5]


def test_polygon_with_more_than_eight_vertices(ctx):
    """C4 scene 7280: a clip plane coplanar with the polygon makes the tet-tet clip return a 9-gon
    (the reference's vectors have no bound; its "at most 8" comment assumes general position). The
    exceptional path stages it in two slots; the surface must match the oracle face for face."""
    from drake_b200 import workloads as W
    wl = W.config4(1, 7280)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)[0]
    assert batch.stats()["exceptional_clips"] >= 1
    osc, order = parity.oracle_scene_for(wl)
    want, _ = osc.compute(wl.poses[0][order], ho.KPOLYGON)
    assert max(len(p) for s in want for p in s.polygons()) == 9
    assert [(x.id_M, x.id_N) for x in got] == [(x.id_M, x.id_N) for x in want]
    for g, w in zip(got, want):
        parity.compare_surface(g, w)
    # Same scene in the triangle representation (a 9-fan).
    got_t = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KTRIANGLE)[0]
    want_t, _ = osc.compute(wl.poses[0][order], ho.KTRIANGLE)
    for g, w in zip(got_t, want_t):
        parity.compare_surface(g, w)


def _check_bvh(boxes, left, num_elements):
    """Structure + containment of a BVH in the library's layout."""
    seen = np.zeros(num_elements, int)
    stack = [1]
    depth = {1: 0}
    height = 0
    while stack:
        k = stack.pop()
        if left[k] < 0:
            seen[~left[k]] += 1
            height = max(height, depth[k])
            continue
        for c in (left[k], left[k] + 1):
            lo_p, hi_p = boxes[k, :3] - boxes[k, 3:], boxes[k, :3] + boxes[k, 3:]
            lo_c, hi_c = boxes[c, :3] - boxes[c, 3:], boxes[c, :3] + boxes[c, 3:]
            tol = 1e-12 * (1 + np.abs(boxes[k]).max())
            assert np.all(lo_c >= lo_p - tol) and np.all(hi_c <= hi_p + tol), (k, c)
            depth[c] = depth[k] + 1
            stack.append(c)
    assert np.all(seen == 1)
    return height


def test_gpu_lbvh_builder_gives_the_same_query_results():
    """The device-built linear BVH (Morton sort + Karras + bottom-up fit): a valid tree over every
    element with leaf boxes bit-identical to the host builder's, and -- the candidate set being
    tree-independent -- bit-identical candidate sets and matching surfaces."""
    import drake_b200 as db
    from drake_b200 import engine
    R = 3.0
    soft = ho.VolumeMesh.sphere(R, 1.0)                 # 4096 tets, dense
    fld = ho.VolumeField.sphere(soft, R, 1e5)
    rigid = ho.TriMesh.sphere(R, 1.0)
    poses = np.stack([ho.xform(ho.rpy_matrix(0.3, -0.2, 0.9), [0.1, 0.2, 0.3]), ho.xform(ho.rpy_matrix(1.0, 0.4, 0.2), [2.5, 1.0, 0.5]),
                      ho.xform(None, [-1.5, 2.0, 0.5])])
    out = []
    for builder in (engine.BVH_HOST_MEDIAN_SPLIT, engine.BVH_GPU_LBVH):
        c = db.Context(0)
        c.set_bvh_builder(builder)
        a_s = parity.register_compliant(c, soft, fld)
        a_r = parity.register_rigid(c, rigid)
        for a, ne in ((a_s, len(soft.tets)), (a_r, len(rigid.tris))):
            boxes, left = c.asset_bvh(a)
            assert len(left) == 2 * ne
            h = _check_bvh(boxes, left, ne)
            assert h == c.asset_info(a).bvh_height
        boxes, left = c.asset_bvh(a_s)
        leaf = left < 0
        leaf[0] = False
        order = np.argsort(~left[leaf])
        out_boxes = boxes[leaf][order]
        batch = db.SceneBatch(c, [0, 3], [a_s, a_s, a_r], [1, 2, 3])
        got = batch.compute_contact_surfaces(poses, db.POLYGON)[0]
        cands = [batch.candidates(k) for k in (0, 1)]
        cands = [x[np.lexsort(x.T[::-1])] for x in cands]
        out.append((out_boxes, got, cands, batch.stats()))
    (b0, s0, c0, st0), (b1, s1, c1, st1) = out
    np.testing.assert_array_equal(b0, b1)               # leaf boxes bit-identical
    for x, y in zip(c0, c1):
        np.testing.assert_array_equal(x, y)             # candidate sets bit-identical
    assert len(c0[0]) > 1000 and len(c0[1]) > 1000
    assert [(s.id_M, s.id_N) for s in s0] == [(s.id_M, s.id_N) for s in s1] == [(1, 2), (1, 3), (2, 3)]
    for g, w in zip(_surface_digest(s0), _surface_digest(s1)):
        assert g[:5] == w[:5]
        assert g[5] == pytest.approx(w[5], rel=1e-12)
    for k in ("polygons", "faces", "vertices", "tet_tri_pairs", "tet_tet_pairs"):
        assert st0[k] == st1[k]


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_large_surfaces_reduced_by_the_whole_grid(ctx, rep):
    """Two dense spheres of 32768 tets: one surface of thousands of faces per scene; above 4096 faces
    the per-surface area / centroid reduction is spread over the grid. Totals must match the oracle."""
    from drake_b200 import workloads as W
    wl = W.config2_spheres(2, 0, 0.5)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), rep)
    osc, order = parity.oracle_scene_for(wl)
    biggest = 0
    for s in range(wl.num_scenes):
        want, _ = osc.compute(wl.poses[s][order], rep)
        assert len(got[s]) == len(want) == 1
        parity.compare_surface(got[s][0], want[0])
        biggest = max(biggest, want[0].num_faces)
    if rep == ho.KTRIANGLE:
        assert biggest > 4096


@pytest.mark.parametrize("name", ["c3_16_scenes", "c3_16_scenes_tri", "c4_2_scenes", "c2s_2_scenes"])
def test_against_committed_golden_fixtures(ctx, name):
    """GPU results against tests/golden/*.json (no oracle call): ids, face / vertex counts, the checksum of the
    emitting element-pair keys (bit-exact), total area, centroid and summed face pressure (1e-10)."""
    import json
    import os
    from tests.golden import make_golden as G
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(G.__file__)), name + ".json")))
    make, rep = G.CASES[name]
    wl = make()
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), rep)
    assert len(got) == len(want["scenes"])
    for gs, ws in zip(got, want["scenes"]):
        assert [(g.id_M, g.id_N, g.num_faces, len(g.vertices)) for g in gs] == \
               [(w["id_M"], w["id_N"], w["faces"], w["vertices"]) for w in ws]
        for g, w in zip(gs, ws):
            keys = parity.face_keys(g.elem_M, g.elem_N) * np.int64(2654435761) % np.int64(1 << 61)
            assert int(np.bitwise_xor.reduce(keys)) == w["key_checksum"]
            assert g.total_area == pytest.approx(w["total_area"], rel=1e-10)
            assert float(np.sum(g.pressure_at_centroid)) == pytest.approx(w["sum_face_pressure"], rel=1e-10)
            assert np.allclose(g.centroid, w["centroid"], rtol=1e-10, atol=1e-13)


def test_full_size_config3_properties(ctx):
    """BASELINE config 3 at its full size (16 384 scenes, 49 152 geometry pairs, ~13.6 M candidate element pairs):
    size-independent properties instead of an oracle run. (1) idempotence: the same query twice gives the same
    counters; (2) the per-scene statistics sum to the batch totals; (3) a chunked query gives the same counters and
    total area; (4) moving every geometry of every scene by one rigid transform leaves the number of surfaces
    unchanged and the total contact area within 1e-9; (5) the self-check mode finds no cull violation."""
    import torch
    from drake_b200 import engine, workloads as W
    wl = W.config3(16384)
    batch = wl.register(ctx)
    poses = wl.poses.reshape(-1, 12)
    stats_dev = torch.zeros((wl.num_scenes, 6), dtype=torch.int64, device="cuda")
    batch.set_poses(poses)
    batch.run(ho.KPOLYGON, engine.STRICT, stats_dev.data_ptr())
    a = batch.stats()
    per_scene = stats_dev.cpu().numpy()
    batch.run(ho.KPOLYGON, engine.STRICT)
    b = batch.stats()
    counters = ("bv_pair_tests", "leaf_pairs", "tet_tri_pairs", "tet_tet_pairs", "polygons", "faces", "vertices", "surfaces")
    for k in counters:
        assert a[k] == b[k], k
    assert a["total_area"] == pytest.approx(b["total_area"], rel=1e-12)
    assert a["surfaces"] > 40000 and a["tet_tri_pairs"] + a["tet_tet_pairs"] > 13_000_000
    assert per_scene[:, 3].sum() == a["faces"] and per_scene[:, 4].sum() == a["surfaces"]
    assert per_scene[:, 0].sum() == a["geometry_pairs_tested"]
    batch.set_chunks(5)
    batch.run(ho.KPOLYGON, engine.STRICT)
    c = batch.stats()
    for k in counters:
        assert a[k] == c[k], k
    assert a["total_area"] == pytest.approx(c["total_area"], rel=1e-12)
    batch.set_chunks(1)
    # (4) one rigid transform applied to the whole world.
    X = parity.random_pose(np.random.default_rng(7), 0.3)
    R, p = X[:9].reshape(3, 3).T, X[9:]
    Rg = poses[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    moved = np.concatenate([(R @ Rg).transpose(0, 2, 1).reshape(-1, 9), poses[:, 9:] @ R.T + p], axis=1)
    batch.set_poses(moved)
    batch.run(ho.KPOLYGON, engine.STRICT)
    d = batch.stats()
    assert d["surfaces"] == a["surfaces"]
    assert abs(d["faces"] - a["faces"]) <= 1e-4 * a["faces"]          # ulp-level classifications may flip
    assert d["total_area"] == pytest.approx(a["total_area"], rel=1e-9)
    # (5)
    batch.set_poses(poses)
    batch.set_verify(True)
    batch.run(ho.KPOLYGON, engine.STRICT)
    v = batch.stats()
    assert v["cull_violations"] == 0
    for k in ("polygons", "faces", "vertices", "surfaces"):
        assert v[k] == a[k], k
