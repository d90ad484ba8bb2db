"""The surface-seeded soft-soft intersector on the B200 (dbh_context_set_soft_soft_intersector(DBH_INTERSECTOR_SURFACE_BFS);
hydro_bfs.cuh / hydro_bfs_kernels.cuh: boundary trees for the seeds, breadth-first waves over the face neighbours) against
the oracle, through the C-ABI. It is the reference's second VolumeIntersector::IntersectFields overload
(field_intersection.cc:254-375); the reference's own tests for it compare the emitted tetrahedron pairs with the volume
overload's (field_intersection_test.cc:397-476) and pin the one known difference (FullyPenetrated, :296-344) -- so do
these. Every case runs through both mid-phase implementations (the seeds come from either)."""
import numpy as np
import pytest

from oracle import ho
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import drake_b200 as db
    from drake_b200 import engine
    c = db.Context(0)
    c.set_soft_soft_intersector(engine.INTERSECTOR_SURFACE_BFS)
    return c


@pytest.fixture(scope="module")
def volume_ctx():
    import drake_b200 as db
    return db.Context(0)


@pytest.fixture(scope="module", autouse=True, params=["per_level", "on_chip"])
def mid_phase(request):
    import drake_b200 as db
    db.SceneBatch.test_defaults = (0 if request.param == "per_level" else 1, False)
    yield request.param
    db.SceneBatch.test_defaults = None


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_dense_spheres_against_the_oracle(ctx, rep):
    """Two dense spheres (4096 tetrahedra each) in six scenes of one batch, random relative poses from grazing to deep:
    every face of every surface against the oracle's volume-hierarchy run, emitted element pairs bit-exact."""
    import drake_b200 as db
    s = ho.VolumeMesh.sphere(1.0, 0.25, dense=True)
    f = ho.VolumeField.sphere(s, 1.0, 1e5)
    a = parity.register_compliant(ctx, s, f)
    rng = np.random.default_rng(21)
    poses = []
    for k in range(6):
        d = rng.normal(size=3)
        d *= (1.1 + 0.16 * k) / np.linalg.norm(d)
        poses += [parity.random_pose(rng, 0.05), parity.random_pose(rng, 0.05, center=d)]
    batch = db.SceneBatch(ctx, np.arange(7) * 2, [a, a] * 6, [1, 2] * 6)
    got = batch.compute_contact_surfaces(np.stack(poses), rep)
    st = batch.stats()
    nonnull = 0
    for k in range(6):
        want = parity.memo(("bfs_spheres", k, rep), lambda: ho.soft_soft(s, f, s, f, poses[2 * k], poses[2 * k + 1], 1, 2,
                                                                         rep, mode=ho.MODE_TREE))
        if want is None:
            assert got[k] == []
            continue
        nonnull += 1
        assert len(got[k]) == 1
        parity.compare_surface(got[k][0], want)
    assert nonnull >= 5
    assert st["bfs_waves"] > 0 and st["bfs_pairs"] >= st["polygons"]
    # Work proportional to the surface: far fewer pairs reach the exact stage than the volume hierarchies would offer.
    cand = sum(len(ho.candidates(s, s, ho.invert_and_compose(poses[2 * k], poses[2 * k + 1]), ho.MODE_TREE)[0])
               for k in range(2))
    assert st["bfs_pairs"] < 3 * cand


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_the_reference_fixture_pad_and_octahedron(ctx, rep):
    """field_intersection_test.cc:346-476: Box(0.06, 0.10, 0.14) with medial axis x the octahedron of Sphere(0.03),
    X_MN = 0.03 x (and the no-intersection pose 0.2 x, :532-560), both representations, both id orders."""
    import drake_b200 as db
    box = ho.VolumeMesh.box_ma(0.06, 0.10, 0.14)
    bf = ho.VolumeField.box(box, 0.06, 0.10, 0.14, 1e5)
    octa = ho.VolumeMesh.sphere(0.03, 0.3)
    of = ho.VolumeField.sphere(octa, 0.03, 1e5)
    a_box, a_octa = parity.register_compliant(ctx, box, bf), parity.register_compliant(ctx, octa, of)
    X_WM, X_WN, X_far = ho.xform(), ho.xform(None, [0.03, 0, 0]), ho.xform(None, [0.2, 0, 0])
    batch = db.SceneBatch(ctx, [0, 2, 4, 6], [a_box, a_octa, a_octa, a_box, a_box, a_octa], [1, 2, 1, 2, 1, 2])
    got = batch.compute_contact_surfaces(np.stack([X_WM, X_WN, X_WN, X_WM, X_WM, X_far]), rep)
    want = ho.soft_soft(box, bf, octa, of, X_WM, X_WN, 1, 2, rep, mode=ho.MODE_BRUTE)
    assert want is not None and len(got[0]) == 1
    parity.compare_surface(got[0][0], want)
    want = ho.soft_soft(octa, of, box, bf, X_WN, X_WM, 1, 2, rep, mode=ho.MODE_BRUTE)
    assert len(got[1]) == 1
    parity.compare_surface(got[1][0], want)
    assert got[2] == []


def test_fully_penetrated_body_is_a_subset_of_the_volume_result(ctx, volume_ctx):
    """field_intersection_test.cc:296-344: a small sphere well inside a big one, surfaces apart. The reference's surface
    overload finds no seed (tight OBB leaves) and reports nothing, "not a desirable behavior" in its own words; the
    device seeds through per-triangle axis-aligned boxes, which may reach the inner body. Whatever it reports is made of
    polygons of the volume result (same emitting pairs, same values): never more, possibly fewer."""
    import drake_b200 as db
    big = ho.VolumeMesh.sphere(1.0, 1.0)
    bigf = ho.VolumeField.sphere(big, 1.0, 1e5)
    small = ho.VolumeMesh.sphere(0.125, 0.125)
    smallf = ho.VolumeField.sphere(small, 0.125, 1e5)
    poses = np.stack([ho.xform(), ho.xform(None, [0.5, 0, 0])])
    results = []
    for c in (ctx, volume_ctx):
        a_big, a_small = parity.register_compliant(c, big, bigf), parity.register_compliant(c, small, smallf)
        batch = db.SceneBatch(c, [0, 2], [a_big, a_small], [1, 2])
        results.append(batch.compute_contact_surfaces(poses, ho.KPOLYGON)[0])
    assert len(results[1]) == 1
    full = results[1][0]
    parity.compare_surface(full, ho.soft_soft(big, bigf, small, smallf, poses[0], poses[1], 1, 2, ho.KPOLYGON))
    if results[0]:
        part = results[0][0]
        fk, pk = parity.face_keys(full.elem_M, full.elem_N), parity.face_keys(part.elem_M, part.elem_N)
        assert len(np.unique(pk)) == len(pk) and np.all(np.isin(pk, fk))
        idx = {k: i for i, k in enumerate(fk.tolist())}
        sel = np.array([idx[k] for k in pk.tolist()])
        np.testing.assert_array_equal(part.areas, full.areas[sel])


def test_mixed_scene_batch_against_the_scene_oracle(ctx):
    """The ball / plate / floor batch (soft-rigid and soft-soft pairs side by side, 96 scenes): every surface of every
    scene against the scene-level oracle; only the soft-soft pairs go through the surface intersector."""
    from drake_b200 import workloads as W
    wl = W.config3(96)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    st = batch.stats()
    osc, order = parity.oracle_scene_for(wl)
    compared = 0
    for sc in range(wl.num_scenes):
        want, _ = parity.memo(("bfs_c3", sc), lambda: osc.compute(wl.poses[sc][order], ho.KPOLYGON))
        assert [(x.id_M, x.id_N) for x in got[sc]] == [(x.id_M, x.id_N) for x in want], sc
        for g, w in zip(got[sc], want):
            parity.compare_surface(g, w)
        compared += len(want)
    assert compared >= 150 and st["bfs_waves"] > 0


def test_multi_chunk_equals_single_chunk(ctx):
    """The queue and the visited set are chunk-local: the same batch walked in 1, 2 and 5 chunks of scenes."""
    import drake_b200 as db
    s = ho.VolumeMesh.sphere(1.0, 0.5, dense=True)
    f = ho.VolumeField.sphere(s, 1.0, 1e5)
    a = parity.register_compliant(ctx, s, f)
    rng = np.random.default_rng(4)
    n = 10
    poses = []
    for _ in range(n):
        d = rng.normal(size=3)
        d *= rng.uniform(1.2, 1.9) / np.linalg.norm(d)
        poses += [parity.random_pose(rng, 0.05), parity.random_pose(rng, 0.05, center=d)]
    batch = db.SceneBatch(ctx, np.arange(n + 1) * 2, [a, a] * n, [1, 2] * n)
    base = None
    for chunks in (1, 2, 5):
        batch.set_chunks(chunks)
        got = batch.compute_contact_surfaces(np.stack(poses), ho.KPOLYGON)
        digest = [[(x.id_M, x.id_N, x.num_faces, np.sort(parity.face_keys(x.elem_M, x.elem_N)).tobytes()) for x in sc]
                  for sc in got]
        if base is None:
            base = digest
            assert sum(len(sc) for sc in got) >= 8
        assert digest == base, chunks


def test_one_full_size_config5_scene_equals_the_volume_intersector(ctx, volume_ctx):
    """BASELINE config 5 at its own mesh size (two 2 097 152-tetrahedron spheres): the surface intersector's result --
    boundary trees of 32 768 triangles, tens of waves -- equals the volume intersector's face for face (which
    tests/test_zz_gpu_fullscale.py compares with the oracle), while a small fraction of its pairs reach the exact stage."""
    from drake_b200 import workloads as W
    wl = W.config5(1)
    results, stats = [], []
    for c in (ctx, volume_ctx):
        batch = wl.register(c)
        results.append(batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)[0])
        stats.append(batch.stats())
    (got,), (want,) = results
    assert want.num_faces > 20000 and got.num_faces == want.num_faces
    gk, wk = parity.face_keys(got.elem_M, got.elem_N), parity.face_keys(want.elem_M, want.elem_N)
    go, wo = np.argsort(gk, kind="stable"), np.argsort(wk, kind="stable")
    np.testing.assert_array_equal(gk[go], wk[wo])
    np.testing.assert_array_equal(got.areas[go], want.areas[wo])            # the same arithmetic: bit for bit
    np.testing.assert_array_equal(got.centroids[go], want.centroids[wo])
    assert got.total_area == pytest.approx(want.total_area, rel=1e-12)
    assert stats[0]["bfs_pairs"] * 10 < stats[1]["tet_tet_pairs"]


def test_deforming_volume_meshes_through_the_surface_intersector(ctx):
    """Deformable x deformable (deformable_field_intersection.cc:46-66) on the surface intersector: the device refit
    re-derives the tetrahedra and refits BOTH trees of an asset (the boundary tree's leaves take their tetrahedron's box:
    conservative), the face-neighbour table is topological and stays. After every update the surface equals the oracle's
    on the deformed meshes face for face."""
    import drake_b200 as db
    from drake_b200 import engine
    from oracle import ho_deformable as hd
    from tests.test_refit import _sdf_spheres, deform
    a, sa, b, sb, off = _sdf_spheres()
    va0, vb0 = a.verts.copy(), b.verts.copy()
    ha = ctx.register_compliant_mesh(va0, a.tets, sa)
    hb = ctx.register_compliant_mesh(vb0 + off, b.tets, sb)
    batch = db.SceneBatch(ctx, [0, 2], [ha, hb], [1, 2])
    poses = np.stack([ho.IDENTITY, ho.IDENTITY])
    faces = 0
    for t in (0.0, 0.5, -0.6):
        va, vb = deform(va0, t), deform(vb0, -t) + off
        ctx.update_vertices([ha, hb], [va, vb])
        ma, mb = ho.VolumeMesh.from_arrays(va, a.tets), ho.VolumeMesh.from_arrays(vb, b.tets)
        want = hd.deformable_deformable_contact(ma, sa, mb, sb)
        assert want is not None
        batch.set_poses(poses)
        batch.run(ho.KPOLYGON, engine.STRICT)
        got = batch.copy_results().surfaces()
        assert len(got) == 1
        parity.compare_surface(got[0], want.surface)
        faces += got[0].num_faces
    assert faces > 100 and batch.stats()["bfs_waves"] > 0


def test_registration_builds_the_reference_topology(ctx, volume_ctx):
    """The face-neighbour table registration builds (sorted faces) equals VolumeMeshTopology (volume_mesh_topology.cc:13-68,
    the oracle's restatement, itself pinned by the reference's two fixtures) entry for entry -- which neighbour behind
    which face index -- and the boundary tree has one leaf per triangle of the surface mesh."""
    from drake_b200 import engine
    for mesh in (ho.VolumeMesh.sphere(1.0, 0.25, dense=True), ho.VolumeMesh.box_ma(0.06, 0.10, 0.14),
                 ho.VolumeMesh.ellipsoid(0.3, 0.2, 0.1, 0.05), ho.VolumeMesh.box_grid(0.3, 0.2, 0.1, 0.05)):
        a = ctx.register_compliant_mesh(mesh.verts, mesh.tets, np.zeros(len(mesh.verts)) + mesh.verts[:, 0])
        nb, ntri = ctx.asset_topology(a)
        np.testing.assert_array_equal(nb, mesh.neighbors())
        assert ntri == len(mesh.to_surface().tris) == int((nb < 0).sum())
    plain = volume_ctx.register_compliant_mesh(mesh.verts, mesh.tets, mesh.verts[:, 0].copy())
    with pytest.raises(engine.HydroError):
        volume_ctx.asset_topology(plain)


def test_dense_clutter_is_a_subset_of_the_volume_result(ctx, volume_ctx):
    """Config 4 (32 soft boxes / ellipsoids dropped into a bin, bodies wholly inside others included), 48 scenes: every
    surface the surface intersector reports is one of the volume intersector's, every one of its faces one of that
    surface's faces with the same area bit for bit -- nothing is invented. It is NOT always complete here: the walk stops
    at a pair that produces no polygon (gradient filters, a sliver that dedupes below three vertices), so polygons
    behind such a pair are lost. That is a property of the reference's overload, not of this build: its faithful
    restatement loses faces on the same configuration (tests/test_oracle_surface_bfs.py:
    test_the_surface_overload_loses_polygons_in_dense_clutter, 0.3 % of the soft-soft area)."""
    from drake_b200 import workloads as W
    wl = W.config4(48)
    results = []
    for c in (ctx, volume_ctx):
        batch = wl.register(c)
        results.append(batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON))
    by_surface, by_volume = results
    total = complete = soft_soft = 0
    area_full = area_part = 0.0
    for sc in range(wl.num_scenes):
        full = {(s.id_M, s.id_N): s for s in by_volume[sc]}
        for s in by_surface[sc]:
            f = full[(s.id_M, s.id_N)]            # KeyError = a surface the volume intersector does not have
            fk, pk = parity.face_keys(f.elem_M, f.elem_N), parity.face_keys(s.elem_M, s.elem_N)
            assert len(np.unique(pk)) == len(pk) and np.all(np.isin(pk, fk))
            fo, po = np.argsort(fk, kind="stable"), np.argsort(pk, kind="stable")
            sel = fo[np.searchsorted(fk[fo], pk[po])]
            np.testing.assert_array_equal(s.areas[po], f.areas[sel])
            complete += len(pk) == len(fk)
            if s.grad_eM is not None and s.grad_eN is not None:
                soft_soft += 1
                area_part += float(np.sort(s.areas).sum())
            else:
                assert len(pk) == len(fk)          # soft-rigid pairs do not go through the surface intersector
        for f in by_volume[sc]:
            if f.grad_eM is not None and f.grad_eN is not None:
                area_full += float(np.sort(f.areas).sum())
        total += len(by_volume[sc])
    reported = sum(len(x) for x in by_surface)
    print(f"dense clutter: {reported} of {total} surfaces reported, {complete} complete, {soft_soft} soft-soft, "
          f"soft-soft area {area_part / area_full:.5f} of the volume intersector's")
    assert total > 1500 and soft_soft > 100
    assert reported <= total and reported >= 0.98 * total and complete >= 0.6 * reported
    assert 0.99 * area_full <= area_part <= area_full
