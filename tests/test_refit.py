"""Deformable reuse, first step: triangle-mesh assets whose vertices move from step to step --
DeformableMeshWithBvh::UpdateVertexPositions = new positions + derived quantities + BvhUpdater::Update (topology kept,
boxes refitted bottom-up; bvh_updater.h:45-103). The oracle's restatement is checked against the updater's contract
(bvh_updater_test.cc: every box bounds its contents, the root is the mesh's box; topology unchanged); the device path
(dbh_context_update_vertices: k_refit_tris / k_refit_nodes / k_refit_blocks) must give, after an update, the surfaces
the oracle computes on the deformed mesh."""
import copy
import math

import numpy as np
import pytest

from oracle import ho


def deform(verts, t):
    """A smooth, clearly non-rigid deformation of a point set (stretch + twist + ripple), parameter t."""
    v = np.asarray(verts, float).copy()
    ang = 0.8 * t * v[:, 2] / max(np.abs(v[:, 2]).max(), 1e-9)
    c, s = np.cos(ang), np.sin(ang)
    x, y = v[:, 0] * c - v[:, 1] * s, v[:, 0] * s + v[:, 1] * c
    v[:, 0], v[:, 1] = x * (1 + 0.3 * t), y * (1 - 0.2 * t)
    v[:, 2] = v[:, 2] * (1 + 0.15 * t) + 0.02 * t * np.sin(7 * x)
    return v


def test_oracle_refit_keeps_topology_and_bounds_the_deformed_mesh():
    m = ho.TriMesh.sphere(0.25, 0.05)
    nodes0, leaves0, height0 = m.bvh_stats()
    assert m.bvh_is_fitted()
    v0 = m.verts.copy()
    for t in (0.5, 1.0, -0.7):
        m.update_vertices(deform(v0, t))
        assert m.bvh_stats() == (nodes0, leaves0, height0)      # no change of topology
        assert m.bvh_is_fitted()                                # every box bounds its contents, root = mesh box
        np.testing.assert_array_equal(m.verts, deform(v0, t))
    # Normals follow the vertices: recomputed, unit length.
    n = m.arrays()[2]
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-14)
    # The traversal over the refitted tree still finds every overlapping element pair of the deformed mesh.
    soft = ho.VolumeMesh.sphere(0.3, 0.1)
    X = ho.xform(ho.rpy_matrix(0.2, 0.1, 0.4), [0.1, 0.05, 0.3])
    brute = set(map(tuple, ho.candidates(soft, m, X, ho.MODE_BRUTE)[0]))
    tree = set(map(tuple, ho.candidates(soft, m, X, ho.MODE_REFERENCE_BVH)[0]))
    assert brute and brute <= tree


@pytest.mark.gpu
@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_device_refit_gives_the_surfaces_of_the_deformed_mesh(rep):
    """Two deforming surfaces (one asset each) in contact with a static pressure-field volume (the rigid side of a
    deformable-rigid pair), updated three times: after every update the
    query equals the oracle on the deformed mesh face for face, and a freshly registered copy of the deformed mesh gives
    the same surface as the refitted one."""
    import drake_b200 as db
    from tests import parity
    ctx = db.Context(0)
    vol = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(vol, 0.3, 1e6)
    surf = [ho.TriMesh.sphere(0.25, 0.05), ho.TriMesh.ellipsoid(0.2, 0.3, 0.25, 0.06)]
    v0 = [s.verts.copy() for s in surf]
    a_v = parity.register_compliant(ctx, vol, fld)
    a_s = [parity.register_rigid(ctx, s) for s in surf]
    poses = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(ho.rpy_matrix(0.2, 0.1, 0.4), [0.1, 0.05, 0.3]),
                      ho.xform(ho.rpy_matrix(-0.3, 0.2, 0.1), [-0.15, 0.1, -0.32])])
    batch = db.SceneBatch(ctx, [0, 3], [a_v, a_s[0], a_s[1]], [1, 2, 3], excluded=[[1, 2]])
    batch.set_record_candidates(True)
    for step, t in enumerate((0.0, 0.6, 1.0, -0.8)):
        if step:
            new = [deform(v, t) for v in v0]
            ctx.update_vertices(a_s, new)
            for s, v in zip(surf, new):
                s.update_vertices(v)
        got = batch.compute_contact_surfaces(poses, rep)[0]
        assert [(g.id_M, g.id_N) for g in got] == [(1, 2), (1, 3)]
        for k, g in enumerate(got):
            want = ho.soft_rigid(vol, fld, surf[k], poses[0], poses[1 + k], 1, 2 + k, rep)
            parity.compare_surface(g, want)
        if step:
            fresh = [parity.register_rigid(ctx, ho.TriMesh.from_arrays(s.verts, s.arrays()[1])) for s in surf]
            b2 = db.SceneBatch(ctx, [0, 3], [a_v, fresh[0], fresh[1]], [1, 2, 3], excluded=[[1, 2]])
            again = b2.compute_contact_surfaces(poses, rep)[0]
            for g, h in zip(got, again):
                np.testing.assert_array_equal(np.sort(parity.face_keys(g.elem_M, g.elem_N)),
                                              np.sort(parity.face_keys(h.elem_M, h.elem_N)))
                assert g.total_area == pytest.approx(h.total_area, rel=1e-12)


@pytest.mark.gpu
def test_device_refit_of_many_assets_and_a_half_space_pair():
    """64 scenes, each with its own deforming surface asset (registered one after the other: one upload and three
    launches re-pose them all) over a compliant half space and against a static pressure-field volume; per-scene
    deformations; every surface against the oracle after the update (the half-space pair exercises the refitted
    leaf-block boxes)."""
    import drake_b200 as db
    from tests import parity
    ctx = db.Context(0)
    vol = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(vol, 0.3, 1e6)
    base = ho.TriMesh.sphere(0.25, 0.06)
    v0, tris = base.verts.copy(), base.arrays()[1]
    S = 64
    a_v = parity.register_compliant(ctx, vol, fld)
    a_h = ctx.register_half_space(compliant=True, pressure_scale=2e5)
    a_s = [parity.register_rigid(ctx, base) for _ in range(S)]
    so = np.arange(S + 1) * 3
    ga = np.array([[a_v, a_s[s], a_h] for s in range(S)]).reshape(-1)
    gi = np.tile([1, 2, 3], S)
    an = np.tile([0, 0, 1], S)
    ex = np.array([[3 * s, 3 * s + 2] for s in range(S)])     # volume x half space: rejected combination, filtered
    batch = db.SceneBatch(ctx, so, ga, gi, an, ex)
    rng = np.random.default_rng(3)
    poses = np.stack([np.stack([ho.xform(None, [0, 0, 0.4]), parity.random_pose(rng, 0.05, center=(0.1, 0.0, 0.15)),
                                ho.xform(None, [0, 0, 0])]) for _ in range(S)])
    ts = rng.uniform(-1, 1, S)
    new = [deform(v0, t) for t in ts]
    ctx.update_vertices(a_s, new)
    got = batch.compute_contact_surfaces(poses.reshape(-1, 12), ho.KPOLYGON)
    total = 0
    for s in range(S):
        mesh = ho.TriMesh.from_arrays(new[s], tris)
        osc = ho.Scene()
        osc.add_compliant(1, vol, fld)
        osc.add_rigid(2, mesh)
        osc.add_half_space(3, compliant=True, pressure_scale=2e5, anchored=True)
        osc.exclude(1, 3)
        want, _ = osc.compute(poses[s], ho.KPOLYGON)
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want], s
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w, welded=bool(np.any(np.asarray(w.elem_second) < 0) or np.any(np.asarray(w.elem_first) < 0)))
        total += len(want)
    assert total >= S


@pytest.mark.gpu
@pytest.mark.parametrize("ids", [(1, 2, 3), (7, 2, 3)])
def test_deformable_rigid_contact_data_after_device_refit(ids):
    """SURVEY 8f rank 4, deformable x rigid: k_deformable over the faces of two deforming surfaces against a rigid
    geometry's pressure-field volume, after every device-side vertex update, against oracle/ho_deformable.py
    (AddDeformableRigidContactSurface + the DeformableContactSurface constructor): contact points, normals into the
    deformable body, R_WC, pressures, the reference's point-transformed gradients, triangle indices and barycentric
    centroids, faces matched by their emitting element pair. ids (7, 2, 3): the rigid geometry has the larger id, so the
    ContactSurface is stored swapped and the consumer has to undo it."""
    import drake_b200 as db
    from drake_b200 import engine
    from oracle import ho_deformable as hd
    from tests import parity
    from tests.test_plane_math_cpu import check_deformable_rows
    ctx = db.Context(0)
    vol = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(vol, 0.3, 1e6)
    surf = [ho.TriMesh.sphere(0.25, 0.05), ho.TriMesh.ellipsoid(0.2, 0.3, 0.25, 0.06)]
    v0 = [s.verts.copy() for s in surf]
    a_v = parity.register_compliant(ctx, vol, fld)
    a_s = [parity.register_rigid(ctx, s) for s in surf]
    # The rigid geometry is posed; the deformable surfaces live in World (identity), as in the reference.
    X_WR = ho.xform(ho.rpy_matrix(0.2, 0.1, 0.4), [0.1, 0.05, 0.3])
    offsets = [np.array([0.0, 0.0, 0.0]), np.array([0.3, -0.1, 0.75])]
    poses = np.stack([X_WR, ho.IDENTITY, ho.IDENTITY])
    batch = db.SceneBatch(ctx, [0, 3], [a_v, a_s[0], a_s[1]], list(ids), excluded=[[1, 2]])
    checked = 0
    for t in (0.0, 0.6, -0.8):
        new = [deform(v, t) + o for v, o in zip(v0, offsets)]
        ctx.update_vertices(a_s, new)
        for s, v in zip(surf, new):
            s.update_vertices(v)
        batch.set_poses(poses)
        batch.query(ho.KPOLYGON, engine.STRICT)
        batch.compute_deformable_contact()
        batch.sync()
        rows, _, tri = batch.copy_deformable_contact()
        res = batch.copy_results()
        got = res.surfaces()
        want = [hd.deformable_rigid_contact(vol, fld, s, X_WR) for s in surf]
        assert len(got) == sum(w is not None for w in want) >= 1
        k = 0
        for which, w in enumerate(want):
            if w is None:
                continue
            g = got[k]
            assert {g.id_M, g.id_N} == {ids[0], ids[1 + which]}
            f0 = int(res.surf_face_offset[k])
            mine, mine_tri = rows[f0:f0 + g.num_faces], tri[f0:f0 + g.num_faces]
            tet, tr = g.elem_M, g.elem_N      # elements of the first-constructed body (the volume) / the triangle mesh
            np.testing.assert_array_equal(mine_tri, tr)
            go = np.argsort(parity.face_keys(tet, tr), kind="stable")
            wo = np.argsort(parity.face_keys(w.surface.elem_first, w.surface.elem_second), kind="stable")
            w2 = copy.copy(w)
            for name in ("contact_points_W", "nhats_W", "R_WCs", "pressures", "pressure_gradients_W", "triangle_index",
                         "barycentric_centroids"):
                setattr(w2, name, getattr(w, name)[wo])
            check_deformable_rows(mine[go], mine_tri[go], w2, atol_scale=10.0)
            checked += g.num_faces
            k += 1
    assert checked > 300
    with pytest.raises(engine.HydroError):
        batch.query(ho.KTRIANGLE, engine.STRICT)
        batch.compute_deformable_contact()


def _sdf_spheres():
    a = ho.VolumeMesh.sphere(0.25, 0.08)
    b = ho.VolumeMesh.sphere(0.2, 0.07)
    sa = np.linalg.norm(a.verts, axis=1) - 0.25      # signed distances at rest: negative inside, 0 on the surface
    sb = np.linalg.norm(b.verts, axis=1) - 0.2
    return a, sa, b, sb, np.array([0.1, 0.05, 0.36])


@pytest.mark.gpu
def test_volume_mesh_updated_to_its_own_vertices_keeps_every_bit():
    """hydro_derive.cuh is one function for the host (registration) and the device (k_refit_tets): a volume mesh
    re-posed on the device to the positions it was registered with gives the packed result it gave before, bit for bit
    (per-tetrahedron records, cull copies, leaf boxes and tree boxes all derived again on the device)."""
    import drake_b200 as db
    from drake_b200 import engine
    ctx = db.Context(0)
    a, sa, b, sb, off = _sdf_spheres()
    soft = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(soft, 0.3, 1e6)
    h = [ctx.register_compliant_mesh(a.verts, a.tets, sa), ctx.register_compliant_mesh(b.verts + off, b.tets, sb),
         ctx.register_compliant_mesh(soft.verts, soft.tets, fld.values)]
    poses = np.stack([ho.IDENTITY, ho.IDENTITY, ho.xform(ho.rpy_matrix(0.3, -0.2, 0.1), [-0.2, 0.1, -0.25])])
    batch = db.SceneBatch(ctx, [0, 3], h, [1, 2, 3])

    def run():
        batch.set_poses(poses)
        batch.query(ho.KPOLYGON, engine.STRICT)
        batch.sync()
        res = batch.copy_results()
        surf_of_face = np.repeat(np.arange(res.num_surfaces), np.diff(res.surf_face_offset))
        order = np.lexsort((res.elem_N, res.elem_M, surf_of_face))
        return res, order

    r0, o0 = run()
    nf = len(r0.areas)
    assert nf > 100 and r0.num_surfaces >= 2
    ctx.update_vertices(h, [a.verts, b.verts + off, soft.verts])
    r1, o1 = run()
    assert len(r1.areas) == nf
    for name in ("areas", "normals", "centroids", "pressure_at_centroid", "grad_eM", "grad_eN", "elem_M", "elem_N"):
        x0, x1 = getattr(r0, name), getattr(r1, name)
        np.testing.assert_array_equal(x0.reshape(nf, -1)[o0], x1.reshape(nf, -1)[o1], err_msg=name)
    # (Totals are sums in face order, which is the order of the atomics: equal to rounding only.)
    np.testing.assert_allclose(r0.surf_total_area, r1.surf_total_area, rtol=1e-13)


@pytest.mark.gpu
def test_deformable_deformable_contact_data_after_device_refit():
    """SURVEY 8f rank 4, deformable x deformable: two deforming volume meshes carrying signed-distance fields, updated on
    the device every step (k_refit_tets: gradients, planes, cull copies, boxes), then the tet-tet pipeline and k_deformable
    against oracle/ho_deformable.py (AddDeformableDeformableContactSurface): the contact mesh face for face, contact
    points, normals, R_WC, signed distances and the barycentric coordinates in both tetrahedra."""
    import drake_b200 as db
    from drake_b200 import engine
    from oracle import ho_deformable as hd
    from tests import parity
    from tests.test_plane_math_cpu import check_deformable_pair_rows
    ctx = db.Context(0)
    a, sa, b, sb, off = _sdf_spheres()
    va0, vb0 = a.verts.copy(), b.verts.copy()
    ha = ctx.register_compliant_mesh(va0, a.tets, sa)
    hb = ctx.register_compliant_mesh(vb0 + off, b.tets, sb)
    # M = the first mesh: the adapter gives the reference's `deformable0` (its larger id) the smaller batch id.
    batch = db.SceneBatch(ctx, [0, 2], [ha, hb], [1, 2])
    poses = np.stack([ho.IDENTITY, ho.IDENTITY])
    checked = 0
    for t in (0.0, 0.5, -0.6):
        va, vb = deform(va0, t), deform(vb0, -t) + off
        ctx.update_vertices([ha, hb], [va, vb])
        ma, mb = ho.VolumeMesh.from_arrays(va, a.tets), ho.VolumeMesh.from_arrays(vb, b.tets)
        want = hd.deformable_deformable_contact(ma, sa, mb, sb)
        assert want is not None
        batch.set_poses(poses)
        batch.query(ho.KPOLYGON, engine.STRICT)
        batch.compute_deformable_contact()
        batch.sync()
        rows, ea, eb = batch.copy_deformable_contact()
        got = batch.copy_results().surfaces()
        assert len(got) == 1 and (got[0].id_M, got[0].id_N) == (1, 2)
        parity.compare_surface(got[0], want.surface)
        go = np.argsort(parity.face_keys(ea, eb), kind="stable")
        wo = np.argsort(parity.face_keys(want.tet0, want.tet1), kind="stable")
        np.testing.assert_array_equal(np.stack([ea, eb], 1)[go], np.stack([want.tet0, want.tet1], 1)[wo])
        w2 = copy.copy(want)
        for name in ("contact_points_W", "nhats_W", "R_WCs", "signed_distances", "tet0", "tet1", "barycentric0", "barycentric1"):
            setattr(w2, name, getattr(want, name)[wo])
        check_deformable_pair_rows(rows[go], w2, atol_scale=10.0)
        checked += len(rows)
    assert checked > 100
    # A volume mesh squashed flat has no field gradient: the update fails like the reference's field constructor.
    flat = va0.copy()
    flat[:, 2] = 0.0
    with pytest.raises(engine.HydroError):
        ctx.update_vertices([ha], [flat])


@pytest.mark.gpu
def test_update_from_device_memory_and_pool_regrowth():
    """dbh_context_update_vertices_device: the new positions come from device memory (torch tensors here); the query
    equals the one after the host-pointer update. Then a registration that regrows the pools: the re-posed meshes are
    re-uploaded from the refreshed host copy and derived again -- same surfaces as before the regrowth."""
    import torch
    import drake_b200 as db
    from tests import parity
    ctx = db.Context(0)
    vol = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(vol, 0.3, 1e6)
    surf = ho.TriMesh.sphere(0.25, 0.05)
    soft2 = ho.VolumeMesh.sphere(0.25, 0.08)
    sdf2 = np.linalg.norm(soft2.verts, axis=1) - 0.25
    a_v = parity.register_compliant(ctx, vol, fld)
    a_s = parity.register_rigid(ctx, surf)
    a_d = ctx.register_compliant_mesh(soft2.verts, soft2.tets, 0.25 - np.linalg.norm(soft2.verts, axis=1))
    poses = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(ho.rpy_matrix(0.2, 0.1, 0.4), [0.1, 0.05, 0.3]),
                      ho.xform(None, [-0.2, 0.1, -0.3])])
    batch = db.SceneBatch(ctx, [0, 3], [a_v, a_s, a_d], [1, 2, 3])
    new_s, new_d = deform(surf.verts, 0.7), deform(soft2.verts, -0.5)
    ctx.update_vertices([a_s, a_d], [new_s, new_d])
    want = batch.compute_contact_surfaces(poses, ho.KPOLYGON)[0]
    assert len(want) >= 2
    # Back to the registered shapes, then the same deformation from device memory.
    ctx.update_vertices([a_s, a_d], [surf.verts, soft2.verts])
    t = torch.from_numpy(np.concatenate([new_s.reshape(-1), new_d.reshape(-1)])).cuda()
    torch.cuda.synchronize()
    ctx.update_vertices_device([a_s, a_d], t.data_ptr())
    got = batch.compute_contact_surfaces(poses, ho.KPOLYGON)[0]

    def same(xs, ys):
        assert [(x.id_M, x.id_N) for x in xs] == [(y.id_M, y.id_N) for y in ys]
        for x, y in zip(xs, ys):
            np.testing.assert_array_equal(np.sort(parity.face_keys(x.elem_M, x.elem_N)), np.sort(parity.face_keys(y.elem_M, y.elem_N)))
            ox, oy = np.argsort(parity.face_keys(x.elem_M, x.elem_N), kind="stable"), np.argsort(parity.face_keys(y.elem_M, y.elem_N), kind="stable")
            np.testing.assert_array_equal(x.areas[ox], y.areas[oy])
            np.testing.assert_array_equal(x.centroids[ox], y.centroids[oy])

    same(got, want)
    # A large registration regrows every pool; the deformed assets must survive it.
    big = ho.VolumeMesh.sphere(0.3, 0.03, dense=True)
    ctx.register_compliant_mesh(big.verts, big.tets, ho.VolumeField.sphere(big, 0.3, 1e6).values)
    b2 = db.SceneBatch(ctx, [0, 3], [a_v, a_s, a_d], [1, 2, 3])
    same(b2.compute_contact_surfaces(poses, ho.KPOLYGON)[0], want)
