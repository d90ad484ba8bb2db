"""GPU-vs-oracle parity of the half-space special cases through the C-ABI: compliant
mesh x rigid half space (k_plane over tetrahedra, mesh_plane_intersection.cc:182-342) and compliant half
space x rigid mesh (k_plane over triangles, mesh_half_space_intersection.cc:17-430). Run on the B200 box
with -m gpu; the file sorts after test_gpu_parity.py on purpose (newest kernels last)."""
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import ho
from tests import parity
from tests.test_oracle_halfspace import box_mesh, trivial_volume_mesh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    import drake_b200 as db
    return db.Context(0)


def is_welded(surface):
    return bool(np.any(np.asarray(surface.elem_first) < 0) or np.any(np.asarray(surface.elem_second) < 0))


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("ids", [(1, 2), (2, 1)])
def test_compliant_volume_rigid_half_space_kat(ctx, rep, ids):
    """mesh_plane_intersection_test.cc:1217-1316: six triangles / two polygons at x = 0.5 in S, normals along
    Sx (reversed with the ids), null when moved apart."""
    import drake_b200 as db
    mesh = trivial_volume_mesh()
    fld = ho.VolumeField.from_values(mesh, [0.25, 0.5, 0.75, 1, -1])
    X_WS = ho.xform(ho.angle_axis_matrix(math.pi / 3, np.array([-1, 1, 1]) / math.sqrt(3)), [1.25, 2.5, -3.75])
    X_SR = ho.xform(ho.angle_axis_matrix(math.pi / 2, [0, 1, 0]), [0.5, 2.25, -1.25])
    X_WR = ho.compose(X_WS, X_SR)
    a_s, a_h = parity.register_compliant(ctx, mesh, fld), ctx.register_half_space()
    assert ctx.asset_info(a_h).is_half_space == 1
    batch = db.SceneBatch(ctx, [0, 2], [a_s, a_h], list(ids))
    assert batch.info().num_half_space_pairs == 1
    got = batch.compute_contact_surfaces(np.stack([X_WS, X_WR]), rep)[0]
    assert len(got) == 1
    assert got[0].num_faces == (6 if rep == ho.KTRIANGLE else 2)
    want = ho.soft_mesh_rigid_half_space(mesh, fld, X_WS, X_WR, ids[0], ids[1], rep)
    parity.compare_surface(got[0], want, welded=True)
    R_WS, _ = ho.xform_parts(X_WS)
    sign = 1.0 if ids == (1, 2) else -1.0
    np.testing.assert_allclose(got[0].normals, np.tile(sign * R_WS[:, 0], (got[0].num_faces, 1)), atol=1e-14)
    assert set(got[0].elem_N) == {-1} and sorted(set(got[0].elem_M)) == [0, 1]
    # Moved out of contact (cases 2 and 3 of the reference test).
    X_WS2 = ho.compose(X_WS, ho.xform(None, [-0.51, 0, 0]))
    assert batch.compute_contact_surfaces(np.stack([X_WS2, X_WR]), rep)[0] == []
    R_SR, _ = ho.xform_parts(X_SR)
    X_WR2 = ho.compose(X_WS, ho.xform(R_SR, [1.01, 0, 0]))
    assert batch.compute_contact_surfaces(np.stack([X_WS, X_WR2]), rep)[0] == []


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("ids", [(1, 2), (2, 1)])
def test_compliant_half_space_rigid_box_kat(ctx, rep, ids):
    """mesh_half_space_intersection_test.cc:886-1007: 10 polygons (34 triangles) when the boundary passes
    through the box centre; pressure = 17 x depth; total area 12. ids = (half space, mesh)."""
    import drake_b200 as db
    X_WF = ho.xform(ho.angle_axis_matrix(math.pi / 7, np.array([1, 2, 3]) / math.sqrt(14)), [-0.25, 0.5, 0.75])
    X_FB = ho.xform(ho.rpy_matrix(math.pi / 4, math.pi / 4, math.pi / 4), [1.0, 2.0, 3.0])
    box = box_mesh(X_FB)
    X_WH = ho.compose(X_WF, X_FB)
    a_h, a_r = ctx.register_half_space(compliant=True, pressure_scale=17.0), parity.register_rigid(ctx, box)
    batch = db.SceneBatch(ctx, [0, 2], [a_h, a_r], list(ids))
    got = batch.compute_contact_surfaces(np.stack([X_WH, X_WF]), rep)[0]
    assert len(got) == 1
    assert got[0].num_faces == (34 if rep == ho.KTRIANGLE else 10)
    want = ho.soft_half_space_rigid_mesh(box, X_WH, 17.0, X_WF, ids[0], ids[1], rep)
    parity.compare_surface(got[0], want, welded=True)
    assert abs(got[0].total_area - 12.0) < 1e-12
    assert abs(got[0].pressures.max() - 17.0) < 1e-12
    # The half space's gradient is grad_eM when it is M (smaller id), grad_eN otherwise.
    R_WH, _ = ho.xform_parts(X_WH)
    grad = got[0].grad_eM if ids == (1, 2) else got[0].grad_eN
    np.testing.assert_allclose(grad, np.tile(-17.0 * R_WH[:, 2], (got[0].num_faces, 1)), atol=1e-13)
    # Well below the box: null.
    assert batch.compute_contact_surfaces(np.stack([ho.xform(None, [0, 0, -50.0]), X_WF]), rep)[0] == []


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_welded_half_space_surfaces_have_the_reference_vertex_counts(ctx, rep):
    """The reference shares the vertices of a half-space surface between its polygons (cut_edges, mesh_plane_
    intersection.cc:205-231; mesh_half_space_intersection.cc:52-101); the kernels give every polygon its own and say where
    each comes from (vertex_origin). Welded on the host (ContactSurface.weld) the surfaces have the reference's vertex
    counts -- 12 vertices / 10 polygons and 22 / 34 for the box of mesh_half_space_intersection_test.cc:886-1007 -- and
    compare with the oracle's welded surfaces as any mesh-mesh surface does (vertex count equal, not just >=). Second
    case: a 4096-tetrahedron compliant sphere sliced by a rigid half space."""
    import drake_b200 as db
    X_WF = ho.xform(ho.angle_axis_matrix(math.pi / 7, np.array([1, 2, 3]) / math.sqrt(14)), [-0.25, 0.5, 0.75])
    X_FB = ho.xform(ho.rpy_matrix(math.pi / 4, math.pi / 4, math.pi / 4), [1.0, 2.0, 3.0])
    box = box_mesh(X_FB)
    X_WH = ho.compose(X_WF, X_FB)
    a_h, a_r = ctx.register_half_space(compliant=True, pressure_scale=17.0), parity.register_rigid(ctx, box)
    batch = db.SceneBatch(ctx, [0, 2], [a_h, a_r], [2, 1])
    batch.set_poses(np.stack([X_WH, X_WF]))
    batch.run(rep, db.STRICT)
    raw = batch.copy_results(vertex_origin=True).surfaces()[0]
    assert len(raw.vertices) == (44 if rep == ho.KTRIANGLE else 34)   # one set of vertices per polygon
    welded = raw.weld(box.arrays()[1])
    assert (len(welded.vertices), welded.num_faces) == ((22, 34) if rep == ho.KTRIANGLE else (12, 10))
    parity.compare_surface(welded, ho.soft_half_space_rigid_mesh(box, X_WH, 17.0, X_WF, 2, 1, rep))
    soft = ho.VolumeMesh.sphere(1.0, 0.25, dense=True)
    fld = ho.VolumeField.sphere(soft, 1.0, 1e5)
    a_s, a_g = parity.register_compliant(ctx, soft, fld), ctx.register_half_space()
    batch = db.SceneBatch(ctx, [0, 2], [a_s, a_g], [1, 2])
    X_WS = ho.xform(ho.rpy_matrix(0.3, -0.4, 0.2), [0.1, -0.2, 0.55])
    X_WG = ho.xform(ho.rpy_matrix(0.05, 0.1, 0.0), [0, 0, 0])
    batch.set_poses(np.stack([X_WS, X_WG]))
    batch.run(rep, db.STRICT)
    raw = batch.copy_results(vertex_origin=True).surfaces()[0]
    want = ho.soft_mesh_rigid_half_space(soft, fld, X_WS, X_WG, 1, 2, rep)
    welded = raw.weld(soft.tets)
    assert len(welded.vertices) == len(want.vertices) < len(raw.vertices)
    assert welded.num_faces == want.num_faces > 100
    parity.compare_surface(welded, want)


def test_compliant_half_space_margin(ctx):
    """mesh_half_space_intersection_test.cc:1107-1186: the box only reaches into the margin layer."""
    import drake_b200 as db
    X_WF = ho.xform(ho.angle_axis_matrix(math.pi / 7, np.array([1, 2, 3]) / math.sqrt(14)), [-0.25, 0.5, 0.75])
    R_WF, p_WF = ho.xform_parts(X_WF)
    box = box_mesh(ho.xform(None, [0, 0, 1.0]))
    scale, margin = 1.5, 0.012
    X_WH = ho.xform(R_WF, p_WF + (-margin / 2.0) * R_WF[:, 2])
    a_h = ctx.register_half_space(compliant=True, pressure_scale=scale, margin=margin)
    batch = db.SceneBatch(ctx, [0, 2], [a_h, parity.register_rigid(ctx, box)], [1, 2])
    for rep in (ho.KTRIANGLE, ho.KPOLYGON):
        got = batch.compute_contact_surfaces(np.stack([X_WH, X_WF]), rep)[0]
        assert len(got) == 1
        eps = np.finfo(float).eps
        assert abs(got[0].pressures.min() - (-margin * scale)) <= 4 * eps
        assert abs(got[0].pressures.max() - (-margin * scale / 2.0)) <= 4 * eps
        parity.compare_surface(got[0], ho.soft_half_space_rigid_mesh(box, X_WH, scale, X_WF, 1, 2, rep, margin=margin),
                               welded=True)


def mixed_scene(ctx):
    """A compliant ball, a rigid ball, a rigid half space (floor) and a compliant half space (pad), plus a
    rigid box: mesh pairs, both half-space kinds and the rejected combinations (filtered) in one scene."""
    ball = ho.VolumeMesh.sphere(0.05, 0.015)
    fld = ho.VolumeField.sphere(ball, 0.05, 3e4)
    rigid_ball = ho.VolumeMesh.sphere(0.04, 0.02).to_surface()
    box = ho.TriMesh.box(0.2, 0.2, 0.01)
    assets = [parity.register_compliant(ctx, ball, fld), parity.register_rigid(ctx, rigid_ball),
              ctx.register_half_space(), ctx.register_half_space(compliant=True, pressure_scale=3e4 / 0.02, margin=1e-4),
              parity.register_rigid(ctx, box)]
    ids = [10, 7, 3, 12, 5]  # half spaces on both sides of the mesh ids: both swap states occur
    osc = ho.Scene()
    osc.add_compliant(10, ball, fld)
    osc.add_rigid(7, rigid_ball)
    osc.add_half_space(3, anchored=True)
    osc.add_half_space(12, compliant=True, pressure_scale=3e4 / 0.02, margin=1e-4, anchored=True)
    osc.add_rigid(5, box)
    # Rigid-rigid and the rejected half-space combinations are filtered out, as a plant would do.
    excluded_ids = [(7, 3), (7, 5), (3, 5), (10, 12), (3, 12)]
    for a, b in excluded_ids:
        osc.exclude(a, b)
    return assets, ids, osc, excluded_ids


def mixed_poses(rng, num_scenes):
    P = np.zeros((num_scenes, 5, 12))
    for s in range(num_scenes):
        tilt = ho.rpy_matrix(*(0.2 * rng.uniform(-1, 1, size=3)))
        P[s, 0] = parity.random_pose(rng, 0.0, center=(rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05),
                                                        rng.uniform(0.0, 0.06)))        # soft ball near the floor
        P[s, 1] = parity.random_pose(rng, 0.0, center=(0.3 + rng.uniform(-0.02, 0.02), 0.0, rng.uniform(0.0, 0.05)))
        P[s, 2] = ho.xform(tilt, [0, 0, 0])                                               # rigid floor, tilted
        P[s, 3] = ho.xform(tilt.T, [0.3, 0, 0.01])                                        # compliant pad under ball 7
        P[s, 4] = ho.xform(None, [rng.uniform(-0.03, 0.03), 0, 0.08 + rng.uniform(0, 0.02)])  # plate above ball 10
    return P


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_mixed_scenes_with_half_spaces(ctx, rep):
    """48 scenes mixing mesh pairs with both half-space kinds against the scene-level oracle: ordering by
    (id_M, id_N), M/N swaps, per-pair offsets through the shared scan, chunked = unchunked."""
    import drake_b200 as db
    assets, ids, osc, excluded_ids = mixed_scene(ctx)
    S = 48
    G = len(ids)
    flat = lambda s, gid: s * G + ids.index(gid)
    excluded = [(flat(s, a), flat(s, b)) for s in range(S) for a, b in excluded_ids]
    anchored = [0, 0, 1, 1, 0] * S
    batch = db.SceneBatch(ctx, [G * s for s in range(S + 1)], assets * S, ids * S, anchored, excluded)
    info = batch.info()
    assert info.num_half_space_pairs == 3 * S   # (10,3), (7,12), (5,12)
    P = mixed_poses(np.random.default_rng(5), S)
    got = batch.compute_contact_surfaces(P.reshape(-1, 12), rep)
    order = np.argsort(ids)
    counts = {}
    for s in range(S):
        want, _ = osc.compute(P[s][order], rep)
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want], s
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w, welded=is_welded(w))
            counts[(w.id_M, w.id_N)] = counts.get((w.id_M, w.id_N), 0) + 1
    # Every kind of pair occurs: soft ball on the floor, rigid ball in the pad, soft ball under the plate.
    assert counts.get((3, 10), 0) >= 8 and counts.get((7, 12), 0) >= 8 and counts.get((5, 10), 0) >= 4, counts
    # Chunking does not change anything.
    batch.set_chunks(5)
    again = batch.compute_contact_surfaces(P.reshape(-1, 12), rep)
    for s in range(S):
        assert len(again[s]) == len(got[s])
        for x, y in zip(again[s], got[s]):
            assert (x.id_M, x.id_N, x.num_faces) == (y.id_M, y.id_N, y.num_faces)
            np.testing.assert_allclose(x.total_area, y.total_area, rtol=1e-12)


def test_fine_mesh_on_half_space(ctx):
    """A dense sphere (32768 tetrahedra) pressed into a rigid half space: thousands of slices in one pair
    (warp-aggregated counters, the big-surface reduction) against the oracle."""
    import drake_b200 as db
    ball = ho.VolumeMesh.sphere(1.0, 0.12, dense=True)
    fld = ho.VolumeField.sphere(ball, 1.0, 1e5)
    batch = db.SceneBatch(ctx, [0, 2], [parity.register_compliant(ctx, ball, fld), ctx.register_half_space()], [1, 2])
    rng = np.random.default_rng(3)
    X_WS = parity.random_pose(rng, 0.0, center=(0.1, -0.2, 0.4))
    X_WR = ho.xform(ho.rpy_matrix(0.1, -0.2, 0.3), [0, 0, 0])
    for rep, at_least in ((ho.KPOLYGON, 1500), (ho.KTRIANGLE, 4097)):  # > 4096 faces: summed by the whole grid
        got = batch.compute_contact_surfaces(np.stack([X_WS, X_WR]), rep)[0]
        want = ho.soft_mesh_rigid_half_space(ball, fld, X_WS, X_WR, 1, 2, rep)
        assert len(got) == 1 and want.num_faces >= at_least
        parity.compare_surface(got[0], want, welded=True)


def test_rejected_half_space_combinations(ctx):
    """hydroelastic_calculator.cc:113-123,145-147 + proximity_engine.cc:1361-1390: the strict query fails,
    the WithFallback query skips the pair."""
    import drake_b200 as db
    from drake_b200 import engine
    ball = ho.VolumeMesh.sphere(0.5, 0.25)
    fld = ho.VolumeField.sphere(ball, 0.5, 1e8)
    a_soft = parity.register_compliant(ctx, ball, fld)
    a_rigid = parity.register_rigid(ctx, ball.to_surface())
    hs_r, hs_c = ctx.register_half_space(), ctx.register_half_space(compliant=True, pressure_scale=1.0)
    poses = np.stack([ho.xform(None, [0, 0, 0.3]), ho.IDENTITY])
    for assets, status in [([hs_r, hs_c], engine.ERR_HALF_SPACE_HALF_SPACE),
                           ([hs_c, hs_c], engine.ERR_HALF_SPACE_HALF_SPACE),
                           ([a_soft, hs_c], engine.ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH),
                           ([a_rigid, hs_r], engine.ERR_RIGID_RIGID)]:
        batch = db.SceneBatch(ctx, [0, 2], assets, [1, 2])
        batch.set_poses(poses)
        with pytest.raises(db.HydroError) as ei:
            batch.run(ho.KPOLYGON, engine.STRICT)
        assert ei.value.status == status
        batch.run(ho.KPOLYGON, engine.WITH_FALLBACK)
        assert batch.sizes().num_surfaces == 0


def test_cpp_adapter_half_space():
    """The C++ mirror (HalfSpace shape, slab_thickness) through tests/cpp/test_half_space_engine.cc."""
    from drake_b200 import build
    lib = build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_half_space_engine")
    src = os.path.join(ROOT, "tests", "cpp", "test_half_space_engine.cc")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                    "-o", exe, "-L", os.path.dirname(lib), "-ldrake_b200_hydro", "-Wl,-rpath," + os.path.dirname(lib)],
                   check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


def test_config3_on_a_half_space_ground(ctx):
    """The bench workload c3h (ball on a rigid HalfSpace ground, plate on the ball): 32 scenes against the
    scene-level oracle, and the half-space element counter of the stats call."""
    from drake_b200 import workloads as W
    wl = W.config3_half_space(32)
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    osc, order = parity.oracle_scene_for(wl)
    total = 0
    for s in range(wl.num_scenes):
        want, _ = osc.compute(wl.poses[s][order], ho.KPOLYGON)
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want]
        for g, w in zip(got[s], want):
            parity.compare_surface(g, w, welded=is_welded(w))
        total += len(want)
    assert total == 64  # ball x ground and ball x plate in every scene
    st = batch.stats()
    # Elements sliced = the leaf blocks of the ball's tree that reach the ground plane (Bvh::Collide(plane)), not the
    # whole 512-tetrahedron mesh of every scene.
    assert 0 < st["half_space_elements"] < 32 * 512 and st["half_space_elements"] % 128 == 0
    assert st["geometry_pairs_tested"] == 64


@pytest.mark.parametrize("name", ["c3h_16_scenes", "c3h_16_scenes_tri"])
def test_half_space_golden_fixtures(ctx, name):
    """GPU results against tests/golden/c3h_*.json (no oracle call), as test_against_committed_golden_fixtures does
    for the mesh-only configs. The fixture's vertex count of a half-space surface is the welded one (see
    make_golden.py): not compared for surfaces whose faces carry -1 as an emitting element."""
    import json
    from tests.golden import make_golden as G
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(G.__file__)), name + ".json")))
    make, rep = G.CASES[name]
    wl = make()
    batch = wl.register(ctx)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), rep)
    assert len(got) == len(want["scenes"])
    for gs, ws in zip(got, want["scenes"]):
        assert [(g.id_M, g.id_N, g.num_faces) for g in gs] == [(w["id_M"], w["id_N"], w["faces"]) for w in ws]
        for g, w in zip(gs, ws):
            if np.all(np.asarray(g.elem_M) >= 0) and np.all(np.asarray(g.elem_N) >= 0):
                assert len(g.vertices) == w["vertices"]
            else:
                assert len(g.vertices) >= w["vertices"]
            keys = parity.face_keys(g.elem_M, g.elem_N) * np.int64(2654435761) % np.int64(1 << 61)
            assert int(np.bitwise_xor.reduce(keys)) == w["key_checksum"]
            assert g.total_area == pytest.approx(w["total_area"], rel=1e-10)
            assert float(np.sum(g.pressure_at_centroid)) == pytest.approx(w["sum_face_pressure"], rel=1e-10)
            assert np.allclose(g.centroid, w["centroid"], rtol=1e-10, atol=1e-13)
