"""GPU parity of k_wrench (one spatial force per contact surface) against the oracle's
restatement of HydroelasticTractionCalculator.

All four cases passed on the B200 with the last GPU seconds of round 1 (profiles/r1h_gpu_tests_wrench.log); the kernel
has not been timed yet. The per-face arithmetic is also checked on the host against the oracle
(tests/test_plane_math_cpu.py), and the oracle against the reference's own tests (tests/test_oracle_traction.py).
"""
import os

import numpy as np
import pytest

from oracle import ho
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import drake_b200 as db
    return db.Context(0)


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_spatial_forces_of_config3_scenes(ctx, rep):
    """24 ball / plate / floor scenes (soft-rigid + soft-soft surfaces, M/N swaps): every surface's F_Ac_W, F_Ao_W,
    F_Bo_W against the oracle on the oracle's own surface, random body velocities, dissipation and friction per
    geometry combined by the reference's rules. The GPU sums faces in another order: 1e-10 of the force scale."""
    from drake_b200 import workloads as W
    wl = W.config3(24)
    batch = wl.register(ctx)
    S, G = wl.num_scenes, wl.geoms_per_scene
    rng = np.random.default_rng(4)
    V = rng.uniform(-0.5, 0.5, size=(S, G, 6))
    origins = wl.poses[:, :, 9:12] + rng.uniform(-0.05, 0.05, size=(S, G, 3))
    E = np.tile([3e4, np.inf, 3e4], S)      # floor, plate (rigid), ball
    dis = np.tile([1.0, 0.3, 2.0], S)
    mu = np.tile([0.5, 0.2, 1.0], S)
    batch.set_contact_properties(E, dis, mu)
    batch.set_body_kinematics(origins.reshape(-1, 3), V.reshape(-1, 6))
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), rep)
    F_Ac, F_Ao, F_Bo = batch.spatial_forces(1e-4)
    osc, order = parity.oracle_scene_for(wl)
    ids = list(wl.geom_id)
    k = 0
    with ho.keep_handles():
        for s in range(S):
            want, _ = osc.compute(wl.poses[s][order], rep)
            assert len(want) == len(got[s])
            for w in want:
                gM, gN = ids.index(w.id_M), ids.index(w.id_N)
                d = ho.combined_dissipation(E[gM], E[gN], dis[gM], dis[gN])
                m = ho.combined_friction(mu[gM], mu[gN])
                ref = ho.traction(w, origins[s, gM], origins[s, gN], V[s, gM], V[s, gN], d, m, 1e-4)
                scale = max(np.max(np.abs(np.concatenate(ref))), 1e-300)
                for g, r in zip((F_Ac[k], F_Ao[k], F_Bo[k]), ref):
                    assert np.max(np.abs(g - r)) <= 1e-9 * scale, (s, w.id_M, w.id_N)
                k += 1
    assert k == len(F_Ac) and k >= 40


def test_reference_fixture_through_the_c_abi(ctx):
    """FrictionalTractionOverPatch (hydroelastic_traction_calculator_test.cc:565-608) rebuilt from geometry: a
    rigid unit box pressed 0.5 into a compliant half space of pressure scale 1 gives e = 0.5 on its bottom face
    (area 1): f_Bo = (-mu e, 0, e), tau_Bo_y = 0.25 when the box slides along +x."""
    import drake_b200 as db
    from tests.test_oracle_halfspace import box_mesh
    X = ho.xform(None, [0.5, 0.5, 0.5])
    box = ho.TriMesh.from_arrays(0.5 * box_mesh(ho.IDENTITY).verts, box_mesh(ho.IDENTITY).tris)
    a_h = ctx.register_half_space(compliant=True, pressure_scale=1.0)
    batch = db.SceneBatch(ctx, [0, 2], [a_h, parity.register_rigid(ctx, box)], [1, 2])
    batch.set_contact_properties([1.0, np.inf], [0.0, 0.0], [1.0, 1.0])
    batch.set_body_kinematics(None, [[0, 0, 0, 0, 0, 0], [0, 0, 0, 1.0, 0, 0]])
    surfaces = batch.compute_contact_surfaces(np.stack([ho.IDENTITY, ho.IDENTITY]), ho.KPOLYGON)[0]
    assert len(surfaces) == 1
    _, F_Ao, F_Bo = batch.spatial_forces(1e-6)
    # Bottom face at z = -0.5 (pressure 0.5, area 1) plus the four half-submerged side faces, whose normals are
    # horizontal: they add no vertical force; with mu = 1 they add friction, so only f_z is compared exactly.
    assert abs(F_Bo[0][5] - 0.5) < 1e-12
    np.testing.assert_allclose(F_Ao[0][3:6] + F_Bo[0][3:6], 0, atol=1e-12)


def test_cpp_adapter_contact_forces():
    """ComputeHydroelasticContactForces of the C++ mirror (tests/cpp/test_traction_engine.cc)."""
    import subprocess
    from drake_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = build.build()
    exe = os.path.join(root, "tests", "cpp", "_build", "test_traction_engine")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "test_traction_engine.cc"), "-o", exe, "-L", os.path.dirname(lib),
                    "-ldrake_b200_hydro", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_spatial_forces_with_a_half_space_surface(ctx, rep):
    """Mesh pairs and a half-space pair in one scene, forces of every surface against the oracle (the half-space
    surface's vertices are per polygon here and welded in the oracle; its face quantities agree)."""
    import math
    import drake_b200 as db
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.25)
    fld = ho.VolumeField.sphere(soft, R, 1e8)
    rigid = ho.TriMesh.sphere(R, 0.1)
    a_s, a_r, a_h = parity.register_compliant(ctx, soft, fld), parity.register_rigid(ctx, rigid), ctx.register_half_space()
    poses = np.stack([ho.xform(None, [0, 0, 0]), ho.xform(ho.rpy_matrix(math.pi, 0, 0), [0, 0, 0.375]),
                      ho.xform(None, [0, 0.2, 0.2]), ho.xform(None, [0, 0, 0])])
    ids = [10, 20, 30, 40]
    batch = db.SceneBatch(ctx, [0, 4], [a_r, a_s, a_s, a_h], ids, [0, 0, 0, 1], [(0, 3)])
    osc = ho.Scene()
    osc.add_rigid(10, rigid)
    osc.add_compliant(20, soft, fld)
    osc.add_compliant(30, soft, fld)
    osc.add_half_space(40, anchored=True)
    osc.exclude(10, 40)
    E, dis, mu = np.array([np.inf, 1e8, 1e8, np.inf]), np.array([0.0, 1.0, 2.0, 0.0]), np.array([0.3, 0.5, 1.0, 0.8])
    V = np.array([[0, 0, 0, 0, 0, 0], [0.1, 0, 0.2, 0, 0.05, -0.1], [0, 0.3, 0, 0.02, 0, -0.2], [0, 0, 0, 0, 0, 0]], float)
    batch.set_contact_properties(E, dis, mu)
    batch.set_body_kinematics(None, V)
    got = batch.compute_contact_surfaces(poses, rep)[0]
    forces = batch.spatial_forces(1e-4)
    with ho.keep_handles():
        want, _ = osc.compute(poses, rep)
    assert [(s.id_M, s.id_N) for s in got] == [(s.id_M, s.id_N) for s in want] == [(10, 20), (10, 30), (20, 30), (30, 40)]
    for k, (g, w) in enumerate(zip(got, want)):
        parity.compare_surface(g, w, welded=bool(np.any(np.asarray(w.elem_second) < 0)))
        gM, gN = ids.index(w.id_M), ids.index(w.id_N)
        ref = ho.traction(w, poses[gM][9:12], poses[gN][9:12], V[gM], V[gN],
                          ho.combined_dissipation(E[gM], E[gN], dis[gM], dis[gN]), ho.combined_friction(mu[gM], mu[gN]), 1e-4)
        scale = max(np.max(np.abs(np.concatenate(ref))), 1e-300)
        for f_gpu, f_ref in zip(forces, ref):
            assert np.max(np.abs(f_gpu[k] - f_ref)) <= 1e-9 * scale


@pytest.mark.parametrize("rep", [ho.KPOLYGON, ho.KTRIANGLE])
def test_discrete_contact_pairs_match_the_oracle(ctx, rep):
    """k_discrete (the geometry-only half of AppendDiscreteContactPairsForHydroelasticContact,
    discrete_update_manager.cc:896-1036) over every surface of 24 ball / plate / floor scenes -- soft-rigid in both
    id orders and soft-soft -- against the oracle's restatement, face by face (faces matched by their emitting element
    pair): p_WC, nhat, phi0, fn0, stiffness to 1e-10; the same faces skipped."""
    from drake_b200 import engine, workloads as W
    wl = W.config3(24)
    batch = wl.register(ctx)
    batch.set_poses(wl.poses.reshape(-1, 12))
    batch.query(rep, engine.STRICT)
    batch.compute_discrete_pairs()
    batch.sync()
    pairs = batch.copy_discrete_pairs()
    res = batch.copy_results()
    surfaces = res.surfaces()
    assert len(pairs) == batch.sizes().num_faces
    osc, order = parity.oracle_scene_for(wl)
    by_scene = {}
    for k, s in enumerate(surfaces):
        by_scene.setdefault(s.scene, []).append((k, s))
    kept = 0
    for sc in range(wl.num_scenes):
        want, _ = osc.compute(wl.poses[sc][order], rep)
        got = by_scene.get(sc, [])
        assert [(s.id_M, s.id_N) for _, s in got] == [(w.id_M, w.id_N) for w in want]
        for (k, g), w in zip(got, want):
            f0 = int(res.surf_face_offset[k])
            mine = pairs[f0:f0 + g.num_faces]
            ref = ho.discrete_pairs(w)
            go = np.argsort(parity.face_keys(g.elem_M, g.elem_N), kind="stable")
            wo = np.argsort(parity.face_keys(w.elem_first, w.elem_second), kind="stable")
            if rep == ho.KTRIANGLE:   # several triangles per element pair: order them by centroid as well
                go = np.lexsort((g.centroids[:, 2], g.centroids[:, 1], g.centroids[:, 0], parity.face_keys(g.elem_M, g.elem_N)))
                wo = np.lexsort((w.centroids[:, 2], w.centroids[:, 1], w.centroids[:, 0], parity.face_keys(w.elem_first, w.elem_second)))
            a, b = mine[go], ref[wo]
            np.testing.assert_array_equal(a[:, 8] == 0, b[:, 8] == 0)
            L = max(np.max(np.abs(b[:, 0:3])), 1e-300)
            assert np.max(np.abs(a[:, 0:3] - b[:, 0:3])) <= 1e-10 * L
            assert np.max(np.abs(a[:, 3:6] - b[:, 3:6])) <= 1e-10
            for col in (6, 7, 8):
                scale = max(np.max(np.abs(b[:, col])), 1e-300)
                assert np.max(np.abs(a[:, col] - b[:, col])) <= 1e-10 * scale, col
            kept += int(np.count_nonzero(b[:, 8]))
    assert kept > 1000
    assert batch.stats()["discrete_pairs"] == kept
