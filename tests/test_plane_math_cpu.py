"""The per-element arithmetic of the half-space kernels (drake_b200/csrc/hydro_plane.cuh, __host__ __device__)
compiled with g++ (tests/cpp/plane_host_check.cc) and compared with the oracle on this CPU-only machine. The GPU
tests (test_zz_gpu_halfspace.py) run the same expressions through k_plane and the C-ABI; this file is what
lets them be checked before a GPU is available. Test infrastructure only: not a CPU path of the product."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import ho
from tests.parity import REL_TOL, assert_close, face_keys, random_pose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "plane_host_check.cc")
HDRS = [os.path.join(ROOT, "drake_b200", "csrc", n) for n in ("hydro_plane.cuh", "hydro_wrench.cuh", "hydro_clip.cuh",
                                                              "hydro_math.cuh", "hydro_types.h", "hydro_bfs.cuh",
                                                              "hydro_derive.cuh")]
LIB = os.path.join(ROOT, "tests", "cpp", "_build", "libplane_host_check.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def harness():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        # -ffp-contract=off: the device code is compiled with -fmad=false.
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
                        "-Wno-unknown-pragmas", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


class HostSurface:
    """What k_plane would pack for one geometry pair (faces in element order)."""

    def __init__(self, rep, cap_f, cap_v):
        self.rep = rep
        self.elem = np.zeros(cap_f, np.int32)
        self.face_n = np.zeros(cap_f, np.int32)
        self.areas = np.zeros(cap_f)
        self.normals = np.zeros((cap_f, 3))
        self.centroids = np.zeros((cap_f, 3))
        self.e_c = np.zeros(cap_f)
        self.grad = np.zeros((cap_f, 3))
        self.vertices = np.zeros((cap_v, 3))
        self.pressures = np.zeros(cap_v)
        self.face_data = np.zeros(4 * cap_f + cap_v, np.int32)
        self.counts = np.zeros(4, np.int32)

    def args(self):
        d = lambda a: a.ctypes.data_as(_dp)
        i = lambda a: a.ctypes.data_as(_ip)
        return [i(self.elem), i(self.face_n), d(self.areas), d(self.normals), d(self.centroids), d(self.e_c),
                d(self.grad), d(self.vertices), d(self.pressures), i(self.face_data), i(self.counts)]

    def trim(self):
        nf, nv, nfd, _ = (int(c) for c in self.counts)
        for name in ("elem", "face_n", "areas", "normals", "centroids", "e_c", "grad"):
            setattr(self, name, getattr(self, name)[:nf])
        self.vertices, self.pressures, self.face_data = self.vertices[:nv], self.pressures[:nv], self.face_data[:nfd]
        return self

    def polygons(self):
        if self.rep == ho.KTRIANGLE:
            return list(self.face_data.reshape(-1, 3))
        out, i = [], 0
        while i < len(self.face_data):
            n = self.face_data[i]
            out.append(self.face_data[i + 1:i + 1 + n])
            i += n + 1
        return out


def host_mesh_vs_half_space(mesh, fld, X_WS, X_WR, rep, swapped):
    verts, tets = mesh.arrays()[:2]
    values, grads = fld.arrays()[:2]
    tet_v = np.ascontiguousarray(verts[tets].reshape(-1, 12))
    tet_vp = np.ascontiguousarray(values[tets])
    hs = HostSurface(rep, 4 * len(tets), 5 * len(tets))
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    rc = harness().phc_mesh_vs_rigid_half_space(len(tets), d(tet_v), d(grads), d(tet_vp), d(X_WS), d(X_WR), rep,
                                                int(swapped), *hs.args())
    assert rc == 0
    return hs.trim()


def host_half_space_vs_mesh(mesh, X_WS, pressure_scale, margin, X_WR, rep, swapped, filter_normals=True):
    verts, tris, normals = mesh.arrays()[:3]
    tri_v = np.ascontiguousarray(verts[tris].reshape(-1, 9))
    hs = HostSurface(rep, 4 * len(tris), 5 * len(tris))
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    lib = harness()
    lib.phc_half_space_vs_rigid_mesh.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int, C.c_int,
                                                 C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    rc = lib.phc_half_space_vs_rigid_mesh(len(tris), d(tri_v), d(normals), d(X_WS), pressure_scale, margin, d(X_WR),
                                          rep, int(swapped), int(filter_normals), *hs.args())
    assert rc == 0
    return hs.trim()


def polygon_areas_behind_faces(surface_polygons, vertices, rep, keys_sorted, order):
    """Area of the contact polygon behind every face (sorted order). Polygons: the face itself. Triangle fans:
    the polygon is recovered from the second vertex of each of the n consecutive fan triangles of one element."""
    out = np.zeros(len(order))
    if rep == ho.KPOLYGON:
        for j, f in enumerate(order):
            P = vertices[np.asarray(surface_polygons[f])]
            out[j] = 0.5 * np.linalg.norm(sum((np.cross(P[k] - P[0], P[k + 1] - P[0]) for k in range(1, len(P) - 1)),
                                              np.zeros(3)))
        return out
    j = 0
    while j < len(order):
        k = j
        while k < len(order) and keys_sorted[k] == keys_sorted[j]:
            k += 1
        P = np.array([vertices[surface_polygons[f][1]] for f in order[j:k]])
        area = 0.5 * np.linalg.norm(sum((np.cross(P[i] - P[0], P[i + 1] - P[0]) for i in range(1, len(P) - 1)), np.zeros(3)))
        out[j:k] = area
        j = k
    return out


def compare_with_oracle(host, ora, elem_of_oracle, grad_of_oracle):
    """Faces matched by emitting element (stable: a polygon's fan triangles keep the builder's order).

    A degenerate contact polygon (a vertex, an edge or a whole mesh face exactly on the boundary: area ~ 0) has no
    meaningful normal or centroid -- they are quotients of the rounding noise of its cross products
    (polygon_surface_mesh.cc:189-207, contact_surface_utility.cc:218-286), which differs between the welded
    vertices of the oracle / reference and the per-polygon vertices of the kernels -- and in the triangle
    representation neither have the fan triangles hung on that centroid. Such polygons are compared by their
    real vertices, pressures and gradient only."""
    assert host.counts[0] == ora.num_faces
    go, oo = np.argsort(host.elem, kind="stable"), np.argsort(elem_of_oracle, kind="stable")
    keys = np.asarray(elem_of_oracle)[oo]
    np.testing.assert_array_equal(host.elem[go], keys)  # bit-exact emitted set
    L = max(np.max(np.abs(ora.vertices[np.isfinite(ora.vertices).all(axis=1)])), 1e-300)
    pscale = max(np.max(np.abs(ora.pressures)), 1e-300)
    hp, op = host.polygons(), ora.polygons()
    poly_area = polygon_areas_behind_faces(op, ora.vertices, ora.rep, keys, oo)
    solid = poly_area > 1e-6 * max(poly_area.max(), 1e-300)
    degenerate_polygon = ~solid
    # ... and a fan triangle between two coincident polygon vertices is degenerate inside a solid polygon.
    solid &= ora.areas[oo] > 1e-6 * max(poly_area.max(), 1e-300)
    assert_close("areas", host.areas[go][solid], ora.areas[oo][solid], scale=max(ora.areas[oo][solid].max(initial=0), 1e-300))
    if ora.rep == ho.KPOLYGON:
        assert_close("areas of degenerate polygons", host.areas[go][~solid], ora.areas[oo][~solid],
                     scale=max(poly_area.max(), 1e-300))
    else:
        thin = ~solid & ~degenerate_polygon
        assert_close("areas of degenerate fan triangles", host.areas[go][thin], ora.areas[oo][thin],
                     scale=max(poly_area.max(), 1e-300))
    assert_close("normals", host.normals[go][solid], ora.normals[oo][solid], scale=1.0)
    assert_close("centroids", host.centroids[go][solid], ora.centroids[oo][solid], scale=L)
    assert_close("pressure_at_centroid", host.e_c[go][solid], ora.e_centroid[oo][solid], scale=pscale)
    assert_close("grad", host.grad[go], grad_of_oracle[oo])
    for a, b, bad in zip(go, oo, degenerate_polygon):
        ia, ib = np.asarray(hp[a]), np.asarray(op[b])
        assert len(ia) == len(ib)
        if bad and ora.rep == ho.KTRIANGLE:
            ia, ib = ia[:2], ib[:2]   # the fan's third vertex is the degenerate polygon's (meaningless) centroid
        # The oracle welds vertices (as the reference does), the kernel keeps them per polygon.
        assert_close("vertices", host.vertices[ia], ora.vertices[ib], scale=L)
        assert_close("pressures", host.pressures[ia], ora.pressures[ib], scale=pscale)
    assert_close("total_area", host.areas[solid[np.argsort(go)]].sum() if False else host.areas[go][solid].sum(),
                 ora.areas[oo][solid].sum())


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("swapped", [False, True])
def test_mesh_vs_rigid_half_space_matches_the_oracle(rep, swapped):
    rng = np.random.default_rng(11)
    mesh = ho.VolumeMesh.ellipsoid(0.3, 0.4, 0.5, 0.2, dense=True)
    fld = ho.VolumeField.ellipsoid(mesh, 0.3, 0.4, 0.5, 1e5)
    id_S, id_R = (2, 1) if swapped else (1, 2)
    hits = 0
    for _ in range(12):
        X_WS = random_pose(rng, 0.5)
        X_WR = random_pose(rng, 0.5)
        ora = ho.soft_mesh_rigid_half_space(mesh, fld, X_WS, X_WR, id_S, id_R, rep)
        host = host_mesh_vs_half_space(mesh, fld, X_WS, X_WR, rep, swapped)
        if ora is None:
            assert host.counts[0] == 0
            continue
        hits += 1
        compare_with_oracle(host, ora, ora.elem_first, ora.grad_eN if swapped else ora.grad_eM)
    assert hits >= 4


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
@pytest.mark.parametrize("swapped", [False, True])
def test_half_space_vs_rigid_mesh_matches_the_oracle(rep, swapped):
    rng = np.random.default_rng(12)
    mesh = ho.VolumeMesh.ellipsoid(0.3, 0.4, 0.5, 0.2, dense=True).to_surface()
    id_S, id_R = (2, 1) if swapped else (1, 2)
    hits = 0
    for k in range(12):
        X_WS = random_pose(rng, 0.5)
        X_WR = random_pose(rng, 0.5)
        margin = 0.0 if k % 2 else 0.01
        ora = ho.soft_half_space_rigid_mesh(mesh, X_WS, 4e5, X_WR, id_S, id_R, rep, margin=margin)
        host = host_half_space_vs_mesh(mesh, X_WS, 4e5, margin, X_WR, rep, swapped)
        if ora is None:
            assert host.counts[0] == 0
            continue
        hits += 1
        compare_with_oracle(host, ora, ora.elem_second, ora.grad_eN if swapped else ora.grad_eM)
    assert hits >= 4


def test_reference_fixtures_through_the_kernel_arithmetic():
    # mesh_plane_intersection_test.cc:1217-1262 (six triangles, x = 0.5 in S) and
    # mesh_half_space_intersection_test.cc:886-1007 (10 polygons / 34 triangles).
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1]], float)
    mesh = ho.VolumeMesh.from_arrays(v, [[0, 1, 2, 3], [0, 2, 1, 4]])
    fld = ho.VolumeField.from_values(mesh, [0.25, 0.5, 0.75, 1, -1])
    X_WS = ho.xform(ho.angle_axis_matrix(math.pi / 3, np.array([-1, 1, 1]) / math.sqrt(3)), [1.25, 2.5, -3.75])
    X_WR = ho.compose(X_WS, ho.xform(ho.angle_axis_matrix(math.pi / 2, [0, 1, 0]), [0.5, 2.25, -1.25]))
    host = host_mesh_vs_half_space(mesh, fld, X_WS, X_WR, ho.KTRIANGLE, False)
    assert list(host.counts) == [6, 8, 18, 2]
    R_WS, p_WS = ho.xform_parts(X_WS)
    np.testing.assert_allclose(host.normals, np.tile(R_WS[:, 0], (6, 1)), atol=16 * np.finfo(float).eps)
    np.testing.assert_allclose(((host.vertices - p_WS) @ R_WS)[:, 0], 0.5, atol=16 * np.finfo(float).eps)
    from tests.test_oracle_halfspace import box_mesh
    X_WF = ho.xform(ho.angle_axis_matrix(math.pi / 7, np.array([1, 2, 3]) / math.sqrt(14)), [-0.25, 0.5, 0.75])
    X_FB = ho.xform(ho.rpy_matrix(math.pi / 4, math.pi / 4, math.pi / 4), [1.0, 2.0, 3.0])
    box = box_mesh(X_FB)
    X_WH = ho.compose(X_WF, X_FB)
    host = host_half_space_vs_mesh(box, X_WH, 17.0, 0.0, X_WF, ho.KPOLYGON, False)
    assert list(host.counts) == [10, 6 * 3 + 4 * 4, 10 + 6 * 3 + 4 * 4, 10]
    np.testing.assert_allclose(host.areas.sum(), 12.0, atol=1e-12)
    host = host_half_space_vs_mesh(box, X_WH, 17.0, 0.0, X_WF, ho.KTRIANGLE, False)
    assert list(host.counts) == [34, 34 + 10, 3 * 34, 10]


def test_work_item_to_pair_lookup():
    """find_plane_pair (interpolated guess, <= 4-step walk, binary-search fallback) for equal, mixed and wildly
    different mesh sizes, on whole lists and on chunk sub-ranges."""
    lib = harness()
    lp = C.POINTER(C.c_longlong)
    lib.phc_check_find_plane_pair.argtypes = [C.c_int, lp, C.c_int, C.c_int]
    rng = np.random.default_rng(2)
    for sizes in ([512] * 200, [24, 512, 116] * 50, list(rng.integers(1, 40, size=300)),
                  [1] * 100 + [5000] + [1] * 100, [3000, 1, 1, 1, 1, 1, 1, 1, 2000], [7]):
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n = len(sizes)
        for qb, qe in [(0, n), (n // 3, n), (0, max(1, n // 2)), (n // 4, max(n // 4 + 1, 3 * n // 4))]:
            assert lib.phc_check_find_plane_pair(n, off.ctypes.data_as(lp), qb, qe) == 0, (sizes[:5], qb, qe)


def host_wrench(surface, p_WAo, p_WBo, V_WA, V_WB, dissipation, mu, vslip=1e-6):
    """k_wrench's per-face arithmetic (hydro_wrench.cuh) summed in face order over an oracle Surface."""
    lib = harness()
    lib.phc_surface_wrench.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, C.c_double, C.c_double, _dp]
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    bodies = np.concatenate([p_WAo, p_WBo, V_WA[:3], V_WA[3:], V_WB[:3], V_WB[3:], surface.centroid, [dissipation]])
    tris = np.ascontiguousarray(surface.face_data if surface.rep == ho.KTRIANGLE else np.zeros((1, 3)), np.int32)
    out = np.zeros(18)
    lib.phc_surface_wrench(surface.rep, surface.num_faces, d(surface.areas), d(surface.normals), d(surface.centroids),
                           d(surface.e_centroid), d(surface.vertices), d(surface.pressures), tris.ctypes.data_as(_ip),
                           d(bodies), mu, vslip, d(out))
    return out[0:6], out[6:12], out[12:18]


@pytest.mark.parametrize("rep", [ho.KTRIANGLE, ho.KPOLYGON])
def test_wrench_arithmetic_matches_the_oracle(rep):
    """hydro_wrench.cuh against oracle/ho_traction.h on real contact surfaces (soft sphere on a rigid sphere,
    soft sphere on a half space, both id orders), random body velocities, friction and dissipation: same
    expressions in the same face order, so the agreement is to the last bits."""
    rng = np.random.default_rng(21)
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.08)
    fld = ho.VolumeField.sphere(soft, R, 1e6)
    rigid = ho.TriMesh.sphere(R, 0.05)
    X_WA = ho.xform(ho.rpy_matrix(0.3, 0.1, -0.2), [0.01, -0.02, 0.0])
    X_WB = ho.xform(None, [0.02, 0.01, -(R + R) + 0.08])
    cases = []
    with ho.keep_handles():
        for ids in ((1, 2), (2, 1)):
            cases.append(ho.soft_rigid(soft, fld, rigid, X_WA, X_WB, ids[0], ids[1], rep))
            cases.append(ho.soft_mesh_rigid_half_space(soft, fld, X_WA, ho.xform(None, [0, 0, -0.2]), ids[0], ids[1], rep))
        cases.append(ho.soft_soft(soft, fld, soft, fld, X_WA, X_WB, 1, 2, rep))
    for s in cases:
        assert s is not None and s.num_faces > 10
        for _ in range(3):
            p_WAo, p_WBo = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
            V_WA, V_WB = rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6)
            d, mu = float(rng.uniform(0, 2)), float(rng.uniform(0, 1))
            want = ho.traction(s, p_WAo, p_WBo, V_WA, V_WB, d, mu)
            got = host_wrench(s, p_WAo, p_WBo, V_WA, V_WB, d, mu)
            scale = max(np.max(np.abs(want[0])), 1e-300)
            for g, w in zip(got, want):
                assert np.max(np.abs(g - w)) <= 1e-13 * scale
    lib = harness()
    for name in ("phc_combined_dissipation", "phc_combined_friction"):
        getattr(lib, name).restype = C.c_double
    lib.phc_combined_dissipation.argtypes = [C.c_double] * 4
    lib.phc_combined_friction.argtypes = [C.c_double] * 2
    inf = float("inf")
    for args in [(inf, 1e5, 3.0, 7.0), (1e5, inf, 3.0, 7.0), (inf, inf, 3.0, 7.0), (1e5, 3e5, 4.0, 8.0), (2.0, 8.0, 1.0, 4.0)]:
        assert lib.phc_combined_dissipation(*args) == ho.combined_dissipation(*args)
    for args in [(0.0, 0.0), (0.5, 0.5), (1.0, 0.25)]:
        assert lib.phc_combined_friction(*args) == ho.combined_friction(*args)


def check_deformable_rows(rows, tri_index, want, atol_scale=1.0):
    """Rows of k_deformable's layout (p_WC, nhat_W, R_WC column-major, pressure, grad_W, barycentric) against
    oracle/ho_deformable.py's DeformableRigidContact, faces matched by the caller."""
    assert rows.shape[0] == len(want.pressures) and rows.shape[1] in (22, 24) and np.array_equal(tri_index, want.triangle_index)
    L = max(np.max(np.abs(want.contact_points_W)), 1e-300)
    assert np.max(np.abs(rows[:, 0:3] - want.contact_points_W)) <= 1e-12 * L * atol_scale
    assert np.max(np.abs(rows[:, 3:6] - want.nhats_W)) <= 1e-12 * atol_scale
    R = rows[:, 6:15].reshape(-1, 3, 3).transpose(0, 2, 1)          # column-major -> [row, col]
    assert np.max(np.abs(R - want.R_WCs)) <= 1e-11 * atol_scale
    P = max(np.max(np.abs(want.pressures)), 1e-300)
    assert np.max(np.abs(rows[:, 15] - want.pressures)) <= 1e-10 * P
    G = max(np.max(np.abs(want.pressure_gradients_W)), 1e-300)
    assert np.max(np.abs(rows[:, 16:19] - want.pressure_gradients_W)) <= 1e-10 * G
    assert np.max(np.abs(rows[:, 19:22] - want.barycentric_centroids)) <= 1e-9 * atol_scale
    assert np.max(np.abs(rows[:, 19:22].sum(axis=1) - 1.0)) <= 1e-12


def check_deformable_pair_rows(rows, want, atol_scale=1.0):
    """Rows of k_deformable's tet-tet layout (p_WC, nhat_W, R_WC column-major, signed distance, barycentric[4] x 2) against
    oracle/ho_deformable.py's DeformableDeformableContact, faces matched by the caller."""
    assert rows.shape == (len(want.signed_distances), 24)
    L = max(np.max(np.abs(want.contact_points_W)), 1e-300)
    assert np.max(np.abs(rows[:, 0:3] - want.contact_points_W)) <= 1e-12 * L * atol_scale
    assert np.max(np.abs(rows[:, 3:6] - want.nhats_W)) <= 1e-12 * atol_scale
    R = rows[:, 6:15].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.max(np.abs(R - want.R_WCs)) <= 1e-11 * atol_scale
    D = max(np.max(np.abs(want.signed_distances)), 1e-300)
    assert np.max(np.abs(rows[:, 15] - want.signed_distances)) <= 1e-10 * D
    # Barycentric coordinates of a point in a thin tetrahedron amplify rounding by its aspect ratio: compare the points
    # they reproduce as well as the coordinates.
    assert np.max(np.abs(rows[:, 16:20] - want.barycentric0)) <= 1e-8 * atol_scale
    assert np.max(np.abs(rows[:, 20:24] - want.barycentric1)) <= 1e-8 * atol_scale
    assert np.max(np.abs(rows[:, 16:20].sum(axis=1) - 1.0)) <= 1e-10 and np.max(np.abs(rows[:, 20:24].sum(axis=1) - 1.0)) <= 1e-10


def sphere_sdf(mesh, r):
    """Signed distances of a sphere mesh's vertices at rest (negative inside, 0 on the surface)."""
    return np.linalg.norm(mesh.verts, axis=1) - r


def test_deformable_pair_arithmetic_matches_the_oracle():
    """hydro_wrench.cuh's deformable_pair_contact (Cramer's rule for VolumeMesh::CalcBarycentric) on the faces of the
    equilibrium surface of two signed-distance fields against oracle/ho_deformable.py, incl. the reference's two-tetrahedra
    fixture (deformable_field_intersection_test.cc:114-173)."""
    from oracle import ho_deformable as hd
    lib = harness()
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)

    def host_rows(want, m0, m1):
        s = want.surface
        v0, t0 = m0.arrays()[:2]
        v1, t1 = m1.arrays()[:2]
        out = np.zeros((s.num_faces, 24))
        lib.phc_deformable_pair_contact(s.num_faces, d(s.normals), d(s.centroids), d(s.e_centroid), d(ho.IDENTITY), d(ho.IDENTITY),
                                        d(v0[t0[want.tet0]].reshape(-1, 12)), d(v1[t1[want.tet1]].reshape(-1, 12)), d(out))
        return out

    d0 = ho.VolumeMesh.from_arrays([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 2, 3]])
    d1 = ho.VolumeMesh.from_arrays(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, -1.0]]) + [0, 0, 0.2], [[1, 0, 2, 3]])
    sdf = [0.0, 0.0, 0.0, -1.0]
    want = hd.deformable_deformable_contact(d1, sdf, d0, sdf)
    rows = host_rows(want, d1, d0)
    check_deformable_pair_rows(rows, want)
    assert np.array_equal(rows[0, 3:6], [0, 0, 1.0]) and abs(rows[0, 15] + 0.1) < 1e-15

    a = ho.VolumeMesh.sphere(0.25, 0.08)
    b = ho.VolumeMesh.sphere(0.2, 0.07)
    sa, sb = sphere_sdf(a, 0.25), sphere_sdf(b, 0.2)
    b = ho.VolumeMesh.from_arrays(b.verts + [0.1, 0.05, 0.38], b.tets)
    want = hd.deformable_deformable_contact(a, sa, b, sb)
    assert want is not None and len(want.signed_distances) > 30
    check_deformable_pair_rows(host_rows(want, a, b), want)
    assert want.signed_distances.max() <= 1e-12       # the equilibrium surface lies inside both bodies


def test_deformable_contact_arithmetic_matches_the_oracle():
    """hydro_wrench.cuh's deformable_face_contact (MakeFromOneUnitVector, CalcBarycentric through the normal equations,
    the reference's point-transformed gradient) on the faces of real tet-triangle surfaces against
    oracle/ho_deformable.py, incl. the reference's one-polygon fixture (deformable_mesh_intersection_test.cc:106-178)."""
    from oracle import ho_deformable as hd
    lib = harness()
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)

    def host_rows(want, surface_mesh, X_WR, X_WD):
        s = want.surface
        verts, tris = surface_mesh.arrays()[:2]
        tv = verts[tris[want.triangle_index]].reshape(-1, 9)
        out = np.zeros((s.num_faces, 22))
        lib.phc_deformable_contact(s.num_faces, d(s.normals), d(s.centroids), d(s.e_centroid), d(s.grad_eM), d(X_WR), d(X_WD),
                                   d(tv), d(out))
        return out

    single = ho.VolumeMesh.from_arrays([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 2, 3]])
    X_WR = ho.xform(None, [1.0 / 6.0, 1.0 / 6.0, -0.5])
    surface = single.to_surface()
    want = hd.deformable_rigid_contact(single, ho.VolumeField.from_values(single, [123.0] * 4), surface, X_WR)
    rows = host_rows(want, surface, X_WR, ho.IDENTITY)
    check_deformable_rows(rows, want.triangle_index, want)
    assert np.allclose(rows[0, 19:22], 1.0 / 3, atol=1e-14) and np.allclose(rows[0, 0:3], [1.0 / 3, 1.0 / 3, 0], atol=1e-14)

    vol = ho.VolumeMesh.sphere(0.3, 0.1, dense=True)
    fld = ho.VolumeField.sphere(vol, 0.3, 1e6)
    surface = ho.TriMesh.sphere(0.25, 0.05)
    rng = np.random.default_rng(41)
    faces = 0
    for _ in range(4):
        X_WR = random_pose(rng, 0.05)
        X_WD = ho.compose(X_WR, ho.xform(ho.rpy_matrix(*rng.uniform(-1, 1, 3)), [0.1, 0.05, 0.3]))
        want = hd.deformable_rigid_contact(vol, fld, surface, X_WR, X_WD)
        assert want is not None
        check_deformable_rows(host_rows(want, surface, X_WR, X_WD), want.triangle_index, want)
        faces += len(want.pressures)
    assert faces > 200


@pytest.mark.parametrize("name,tri,expected", __import__("tests.test_oracle_halfspace", fromlist=["x"]).SINGLE_TRIANGLE_CASES,
                         ids=[c[0] for c in __import__("tests.test_oracle_halfspace", fromlist=["x"]).SINGLE_TRIANGLE_CASES])
def test_single_triangle_cases_through_the_kernel_arithmetic(name, tri, expected):
    """The reference's MeshHalfSpaceValueTest cases (vertices and edges exactly on the boundary, the >= 0 / <= 0
    vertex choices, zero-area results) through clip_tri_by_half_space + pack_world_polygon."""
    from tests.test_oracle_halfspace import check_single_triangle_case

    class AsSurface:
        def __init__(self, h):
            self.h = h
            self.num_faces = int(h.counts[0])
            self.vertices, self.pressures, self.areas = h.vertices, h.pressures, h.areas

        def polygons(self):
            return self.h.polygons()

    def clip(mesh_F, X_WS, scale, X_WF, rep):
        return AsSurface(host_half_space_vs_mesh(mesh_F, X_WS, scale, 0.0, X_WF, rep, False, filter_normals=False))

    check_single_triangle_case(tri, expected, clip=clip)


def test_no_double_counting_through_the_kernel_arithmetic():
    """mesh_plane_intersection_test.cc:810-850 through slice_tet_with_plane / count_tet_slice: a plane coplanar with
    the face two tetrahedra share makes exactly one of them emit (`distance > 0` is strict)."""
    from tests.test_oracle_halfspace import trivial_volume_mesh
    mesh = trivial_volume_mesh()
    fld = ho.VolumeField.from_values(mesh, [0.25, 0.5, 0.75, 1, -1])
    host = host_mesh_vs_half_space(mesh, fld, ho.IDENTITY, ho.IDENTITY, ho.KPOLYGON, False)
    assert list(host.counts) == [1, 3, 4, 1] and list(host.elem) == [0]
    assert {tuple(v) for v in host.vertices + 0.0} == {(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)}
    assert sorted(host.pressures) == [0.25, 0.5, 0.75]


# ---- the exact stage's polygon primitives (hydro_clip.cuh): bit-exact against the oracle ----------------------

def host_clip(verts, normal, point):
    lib = harness()
    v = np.ascontiguousarray(verts, float).reshape(-1, 3)
    out = np.zeros((16, 3))
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    m = lib.phc_clip_polygon_by_half_space(len(v), d(v), d(normal), d(point), d(out))
    assert 0 <= m <= 16
    return out[:m]


def host_dedupe(verts):
    v = np.ascontiguousarray(verts, float).reshape(-1, 3)
    out = np.zeros((max(len(v), 1), 3))
    d = lambda a: a.ctypes.data_as(_dp)
    m = harness().phc_remove_nearly_duplicate(len(v), d(v), d(out))
    return out[:m]


def test_clip_polygon_by_half_space_is_bit_exact():
    """clip_by_half_space (the tet-triangle clipper's inner routine) against the oracle's ClipPolygonByHalfSpace:
    the reference's literal cases (mesh_intersection_test.cc:142-281, through the oracle KAT list) and random convex
    polygons with vertices on, inside and outside the boundary: identical bits, vertex for vertex."""
    n, p = [1.0, 0.0, 0.0], [2.0, 0.0, 0.0]
    square = lambda x0, x1: [[x0, 0, 0], [x1, 0, 0], [x1, 1, 0], [x0, 1, 0]]
    for poly in (square(0, 1), square(1, 2), square(3, 4), square(2, 3), square(1.5, 2.5),
                 [[1, 0, 0], [3, 0, 0], [3, 2, 0]], [[1, 0, 0], [2, 1, 0], [2, 3, 0]], [[2, 0, 0], [2, 1, 0], [2, 3, 0]]):
        want, got = ho.clip_polygon_by_half_space(poly, n, p), host_clip(poly, n, p)
        assert got.shape == want.shape and np.array_equal(got, want), poly
    rng = np.random.default_rng(31)
    for _ in range(300):
        k = int(rng.integers(3, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        poly = np.stack([np.cos(ang), np.sin(ang), np.zeros(k)], axis=1) * rng.uniform(0.1, 3)
        R = parity_rotation(rng)
        poly = poly @ R.T + rng.uniform(-1, 1, 3)
        normal = rng.normal(size=3) * rng.uniform(0.1, 10)
        point = poly[int(rng.integers(k))] if rng.uniform() < 0.3 else rng.uniform(-1, 1, 3)   # a vertex on the boundary
        want, got = ho.clip_polygon_by_half_space(poly, normal, point), host_clip(poly, normal, point)
        assert got.shape == want.shape and np.array_equal(got, want)


def parity_rotation(rng):
    R, _ = ho.xform_parts(random_pose(rng, 0.0))
    return R


def test_remove_nearly_duplicate_vertices_is_bit_exact():
    """mesh_intersection_test.cc:285-360 and random runs of near-duplicates (|p - q|^2 < 1e-28, first vs last)."""
    eps = np.finfo(float).eps
    cases = [[[0, 0, 0], [0, 1, 0], [0, 1, 1], [0, 0, 1]], [[2, 0, 0]] * 3, [[2, 0, 0], [2, 0, 0], [2, 2, 0], [2, 2, 0]],
             [[2, 0, 0], [2, 0, 1], [2, 1, 0], [2, 0, 2 * eps]], [[1, 2, 3]], []]
    rng = np.random.default_rng(32)
    for _ in range(200):
        k = int(rng.integers(1, 9))
        base = rng.uniform(-2, 2, size=(k, 3))
        reps = rng.integers(1, 4, size=k)
        v = np.repeat(base, reps, axis=0) + rng.choice([0.0, 1e-15, 3e-14], size=(int(reps.sum()), 1))
        if rng.uniform() < 0.5:
            v = np.concatenate([v, v[:1] + 1e-16])
        cases.append(v[:16])
    for v in cases:
        want = ho.remove_nearly_duplicate_vertices(v) if len(v) else np.zeros((0, 3))
        got = host_dedupe(v) if len(v) else np.zeros((0, 3))
        assert got.shape == want.shape and np.array_equal(got, want)


def test_intersect_tetrahedra_is_bit_exact():
    """The tet-tet exact stage's sequence (slice by the equilibrium plane, dedupe, clip by the four faces with
    clip_by_half_space_edges, dedupe) on the candidate pairs of two overlapping soft spheres, against the oracle's
    IntersectTetrahedra (field_intersection.cc:85-198): the same polygon, bit for bit, for every candidate; and the
    reference's octagon (field_intersection_test.cc:196-259)."""
    lib = harness()
    lib.phc_intersect_tetrahedra.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, _dp]
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)

    def host(m0, e0, m1, e1, X_MN, nhat, disp):
        v0, t0, _ = m0.arrays()
        v1, t1, inward1 = m1.arrays()
        out = np.zeros((16, 3))
        n = lib.phc_intersect_tetrahedra(d(v0[t0[e0]]), d(v1[t1[e1]]), d(inward1[e1]), d(X_MN), d(nhat), disp, d(out))
        assert n >= 0
        return out[:n]

    m0 = ho.VolumeMesh.from_arrays([[2, 0, 2], [-2, 0, 2], [0, 2, -2], [0, -2, -2]], [[0, 1, 2, 3]])
    m1 = ho.VolumeMesh.from_arrays([[1.5, 1.5, 2.5], [-1.5, -1.5, 2.5], [-1.5, 1.5, -2.5], [1.5, -1.5, -2.5]],
                                   [[0, 1, 2, 3]])
    got = host(m0, 0, m1, 0, ho.IDENTITY, [0, 0, 1.0], 0.0)
    want, _ = ho.intersect_tetrahedra(m0, 0, m1, 0, ho.IDENTITY, [0, 0, 1.0], 0.0)
    assert len(got) == 8 and np.array_equal(got, want)
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.1)
    fld = ho.VolumeField.sphere(soft, R, 1e6)
    rng = np.random.default_rng(33)
    polygons = 0
    for _ in range(3):
        X_MN = ho.invert_and_compose(random_pose(rng, 0.02), random_pose(rng, 0.02, center=(0, 0, 0.3)))
        pairs, _ = ho.candidates(soft, soft, X_MN)
        for e0, e1 in pairs[::7]:
            plane = ho.calc_equilibrium_plane(fld, int(e0), fld, int(e1), X_MN)
            if plane is None:
                continue
            nhat, disp = plane
            want, _ = ho.intersect_tetrahedra(soft, int(e0), soft, int(e1), X_MN, nhat, disp)
            got = host(soft, int(e0), soft, int(e1), X_MN, nhat, disp)
            assert got.shape == want.shape and np.array_equal(got, want), (e0, e1)
            polygons += len(want) > 0
    assert polygons >= 50


def test_clip_triangle_by_tetrahedron_is_bit_exact():
    """The tet-triangle exact stage's sequence against the oracle's ClipTriangleByTetrahedron
    (mesh_intersection.cc:176-264): the reference's heptagon (mesh_intersection_test.cc:660-709) and every candidate
    pair of a soft sphere on a rigid sphere, bit for bit; plus the band-limited cos filter against the plain one."""
    lib = harness()
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)

    def host(vol, tet, surf, tri, X_MN):
        v, t, _ = vol.arrays()
        sv, st = surf.arrays()[:2]
        out = np.zeros((16, 3))
        n = lib.phc_clip_triangle_by_tetrahedron(d(v[t[tet]]), d(sv[st[tri]]), d(X_MN), d(out))
        assert n >= 0
        return out[:n]

    vol = ho.VolumeMesh.from_arrays([[2, 0, 2], [-2, 0, 2], [0, 2, -2], [0, -2, -2]], [[0, 1, 2, 3]])
    surf = ho.TriMesh.from_arrays([[1.5, 1.5, 0], [-1.5, 0, 0], [0, -1.5, 0]], [[0, 1, 2]])
    got, want = host(vol, 0, surf, 0, ho.IDENTITY), ho.clip_triangle_by_tetrahedron(vol, 0, surf, 0, ho.IDENTITY)
    assert len(got) == 7 and np.array_equal(got, want)
    R = 0.25
    soft, rigid = ho.VolumeMesh.sphere(R, 0.1), ho.TriMesh.sphere(R, 0.05)
    fld = ho.VolumeField.sphere(soft, R, 1e6)
    rng = np.random.default_rng(34)
    polygons = 0
    for _ in range(3):
        X_MN = ho.invert_and_compose(random_pose(rng, 0.02), random_pose(rng, 0.02, center=(0, 0, 0.4)))
        pairs, _ = ho.candidates(soft, rigid, X_MN)
        for tet, tri in pairs[::5]:
            want = ho.clip_triangle_by_tetrahedron(soft, int(tet), rigid, int(tri), X_MN)
            got = host(soft, int(tet), rigid, int(tri), X_MN)
            assert got.shape == want.shape and np.array_equal(got, want), (tet, tri)
            polygons += len(want) > 0
    assert polygons >= 50
    # face_normal_along_gradient: the shortcut band around cos(5 pi / 8) never changes the decision.
    lib.phc_face_normal_along_gradient.argtypes = [_dp, _dp]
    grads = fld.arrays()[1]
    R_MN, _ = ho.xform_parts(X_MN)
    normals = rigid.arrays()[2]
    for tet, tri in pairs[::3]:
        g = grads[tet] / np.linalg.norm(grads[tet])
        want = ho.is_face_normal_along_pressure_gradient(fld, rigid, X_MN, int(tet), int(tri))
        # The kernel's ngrad is normalized(grad) computed at registration with the oracle's expression; numpy's
        # normalisation can differ in the last bit, which only matters inside 1e-15 of the threshold.
        got = bool(lib.phc_face_normal_along_gradient(d(g), d(R_MN @ normals[tri])))
        cos = float(g @ (R_MN @ normals[tri]))
        assert got == want or abs(cos - math.cos(5 * math.pi / 8)) < 1e-12


def test_on_boundary_quirk_through_the_kernel_arithmetic():
    """The reference's on-boundary vertex choice (see tests/test_oracle_halfspace.py::on_boundary_quirk) is reproduced
    by clip_tri_by_half_space, not corrected."""
    from tests.test_oracle_halfspace import check_on_boundary_quirk, on_boundary_quirk

    class AsSurface:
        def __init__(self, h):
            self.h, self.num_faces = h, int(h.counts[0])
            self.vertices, self.pressures = h.vertices, h.pressures

        def polygons(self):
            return self.h.polygons()

    check_on_boundary_quirk(on_boundary_quirk(
        lambda m, X_WS, scale, X_WF, rep: AsSurface(host_half_space_vs_mesh(m, X_WS, scale, 0.0, X_WF, rep, False,
                                                                             filter_normals=False))))


def test_parity_compare_surface_on_cpu():
    """tests/parity.py::compare_surface is what every GPU parity test calls; it cannot run here against the GPU, so it
    is exercised against oracle surfaces dressed as GPU results (same arrays, vertices un-welded for the half-space
    ones, faces shuffled): the comparison code itself must be sound, and it must still catch a perturbed value."""
    import copy
    from types import SimpleNamespace
    from tests import parity

    def as_gpu(s, rng, unweld):
        polys = s.polygons()
        order = rng.permutation(len(polys)) if s.rep == ho.KPOLYGON else np.arange(len(polys))
        # Triangle fans must stay contiguous per polygon (stable matching by emitting element): only polygons shuffle.
        verts, pres, fd = [], [], []
        for f in order:
            idx = np.asarray(polys[f])
            if unweld:
                base = len(verts)
                verts += list(s.vertices[idx])
                pres += list(s.pressures[idx])
                idx = np.arange(base, base + len(idx))
            fd.append(idx)
        if not unweld:
            verts, pres = list(s.vertices), list(s.pressures)
        g = SimpleNamespace(id_M=s.id_M, id_N=s.id_N, representation=s.rep, num_faces=s.num_faces,
                            vertices=np.array(verts), pressures=np.array(pres), areas=s.areas[order],
                            normals=s.normals[order], centroids=s.centroids[order], pressure_at_centroid=s.e_centroid[order],
                            grad_eM=None if s.grad_eM is None else s.grad_eM[order],
                            grad_eN=None if s.grad_eN is None else s.grad_eN[order], total_area=s.total_area,
                            centroid=s.centroid, elem_M=s.elem_first[order], elem_N=s.elem_second[order])
        g.polygons = lambda: fd
        return g

    rng = np.random.default_rng(8)
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.1)
    fld = ho.VolumeField.sphere(soft, R, 1e6)
    rigid = ho.TriMesh.sphere(R, 0.05)
    X_WA, X_WB = ho.xform(ho.rpy_matrix(0.3, 0.1, -0.2), [0.01, -0.02, 0.0]), ho.xform(None, [0.02, 0.01, -0.42])
    for rep in (ho.KPOLYGON, ho.KTRIANGLE):
        cases = [(ho.soft_rigid(soft, fld, rigid, X_WA, X_WB, 1, 2, rep), False),
                 (ho.soft_soft(soft, fld, soft, fld, X_WA, X_WB, 1, 2, rep), False),
                 (ho.soft_mesh_rigid_half_space(soft, fld, X_WA, ho.xform(None, [0, 0, -0.2]), 2, 1, rep), True),
                 (ho.soft_half_space_rigid_mesh(rigid, ho.xform(None, [0, 0, -0.2]), 4e6, X_WA, 1, 2, rep), True)]
        for s, welded in cases:
            assert s is not None and s.num_faces > 10
            g = as_gpu(s, rng, unweld=welded)
            parity.compare_surface(g, s, welded=welded)
            bad = copy.copy(g)
            bad.areas = g.areas.copy()
            bad.areas[len(bad.areas) // 2] *= 1 + 1e-8
            with pytest.raises(AssertionError):
                parity.compare_surface(bad, s, welded=welded)
            bad = copy.copy(g)
            bad.normals = g.normals.copy()
            bad.normals[0] = -bad.normals[0]
            with pytest.raises(AssertionError):
                parity.compare_surface(bad, s, welded=welded)


def test_reference_demand_is_reported_not_hidden():
    """Two vertices exactly on the boundary and the third outside, listed so that GetVertexAddIfNeeded's
    DRAKE_DEMAND(sgn(s_a) != sgn(s_b)) (mesh_half_space_intersection.cc:66) fires in the reference: the oracle throws,
    the kernel arithmetic raises DBH_ERR_REFERENCE_DEMAND (13) instead of emitting a polygon with a 0 / 0 vertex."""
    mesh = ho.TriMesh.from_arrays([[0, 0, 0], [1, 0, 0], [0, 1, 1]], [[0, 1, 2]])   # s = (0, 0, +1)
    with pytest.raises(RuntimeError, match="DRAKE_DEMAND"):
        ho.soft_half_space_rigid_mesh(mesh, ho.IDENTITY, 1.0, ho.IDENTITY, 1, 2, ho.KPOLYGON, filter_normals=False)
    verts, tris, normals = mesh.arrays()[:3]
    hs = HostSurface(ho.KPOLYGON, 4, 5)
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    lib = harness()
    lib.phc_half_space_vs_rigid_mesh.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int, C.c_int,
                                                 C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    rc = lib.phc_half_space_vs_rigid_mesh(1, d(verts[tris].reshape(-1, 9)), d(normals), d(ho.IDENTITY), 1.0, 0.0,
                                          d(ho.IDENTITY), ho.KPOLYGON, 0, 0, *hs.args())
    assert rc == 13


def test_nine_gon_of_config4_is_bit_exact_on_the_exceptional_storage():
    """BASELINE config 4, scene 7280: a clip plane coplanar with the polygon makes IntersectTetrahedra return a 9-gon
    (the reference's "at most 8" comment assumes general position). The kernel's primitives on the exceptional path's
    16-slot storage (LocalPoly) reproduce the oracle's polygons of that body pair bit for bit, the 9-gon included."""
    from drake_b200 import workloads as W
    from tests import parity
    lib = harness()
    lib.phc_intersect_tetrahedra.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, _dp]
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
    wl = W.config4(1, 7280)
    osc, order = parity.oracle_scene_for(wl)
    surfaces, _ = osc.compute(wl.poses[0][order], ho.KPOLYGON)
    big = [s for s in surfaces if max(len(p) for p in s.polygons()) == 9]
    assert len(big) == 1
    s = big[0]
    ids = list(wl.geom_id)
    gM, gN = ids.index(s.id_M), ids.index(s.id_N)

    def oracle_mesh(g):
        v, e, p = wl.meshes[wl.geom_mesh[g]].arrays()
        m = ho.VolumeMesh.from_arrays(v, e)
        return m, ho.VolumeField.from_values(m, p)

    (m0, f0), (m1, f1) = oracle_mesh(gM), oracle_mesh(gN)
    X_MN = ho.invert_and_compose(wl.poses[0][gM], wl.poses[0][gN])
    v0, t0, _ = m0.arrays()
    v1, t1, in1 = m1.arrays()
    sizes = []
    for e0, e1 in sorted(set(zip(s.elem_first.tolist(), s.elem_second.tolist()))):
        nhat, disp = ho.calc_equilibrium_plane(f0, e0, f1, e1, X_MN)
        want, _ = ho.intersect_tetrahedra(m0, e0, m1, e1, X_MN, nhat, disp)
        out = np.zeros((16, 3))
        n = lib.phc_intersect_tetrahedra(d(v0[t0[e0]]), d(v1[t1[e1]]), d(in1[e1]), d(X_MN), d(nhat), disp, d(out))
        assert n == len(want) and np.array_equal(out[:n], want)
        sizes.append(n)
    assert max(sizes) == 9


def test_tet_pair_polygon_with_face_codes_is_bit_exact():
    """The surface-seeded intersector's per-pair function (hydro_bfs.cuh: tet_pair_polygon_with_faces, what k_bfs_wave
    runs) against the oracle's CalcContactPolygon pieces (field_intersection.cc:377-428): the same polygon bit for bit and
    the same face codes in the same order (IntersectTetrahedra's second return value, :85-198) -- the reference's
    octagon with its literal face list (field_intersection_test.cc:196-259) and the candidate pairs of two overlapping
    soft spheres under random poses, filters included."""
    lib = harness()
    lib.phc_tet_pair_polygon_faces.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)

    def host(m0, f0, e0, m1, f1, e1, X_MN):
        v0, t0, _ = m0.arrays()
        v1, t1, _ = m1.arrays()
        poly, faces, info = np.zeros((16, 3)), np.zeros(16, np.int32), np.zeros(2, np.int32)
        n = lib.phc_tet_pair_polygon_faces(d(v0[t0[e0]]), d(f0.values[t0[e0]]), d(v1[t1[e1]]), d(f1.values[t1[e1]]),
                                           d(X_MN), d(poly), faces.ctypes.data_as(_ip), info.ctypes.data_as(_ip))
        assert n >= 0, n
        return poly[:n], faces[:info[0]], int(info[1])

    def oracle(m0, f0, e0, m1, f1, e1, X_MN):
        plane = ho.calc_equilibrium_plane(f0, e0, f1, e1, X_MN)
        if plane is None:
            return np.zeros((0, 3)), np.zeros(0, np.int32)
        nhat, disp = plane
        R_NM = np.asarray(X_MN[:9], float).reshape(3, 3)   # column-major R_MN read row-major = its transpose
        if not ho.is_plane_normal_along_pressure_gradient(nhat, e0, f0):
            return np.zeros((0, 3)), np.zeros(0, np.int32)
        if not ho.is_plane_normal_along_pressure_gradient(R_NM @ (-nhat), e1, f1):
            return np.zeros((0, 3)), np.zeros(0, np.int32)
        return ho.intersect_tetrahedra(m0, e0, m1, e1, X_MN, nhat, disp)

    # The reference's octagon: fields chosen so that the equilibrium plane is z = 0 (f0 = z, f1 = -z).
    m0 = ho.VolumeMesh.from_arrays([[2, 0, 2], [-2, 0, 2], [0, 2, -2], [0, -2, -2]], [[0, 1, 2, 3]])
    m1 = ho.VolumeMesh.from_arrays([[1.5, 1.5, 2.5], [-1.5, -1.5, 2.5], [-1.5, 1.5, -2.5], [1.5, -1.5, -2.5]],
                                   [[0, 1, 2, 3]])
    f0 = ho.VolumeField.from_values(m0, m0.verts[:, 2].copy())
    f1 = ho.VolumeField.from_values(m1, -m1.verts[:, 2].copy())
    want_poly, want_faces = oracle(m0, f0, 0, m1, f1, 0, ho.IDENTITY)
    got_poly, got_faces, mask = host(m0, f0, 0, m1, f1, 0, ho.IDENTITY)
    assert len(want_poly) == 8 and np.array_equal(got_poly, want_poly)
    assert got_faces.tolist() == want_faces.tolist()
    assert sorted(set(want_faces.tolist())) == list(range(8)) and mask == 0xFF   # every face of both tetrahedra
    R = 0.25
    soft = ho.VolumeMesh.sphere(R, 0.1)
    fld = ho.VolumeField.sphere(soft, R, 1e6)
    rng = np.random.default_rng(77)
    polygons = rejected = 0
    for _ in range(3):
        X_MN = ho.invert_and_compose(random_pose(rng, 0.02), random_pose(rng, 0.02, center=(0, 0, 0.3)))
        pairs, _ = ho.candidates(soft, soft, X_MN)
        for e0, e1 in pairs[::5]:
            want_poly, want_faces = oracle(soft, fld, int(e0), soft, fld, int(e1), X_MN)
            got_poly, got_faces, mask = host(soft, fld, int(e0), soft, fld, int(e1), X_MN)
            assert got_poly.shape == want_poly.shape and np.array_equal(got_poly, want_poly), (e0, e1)
            assert got_faces.tolist() == want_faces.tolist(), (e0, e1)
            assert mask == sum(1 << c for c in set(want_faces.tolist()))
            polygons += len(want_poly) > 0
            rejected += len(want_poly) == 0
    assert polygons >= 50 and rejected >= 50


# ---- BoxesOverlap: the reference's own test, case by case ----------------------------------------------------------------

def _aa(angle, axis):
    u = np.zeros(3)
    u[axis] = 1.0
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * K @ K


def _corner_transform(a_half, b_half, axis, expect_overlap):
    """CalcCornerTransform (geometry/test_utilities/boxes_overlap_transforms.cc:12-45): B's lowest corner 0.01 inside /
    outside the middle of A's face along `axis`."""
    axis1, axis2 = (axis + 1) % 3, (axis + 2) % 3
    R_AB = _aa(math.pi / 5, axis1) @ _aa(-math.pi / 5, axis2)
    p_BqBo_A = R_AB @ b_half
    assert p_BqBo_A[axis] > 0
    p_AfBq_A = np.zeros(3)
    p_AfBq_A[axis] = -0.01 if expect_overlap else 0.01
    p_AoAf_A = np.array(a_half, float)
    p_AoAf_A[axis1] /= 2
    p_AoAf_A[axis2] /= 2
    return R_AB, p_AoAf_A + p_AfBq_A + p_BqBo_A


def _edge_transform(a_half, b_half, a_axis, b_axis, expect_overlap):
    """CalcEdgeTransform (boxes_overlap_transforms.cc:47-110): an edge of B along b_axis 0.01 inside / outside an edge of
    A along a_axis -- only the axis a_axis x b_axis separates the two."""
    a1, a2, b1, b2 = (a_axis + 1) % 3, (a_axis + 2) % 3, (b_axis + 1) % 3, (b_axis + 2) % 3
    theta = math.pi / 5
    if a_axis == b_axis:
        R_AB = _aa(theta, b1) @ _aa(theta, b2)
    elif a1 == b_axis:
        R_AB = _aa(theta, b1) @ _aa(-theta, b2)
    else:
        R_AB = _aa(-theta, b2) @ _aa(theta, b1)
    p_BqBo_B = np.array(b_half, float)
    p_BqBo_B[b_axis] -= b_half[b_axis] / 2
    p_BqBo_A = R_AB @ p_BqBo_B
    assert p_BqBo_A[a1] > 0 and p_BqBo_A[a2] > 0
    offset = -0.01 if expect_overlap else 0.01
    p_AfBq_A = np.zeros(3)
    p_AfBq_A[a1] = p_AfBq_A[a2] = offset
    p_AoAf_A = np.array(a_half, float)
    p_AoAf_A[a_axis] -= a_half[a_axis] / 2
    return R_AB, p_AoAf_A + p_AfBq_A + p_BqBo_A


def test_boxes_overlap_reference_all_cases():
    """BoxesOverlapTest.AllCases (geometry/proximity/test/boxes_overlap_test.cc:97-163): parallel and contained; for each
    of A's and of B's three face normals, and for each of the nine edge x edge axes, a pose where ONLY that axis separates
    the boxes (expect false) and its colliding twin (expect true); every case both ways round (a, b, X_AB) and
    (b, a, X_AB^-1). Through the oracle's BoxesOverlap and through the product's fp64 box test compiled for the host."""
    L = harness()
    L.phc_boxes_overlap.argtypes = [_dp, _dp, _dp]

    def product(a, b, X):
        a, b, X = (np.ascontiguousarray(v, float) for v in (a, b, X))
        return bool(L.phc_boxes_overlap(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), X.ctypes.data_as(_dp)))

    def both_ways(a, b, R, p, expect, label):
        X = ho.xform(R, p)
        X_inv = ho.xform(R.T, -R.T @ np.asarray(p, float))
        for fn in (ho.boxes_overlap, product):
            assert fn(a, b, X) == expect, (label, fn)
            assert fn(b, a, X_inv) == expect, (label, "inverse", fn)
    both_ways([1, 2, 1], [0.5, 1, 0.5], np.eye(3), [0.2, 0.4, 0.2], True, "parallel_and_contained")
    a, b = np.array([2.0, 4, 3]), np.array([3.5, 2, 1.5])
    n = 1
    for axis in range(3):
        for expect in (False, True):
            R, p = _corner_transform(a, b, axis, expect)
            both_ways(a, b, R, p, expect, f"A{axis} {expect}")
            R, p = _corner_transform(b, a, axis, expect)          # X_BA: invert for X_AB
            both_ways(a, b, R.T, -R.T @ p, expect, f"B{axis} {expect}")
            n += 2
    for a_axis in range(3):
        for b_axis in range(3):
            for expect in (False, True):
                R, p = _edge_transform(a, b, a_axis, b_axis, expect)
                both_ways(a, b, R, p, expect, f"A{a_axis}-B{b_axis} {expect}")
                n += 1
    assert n == 31


def test_tetrahedron_gradients_reference_known_answers():
    """VolumeMeshTest.TestCalcGradBarycentric and TestCalcGradientVectorOfLinearField (geometry/proximity/test/
    volume_mesh_test.cc:290-360,368-385): the tetrahedron (0,0,0), (1,0,0), (0,2,0), (0,0,3) posed at rpy (pi/6, 2pi/3, 5pi/4),
    p (1,2,3) has grad b_i = R (-1,-1/2,-1/3), R x, R y / 2, R z / 3 (1e-14), summing to zero; the field 2, 3, 4, 5 on the unit
    tetrahedron has gradient (1, 2, 3) -- through the oracle's MeshFieldLinear and the product's registration-time
    derivation (derive_tet) compiled for the host. A zero-volume tetrahedron is reported degenerate by both."""
    L = harness()
    L.phc_derive_tet.argtypes = [_dp] * 5

    def product(V, u):
        V, u = np.ascontiguousarray(V, float), np.ascontiguousarray(u, float)
        g, m, n = np.zeros(3), np.zeros(1), np.zeros(12)
        ok = L.phc_derive_tet(V.ctypes.data_as(_dp), u.ctypes.data_as(_dp), g.ctypes.data_as(_dp), m.ctypes.data_as(_dp),
                              n.ctypes.data_as(_dp))
        return bool(ok), g, float(m[0]), n.reshape(4, 3)

    def oracle(V, u):
        f = ho.VolumeField.from_values(ho.VolumeMesh.from_arrays(V, [[0, 1, 2, 3]]), u)
        _, g, m, degenerate = f.arrays()
        return (not degenerate), g[0], float(m[0]), f.mesh.arrays()[2][0]
    X_WM = ho.xform(ho.rpy_matrix(math.pi / 6, 2 * math.pi / 3, 5 * math.pi / 4), [1.0, 2.0, 3.0])
    R, p = X_WM[:9].reshape(3, 3).T, X_WM[9:]
    V_M = np.array([[0, 0, 0], [1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0]])
    V_W = V_M @ R.T + p
    want = [R @ np.array([-1, -1 / 2, -1 / 3]), R @ np.array([1.0, 0, 0]), R @ np.array([0, 0.5, 0]), R @ np.array([0, 0, 1 / 3])]
    for fn in (oracle, product):
        total = np.zeros(3)
        for i in range(4):
            ok, g, m, inward = fn(V_W, np.eye(4)[i])
            assert ok
            np.testing.assert_allclose(g, want[i], rtol=0, atol=1e-14)
            assert abs(m + g @ V_W[0] - (1.0 if i == 0 else 0.0)) <= 1e-13      # f(x) = grad . x + f(Mo) at vertex 0
            total += g
            # inward_normal(e, f) is the unit normal of the face opposite vertex f, pointing at it: along grad b_f.
            np.testing.assert_allclose(inward[i], want[i] / np.linalg.norm(want[i]), rtol=0, atol=1e-14)
        np.testing.assert_allclose(total, 0, atol=1e-14)
        ok, g, m, _ = fn([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [2.0, 3.0, 4.0, 5.0])
        assert ok and m == 2.0
        np.testing.assert_allclose(g, [1, 2, 3], rtol=0, atol=1e-14)
        assert not fn([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], [0.0, 0, 0, 1.0])[0]
    # and the two derivations are the same numbers
    for i in range(4):
        a, b = oracle(V_W, np.eye(4)[i]), product(V_W, np.eye(4)[i])
        np.testing.assert_array_equal(a[1], b[1])
        assert a[2] == b[2]
        np.testing.assert_array_equal(a[3], b[3])
