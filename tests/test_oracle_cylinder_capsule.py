"""Cylinder and Capsule hydroelastic representations (MakeCylinderVolumeMeshWithMa, MakeCapsuleVolumeMesh and their
pressure fields): the oracle's restatement (oracle/ho_cylinder_capsule.h) pinned by the reference's own tests --
make_cylinder_mesh_test.cc, make_capsule_mesh_test.cc, make_cylinder_field_test.cc, make_capsule_field_test.cc -- and
the product's host generators (drake_b200/csrc/hydro_shapes.cu) identical to it bit for bit."""
import math

import numpy as np
import pytest

import drake_b200 as db
from oracle import ho


def _boundary_vertices(T):
    """CollectUniqueVertices(IdentifyBoundaryFaces(tetrahedra)) (volume_to_surface_mesh.cc:31-150)."""
    faces = {}
    for t in T:
        for f in ((1, 2, 3), (3, 2, 0), (1, 3, 0), (2, 1, 0)):
            key = tuple(sorted(int(t[k]) for k in f))
            faces[key] = faces.get(key, 0) + 1
    return sorted({v for key, c in faces.items() if c == 1 for v in key})


def _euler_characteristic(T):
    """ComputeEulerCharacteristic (proximity_utilities.cc): V - E + F - T."""
    V = {int(v) for t in T for v in t}
    E = {tuple(sorted((int(t[a]), int(t[b])))) for t in T for a in range(4) for b in range(a + 1, 4)}
    F = {tuple(sorted(int(t[k]) for k in f)) for t in T for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))}
    return len(V) - len(E) + len(F) - len(T)


def _tet_volumes(V, T):
    a, b, c, d = (V[T[:, k]] for k in range(4))
    return np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0


def _verify_conforming(mesh):
    V, T = mesh.verts, mesh.tets
    assert len(np.unique(V, axis=0)) == len(V), "duplicated vertices"
    assert _euler_characteristic(T) == 1
    assert np.all(_tet_volumes(V, T) > 0)
    boundary = set(_boundary_vertices(T))
    assert all(any(int(v) not in boundary for v in t) for t in T), "a tetrahedron has all its vertices on the boundary"
    return boundary


def _cylinder_distance(radius, length, V):
    """Signed distance to the cylinder for points inside or on it."""
    return -np.minimum(radius - np.hypot(V[:, 0], V[:, 1]), length / 2 - np.abs(V[:, 2]))


def _capsule_distance(radius, length, V):
    z = np.clip(V[:, 2], -length / 2, length / 2)
    return np.sqrt(V[:, 0] ** 2 + V[:, 1] ** 2 + (V[:, 2] - z) ** 2) - radius


@pytest.mark.parametrize("radius,length,hint,tets_per_rim_vertex,extra_vertices", [
    (1.0, 3.0, 0.5, 5, 4),      # Long   (make_cylinder_mesh_test.cc:247-261): 5 n tets, 2 n + 4 vertices
    (1.0, 2.0, 0.5, 4, 3),      # Medium (:263-277): 4 n, 2 n + 3
    (1.0, 1.0, 0.5, 9, None),   # Short  (:279-293): 9 n, 3 n + 3
    (1.0, 1.0, 100.0, 9, None)  # Coarsest (:295-309): n = 3
])
def test_cylinder_mesh_with_medial_axis_reference_cases(radius, length, hint, tets_per_rim_vertex, extra_vertices):
    n = max(3, math.ceil(2 * math.pi * radius / hint))
    mesh = ho.VolumeMesh.cylinder_ma(radius, length, hint)
    assert mesh.num_elements == tets_per_rim_vertex * n
    assert mesh.num_vertices == (3 * n + 3 if extra_vertices is None else 2 * n + extra_vertices)
    boundary = _verify_conforming(mesh)
    V, T = mesh.verts, mesh.tets
    tol = 1e-14 * max(1.0, length / 2, radius)
    d = _cylinder_distance(radius, length, V)
    assert np.all(d <= tol)                                               # every vertex inside or on the cylinder
    assert boundary == set(np.nonzero(np.abs(d) <= tol)[0].tolist())      # and the boundary vertices are those on it
    # VerifyCylinderMeshWithMa (:150-245): every tetrahedron touches the medial axis (a vertex at depth min(r, L/2)).
    depth = min(radius, length / 2)
    assert all(np.any(np.abs(d[t] + depth) <= 1e-14) for t in T)
    # Rim vertices: 2 circles of n (3 for the short class' medial disc is interior).
    assert len(boundary) == 2 * n + 2


def test_cylinder_surface_mesh_reference_case():
    """MakeCylinderSurfaceMesh GenerateSurface (:433-445): radius 1, length 2, hint 3 -> 12 triangles, 8 vertices."""
    s = ho.VolumeMesh.cylinder_ma(1.0, 2.0, 3.0).to_surface()
    assert (s.num_elements, s.num_vertices) == (12, 8)


@pytest.mark.parametrize("length", [3.0, 2.0, 0.5])
def test_capsule_mesh_reference_cases(length):
    """Long / Medium / Short (make_capsule_mesh_test.cc:95-125) and Coarsest (:127-135)."""
    for radius, hint in ((1.0, 0.5), (1.0, 100.0)):
        mesh = ho.VolumeMesh.capsule(radius, length, hint)
        boundary = _verify_conforming(mesh)
        V, T = mesh.verts, mesh.tets
        tol = 1e-14 * max(length / 2, radius)
        d = _capsule_distance(radius, length, V)
        assert np.all(d - tol <= 0)
        assert boundary == set(range(2, len(V)))                  # everything but the two ends of the medial axis
        # IsTetrahedronRespectingMa (:15-38): a boundary vertex and a medial vertex in every tetrahedron.
        assert all(np.any(radius + d[t] < tol) and np.any(np.abs(d[t]) < tol) for t in T)


@pytest.mark.parametrize("n,hint", [(3, (2.0 / 3) * math.pi), (3, math.pi), (4, (2.0 / 4) * math.pi), (706, 2 * math.pi / 706),
                                    (706, math.pi / 706)])
def test_capsule_mesh_counts(n, hint):
    """CoarsestMesh, CoarsestMeshClamped, SecondCoarsestMesh, FinestMesh, FinestMeshClamped (:137-190):
    2 n floor(n/2) + 4 vertices, 4 n (floor(n/2) - 1) + 5 n tetrahedra."""
    mesh = ho.VolumeMesh.capsule(1.0, 2.0, hint)
    assert mesh.num_vertices == 2 * n * (n // 2) + 4
    assert mesh.num_elements == 4 * n * ((n // 2) - 1) + 5 * n


def test_capsule_volume_convergence_and_euler_characteristic():
    """VolumeConvergence (:192-220) and EulerCharacteristic (:222-240)."""
    radius, height, hint = 1.0, 2.0, 3.0
    exact = math.pi * radius * radius * ((4.0 / 3.0) * radius + height)
    m = ho.VolumeMesh.capsule(radius, height, hint)
    prev = exact - float(np.sum(_tet_volumes(m.verts, m.tets)))
    for _ in range(5):
        hint /= 2.0
        m = ho.VolumeMesh.capsule(radius, height, hint)
        err = exact - float(np.sum(_tet_volumes(m.verts, m.tets)))
        assert 0.0 < err < prev
        prev = err
    for n in (3, 4, 5, 6, 7, 8):
        assert _euler_characteristic(ho.VolumeMesh.capsule(radius, height, math.pi * radius / n).tets) == 1


@pytest.mark.parametrize("radius,length", [(1.0, 3.0), (1.0, 2.0), (1.0, 1.0)])
def test_cylinder_pressure_field_reference_cases(radius, length):
    """MakeCylinderPressureFieldWithMaLong / Medium / Short (make_cylinder_field_test.cc:93-124) through
    CheckMinMaxBoundaryValue (:22-74): pressures in [0, E], 0 on every boundary vertex, E at the centre vertex."""
    E = 1.0e5
    mesh = ho.VolumeMesh.cylinder_ma(radius, length, 0.25)
    p = ho.VolumeField.cylinder(mesh, radius, length, E).values
    assert p.max() <= E + 1e-14 * E and abs(p.max() - E) <= 1e-14 * E
    assert p.min() == 0.0
    assert np.all(p[_boundary_vertices(mesh.tets)] == 0.0)
    centre = np.nonzero(np.all(mesh.verts == 0.0, axis=1))[0]
    if len(centre):
        assert abs(p[centre[0]] - E) <= 1e-14 * E
    interior = sorted(set(range(mesh.num_vertices)) - set(_boundary_vertices(mesh.tets)))
    np.testing.assert_allclose(p[interior], E, rtol=1e-14)          # every interior vertex is on the medial axis


def test_cylinder_and_capsule_fields_with_margin():
    """WithMargin of both field tests through VerifyInvariantsOfThePressureFieldWithMargin (pressure_field_invariants.cc:
    15-59): boundary pressure -margin / (depth - margin) E < 0, interior slope depth / (depth - margin)."""
    E, margin, depth = 1.0e5, 0.12, 1.0
    for mesh, field in ((ho.VolumeMesh.cylinder_ma(1.0, 3.0, 0.25), lambda m, mg: ho.VolumeField.cylinder(m, 1.0, 3.0, E, mg)),
                        (ho.VolumeMesh.capsule(1.0, 3.0, 0.25), lambda m, mg: ho.VolumeField.capsule(m, 1.0, 3.0, E, mg))):
        p0, p1 = field(mesh, 0.0).values, field(mesh, margin).values
        boundary = _boundary_vertices(mesh.tets)
        want = -margin / (depth - margin) * E
        assert np.all(p1[boundary] < 0)
        np.testing.assert_allclose(p1[boundary], want, rtol=1e-14)
        interior = sorted(set(range(mesh.num_vertices)) - set(boundary))
        np.testing.assert_allclose((p1[interior] - want) / p0[interior], depth / (depth - margin), rtol=1e-14)


def test_capsule_field_min_max_boundary():
    """CheckMinMaxBoundaryValue of make_capsule_field_test.cc (:22-84) for its three capsules."""
    E = 1.0e5
    for radius, length in ((1.0, 3.0), (1.0, 2.0), (1.0, 0.5)):
        mesh = ho.VolumeMesh.capsule(radius, length, 0.25)
        p = ho.VolumeField.capsule(mesh, radius, length, E).values
        assert p.max() == E and p.min() == 0.0
        assert np.all(p[_boundary_vertices(mesh.tets)] == 0.0)
        assert p[0] == E and p[1] == E


# ---- the product's generators against the oracle's, bit for bit ---------------------------------------------------------

@pytest.mark.parametrize("radius,length,hint", [(1.0, 3.0, 0.5), (1.0, 2.0, 0.5), (1.0, 1.0, 0.5), (1.0, 1.0, 100.0),
                                                (0.05, 0.3, 0.02), (0.37, 0.11, 0.07), (0.2, 0.4 + 1e-15, 0.05)])
def test_product_cylinder_matches_the_oracle(radius, length, hint):
    m = db.Mesh.cylinder_volume_with_ma(radius, length, hint)
    o = ho.VolumeMesh.cylinder_ma(radius, length, hint)
    assert (m.num_vertices, m.num_elements, m.vertices_per_element) == (o.num_vertices, o.num_elements, 4)
    np.testing.assert_array_equal(m.vertices, o.verts)
    np.testing.assert_array_equal(m.elements, o.tets)
    for margin in (0.0, 0.01):
        np.testing.assert_array_equal(m.set_cylinder_pressure(radius, length, 3e4, margin).pressure,
                                      ho.VolumeField.cylinder(o, radius, length, 3e4, margin).values)
    s, so = m.to_surface(), o.to_surface()
    np.testing.assert_array_equal(s.vertices, so.verts)
    np.testing.assert_array_equal(s.elements, so.tris)


@pytest.mark.parametrize("radius,length,hint", [(1.0, 3.0, 0.5), (1.0, 0.5, 0.5), (1.0, 1.0, 100.0), (0.05, 0.3, 0.02),
                                                (0.37, 0.11, 0.07)])
def test_product_capsule_matches_the_oracle(radius, length, hint):
    m = db.Mesh.capsule_volume(radius, length, hint)
    o = ho.VolumeMesh.capsule(radius, length, hint)
    assert (m.num_vertices, m.num_elements, m.vertices_per_element) == (o.num_vertices, o.num_elements, 4)
    np.testing.assert_array_equal(m.vertices, o.verts)
    np.testing.assert_array_equal(m.elements, o.tets)
    for margin in (0.0, 0.01):
        np.testing.assert_array_equal(m.set_capsule_pressure(radius, length, 3e4, margin).pressure,
                                      ho.VolumeField.capsule(o, radius, length, 3e4, margin).values)
    s, so = m.to_surface(), o.to_surface()
    np.testing.assert_array_equal(s.vertices, so.verts)
    np.testing.assert_array_equal(s.elements, so.tris)


def test_bad_parameters_are_rejected():
    with pytest.raises(db.HydroError):
        db.Mesh.cylinder_volume_with_ma(1.0, 1.0, 0.0)
    with pytest.raises(db.HydroError):
        db.Mesh.capsule_volume(-1.0, 1.0, 0.5)
    with pytest.raises(db.HydroError):
        db.Mesh.cylinder_volume_with_ma(1.0, 1.0, 0.5).set_cylinder_pressure(1.0, 1.0, 1e5, 0.6)   # length <= 2 margin
    with pytest.raises(db.HydroError):
        db.Mesh.box_volume_with_ma(1, 1, 1).set_capsule_pressure(1.0, 1.0, 1e5)   # not a capsule mesh


# ---- compliant Convex / Mesh: MakeConvexVolumeMesh, MakeVolumeMeshPressureField -----------------------------------------

# geometry/test/non_convex_mesh.vtk of the reference (6 points, 6 tetrahedra), restated.
NON_CONVEX_V = [[0.1, 0.1, 0.1], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.2, 0.2, 0.2]]
NON_CONVEX_T = [[0, 5, 2, 3], [0, 5, 3, 4], [0, 5, 4, 2], [0, 1, 4, 3], [0, 1, 3, 2], [0, 1, 2, 4]]


def test_volume_mesh_pressure_field_reference_cases():
    """MakeVolumeMeshPressureFieldTest (make_mesh_field_test.cc:15-75): vertex 0 carries the modulus, 1-5 are zero; WithMargin
    through the margin invariants with depth 0.1 + margin; NoInteriorVertex throws the reference's message."""
    E = 1e7
    mesh = ho.VolumeMesh.from_arrays(NON_CONVEX_V, NON_CONVEX_T)
    p = ho.VolumeField.from_mesh(mesh, E).values
    assert p[0] == E and np.all(p[1:] == 0.0)
    E, margin, depth = 1.0e5, 0.012, 0.1 + 0.012
    p0, p1 = ho.VolumeField.from_mesh(mesh, E).values, ho.VolumeField.from_mesh(mesh, E, margin).values
    want = -margin / (depth - margin) * E
    assert np.all(p1[1:] < 0)
    np.testing.assert_allclose(p1[1:], want, rtol=1e-14)
    np.testing.assert_allclose((p1[0] - want) / p0[0], depth / (depth - margin), rtol=1e-14)
    one_tet = ho.VolumeMesh.from_arrays([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 2, 3]])
    with pytest.raises(RuntimeError, match="lacks interior vertices"):
        ho.VolumeField.from_mesh(one_tet, E)
    with pytest.raises(db.HydroError, match="lacks interior vertices"):
        db.Mesh.from_arrays([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 2, 3]]).set_volume_mesh_pressure(E)


def test_distance_to_a_posed_triangle_reference_regions():
    """CalcDistanceToSurfaceMeshTest.SingleTriangle (calc_distance_to_surface_mesh_test.cc:60-180): the triangle (0,0,0),
    (1,0,0), (0,1,0) posed at rpy (1, 2, 3), p (4, 5, 6); points of the seven Voronoi regions below / in / above its
    plane against the analytic distances, 1e-14."""
    X_WT = ho.xform(ho.rpy_matrix(1, 2, 3), [4, 5, 6])
    R, t = X_WT[:9].reshape(3, 3).T, X_WT[9:]
    A, B, C = np.zeros(3), np.array([1.0, 0, 0]), np.array([0, 1.0, 0])
    tri = ho.TriMesh.from_arrays([R @ v + t for v in (A, B, C)], [[0, 1, 2]])

    def line(Q, P0, P1):
        d = (P1 - P0) / np.linalg.norm(P1 - P0)
        return np.linalg.norm(Q - (d @ (Q - P0) * d + P0))
    cases = [((0.25, 0.25), lambda Q: abs(Q[2])), ((0.25, -0.25), lambda Q: line(Q, A, B)),
             ((-0.25, -0.25), lambda Q: np.linalg.norm(Q - A)), ((-0.25, 0.25), lambda Q: line(Q, A, C)),
             ((-0.25, 1.25), lambda Q: np.linalg.norm(Q - C)), ((0.75, 0.75), lambda Q: line(Q, B, C)),
             ((1.25, -0.25), lambda Q: np.linalg.norm(Q - B))]
    for (x, y), want in cases:
        for sz in (-1, 0, 1):
            Q = np.array([x, y, sz * 0.5])
            assert abs(tri.distance_to(R @ Q + t) - want(Q)) <= 1e-14, (x, y, sz)


def test_convex_volume_mesh_reference_properties_and_product():
    """MakeConvexMeshTest.CoherentVertices (make_convex_mesh_test.cc:95-125) on a cube's and an octahedron's surface: the
    surface's vertices in order, then one interior vertex -- the centroid of the enclosed volume -- behind every face;
    one tetrahedron of positive volume per triangle; the product's generator and field equal the oracle's."""
    cube = ho.VolumeMesh.box_ma(2.0, 2.0, 2.0).to_surface()                       # the reference's quad_cube.obj is this cube
    shifted = ho.TriMesh.from_arrays(cube.verts + np.array([0.3, -0.2, 0.5]), cube.tris)
    for surf, centre in ((cube, np.zeros(3)), (shifted, np.array([0.3, -0.2, 0.5]))):
        vol = surf.to_convex_volume()
        assert (vol.num_vertices, vol.num_elements) == (surf.num_vertices + 1, surf.num_elements)
        np.testing.assert_array_equal(vol.verts[:-1], surf.verts)
        np.testing.assert_allclose(vol.verts[-1], centre, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(vol.tets[:, 0], surf.num_vertices)
        np.testing.assert_array_equal(vol.tets[:, 1:], surf.tris)
        assert np.all(_tet_volumes(vol.verts, vol.tets) > 0)
        assert abs(float(np.sum(_tet_volumes(vol.verts, vol.tets))) - 8.0) <= 1e-14
        m = db.Mesh.from_arrays(surf.verts, surf.tris).to_convex_volume()
        np.testing.assert_array_equal(m.vertices, vol.verts)
        np.testing.assert_array_equal(m.elements, vol.tets)
        field = ho.VolumeField.from_mesh(vol, 1e6)
        assert field.values[-1] == 1e6 and np.all(field.values[:-1] == 0.0)
        np.testing.assert_array_equal(m.set_volume_mesh_pressure(1e6).pressure, field.values)


def test_product_mesh_field_matches_the_oracle():
    """MakeVolumeMeshPressureField of the product against the oracle, bit for bit, on meshes WITH many interior vertices
    (a grid box, a dense-interior sphere, the reference's non-convex mesh), with and without margin."""
    for o in (ho.VolumeMesh.box_grid(0.4, 0.3, 0.2, 0.05), ho.VolumeMesh.sphere(0.3, 0.1, True),
              ho.VolumeMesh.from_arrays(NON_CONVEX_V, NON_CONVEX_T)):
        m = db.Mesh.from_arrays(o.verts, o.tets)
        for margin in (0.0, 0.004):
            want = ho.VolumeField.from_mesh(o, 2e5, margin).values
            np.testing.assert_array_equal(m.set_volume_mesh_pressure(2e5, margin).pressure, want)
        assert want.max() == 2e5
    # A grid box: the pressure of an interior grid vertex is its distance to the nearest face over the largest such distance.
    o = ho.VolumeMesh.box_grid(0.4, 0.4, 0.4, 0.1)
    p = ho.VolumeField.from_mesh(o, 1.0).values
    d = np.min(0.2 - np.abs(o.verts), axis=1)
    np.testing.assert_allclose(p, d / d.max(), rtol=0, atol=1e-14)
