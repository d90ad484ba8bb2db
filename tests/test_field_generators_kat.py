"""The reference's own tests of the a17 generators that feed configs 1-5 (sphere / ellipsoid / box pressure fields, the
ellipsoid mesh, the boundary surface of a volume mesh), at the reference's parameters, run through BOTH the oracle's
restatement and the product's host-side generators (drake_b200/csrc/hydro_shapes.cu behind dbh_mesh_*):

  geometry/proximity/test/make_sphere_field_test.cc:21-100
  geometry/proximity/test/make_ellipsoid_field_test.cc:21-137
  geometry/proximity/test/make_box_field_test.cc:20-126
  geometry/proximity/test/make_ellipsoid_mesh_test.cc:15-61
  geometry/proximity/test/volume_to_surface_mesh_test.cc:19-128
  geometry/proximity/test/pressure_field_invariants.cc:15-59

No GPU: the generators are host code on both sides."""
import numpy as np
import pytest

import drake_b200 as db
from oracle import ho

EPS = np.finfo(float).eps
E = 1.0e5


def _boundary_faces(T):
    """IdentifyBoundaryFaces (volume_to_surface_mesh.cc:31-110) restated for the test: the faces owned by one
    tetrahedron only, as a set of sorted triples."""
    count = {}
    for t in T:
        for f in ((1, 2, 3), (3, 2, 0), (1, 3, 0), (2, 1, 0)):
            key = tuple(sorted(int(t[k]) for k in f))
            count[key] = count.get(key, 0) + 1
    return {k for k, c in count.items() if c == 1}


def _boundary_vertices(T):
    return sorted({v for f in _boundary_faces(T) for v in f})


class _Oracle:
    name = "oracle"
    sphere = staticmethod(lambda r, h, dense: ho.VolumeMesh.sphere(r, h, dense))
    ellipsoid = staticmethod(lambda a, b, c, h, dense: ho.VolumeMesh.ellipsoid(a, b, c, h, dense))
    box_grid = staticmethod(lambda x, y, z, h: ho.VolumeMesh.box_grid(x, y, z, h))
    box_ma = staticmethod(lambda x, y, z: ho.VolumeMesh.box_ma(x, y, z))
    from_arrays = staticmethod(lambda v, t: ho.VolumeMesh.from_arrays(v, t))
    sphere_field = staticmethod(lambda m, r, e, mg=0.0: ho.VolumeField.sphere(m, r, e, mg).values)
    ellipsoid_field = staticmethod(lambda m, a, b, c, e, mg=0.0: ho.VolumeField.ellipsoid(m, a, b, c, e, mg).values)
    box_field = staticmethod(lambda m, x, y, z, e, mg=0.0: ho.VolumeField.box(m, x, y, z, e, mg).values)
    ellipsoid_surface = staticmethod(lambda a, b, c, h: ho.TriMesh.ellipsoid(a, b, c, h))
    verts = staticmethod(lambda m: m.verts)
    tets = staticmethod(lambda m: m.tets)
    tris = staticmethod(lambda m: m.tris)


class _Product:
    name = "product"
    sphere = staticmethod(lambda r, h, dense: db.Mesh.sphere_volume(r, h, int(dense)))
    ellipsoid = staticmethod(lambda a, b, c, h, dense: db.Mesh.ellipsoid_volume(a, b, c, h, int(dense)))
    box_grid = staticmethod(lambda x, y, z, h: db.Mesh.box_volume(x, y, z, h))
    box_ma = staticmethod(lambda x, y, z: db.Mesh.box_volume_with_ma(x, y, z))
    from_arrays = staticmethod(lambda v, t: db.Mesh.from_arrays(np.asarray(v, float), np.asarray(t, np.int32)))
    sphere_field = staticmethod(lambda m, r, e, mg=0.0: np.array(m.set_sphere_pressure(r, e, mg).pressure))
    ellipsoid_field = staticmethod(lambda m, a, b, c, e, mg=0.0: np.array(m.set_ellipsoid_pressure(a, b, c, e, mg).pressure))
    box_field = staticmethod(lambda m, x, y, z, e, mg=0.0: np.array(m.set_box_pressure(x, y, z, e, mg).pressure))
    ellipsoid_surface = staticmethod(lambda a, b, c, h: db.Mesh.ellipsoid_surface(a, b, c, h))
    verts = staticmethod(lambda m: m.vertices)
    tets = staticmethod(lambda m: m.elements)
    tris = staticmethod(lambda m: m.elements)


SIDES = pytest.mark.parametrize("G", [_Oracle, _Product], ids=lambda g: g.name)


def _check_min_max_boundary_exact(G, mesh, p, need_centre):
    """CheckMinMaxBoundaryValue of make_sphere_field_test.cc:21-63 / the loops of make_box_field_test.cc:30-54: exact
    comparisons throughout."""
    assert np.all(p <= E) and np.all(p >= 0.0)
    assert p.min() == 0.0 and p.max() == E
    assert np.all(p[_boundary_vertices(G.tets(mesh))] == 0.0)
    centre = np.nonzero(np.all(G.verts(mesh) == 0.0, axis=1))[0]
    if need_centre:
        assert len(centre) >= 1
        assert p[centre[0]] == E


def _verify_margin_invariants(G, mesh, p0, p1, margin, depth, rel_tol=EPS):
    """VerifyInvariantsOfThePressureFieldWithMargin (pressure_field_invariants.cc:15-59)."""
    boundary = _boundary_vertices(G.tets(mesh))
    want = -margin / (depth - margin) * E
    assert np.all(p1[boundary] < 0)
    assert np.all(np.abs(p1[boundary] - want) <= rel_tol * abs(want))
    interior = sorted(set(range(len(p0))) - set(boundary))
    assert interior
    slope = (p1[interior] - want) / p0[interior]
    assert np.all(np.abs(slope - depth / (depth - margin)) <= rel_tol)


@SIDES
def test_make_sphere_pressure_field(G):
    # make_sphere_field_test.cc:65-82
    mesh = G.sphere(2.0, 0.25, True)
    assert len(G.verts(mesh)) > 7 and len(G.tets(mesh)) > 8
    _check_min_max_boundary_exact(G, mesh, G.sphere_field(mesh, 2.0, E), need_centre=True)


@SIDES
def test_make_sphere_pressure_field_with_margin(G):
    # make_sphere_field_test.cc:84-98
    mesh = G.sphere(2.0, 0.5, False)
    _verify_margin_invariants(G, mesh, G.sphere_field(mesh, 2.0, E), G.sphere_field(mesh, 2.0, E, 0.12), 0.12, 2.0)


@SIDES
def test_make_ellipsoid_pressure_field(G):
    # make_ellipsoid_field_test.cc:68-87 through its CheckMinMaxBoundaryValue (:21-66): 1e-14 E on the maximum.
    mesh = G.ellipsoid(0.05, 0.08, 0.03, 0.04, True)
    assert len(G.verts(mesh)) > 7 and len(G.tets(mesh)) > 8
    p = G.ellipsoid_field(mesh, 0.05, 0.08, 0.03, E)
    tol = 1e-14 * E
    assert np.all(p <= E + tol) and np.all(p >= 0.0)
    assert p.min() == 0.0 and abs(p.max() - E) <= tol
    centre = np.nonzero(np.all(G.verts(mesh) == 0.0, axis=1))[0]
    assert len(centre) >= 1 and abs(p[centre[0]] - p.max()) <= tol


@SIDES
def test_make_ellipsoid_pressure_field_with_margin(G):
    # make_ellipsoid_field_test.cc:89-137: at the six axis end points the pressure is -margin / (radius_k - margin) E.
    margin, radii = 0.012, np.array([0.05, 0.08, 0.03])
    mesh = G.ellipsoid(*radii, 0.08, False)
    p = G.ellipsoid_field(mesh, *radii, E, margin)
    tested = 0
    for i, v in enumerate(G.verts(mesh)):
        for k in range(3):
            if abs(v[k]) == radii[k] and v[(k + 1) % 3] == 0 and v[(k + 2) % 3] == 0:
                want = -margin / (radii[k] - margin) * E
                assert abs(p[i] - want) <= 2.0 * EPS * abs(want)
                tested += 1
    assert tested == 6


@SIDES
def test_make_box_pressure_field(G):
    # make_box_field_test.cc:20-55
    mesh = G.box_grid(1.0, 2.0, 3.0, 0.25)
    _check_min_max_boundary_exact(G, mesh, G.box_field(mesh, 1.0, 2.0, 3.0, E), need_centre=False)


@SIDES
def test_make_box_pressure_field_with_margin(G):
    # make_box_field_test.cc:57-71. The reference's test meshes with MakeBoxVolumeMeshWithMa; box_ma is the variant the
    # compliant Box representation uses (MakeBoxVolumeMeshWithMaAndSymmetricTriangles, hydroelastic_internal.cc:451-452):
    # the same medial axis plus the six face centres, so the same invariants must hold.
    mesh = G.box_ma(1.0, 2.0, 3.0)
    _verify_margin_invariants(G, mesh, G.box_field(mesh, 1.0, 2.0, 3.0, E), G.box_field(mesh, 1.0, 2.0, 3.0, E, 0.12),
                              0.12, 0.5)


@SIDES
def test_make_box_pressure_field_in_mesh_with_medial_axis(G):
    # make_box_field_test.cc:73-126: as above, and no tetrahedron of constant pressure (on the representation's mesh,
    # see the note in the test above).
    mesh = G.box_ma(1.0, 2.0, 3.0)
    p = G.box_field(mesh, 1.0, 2.0, 3.0, E)
    _check_min_max_boundary_exact(G, mesh, p, need_centre=False)
    pt = p[G.tets(mesh)]
    same = np.all(np.abs(pt[:, 1:] - pt[:, :1]) <= E * EPS, axis=1)
    assert not same.any()


@SIDES
def test_make_ellipsoid_volume_and_surface_mesh_counts(G):
    # make_ellipsoid_mesh_test.cc:15-61
    coarse = G.ellipsoid(0.05, 0.08, 0.03, 0.16, True)
    assert (len(G.verts(coarse)), len(G.tets(coarse))) == (7, 8)
    assert len(G.tets(G.ellipsoid(0.05, 0.08, 0.03, 0.08, True))) == 64
    s = G.ellipsoid_surface(0.05, 0.08, 0.03, 0.16)
    assert (len(G.verts(s)), len(G.tris(s))) == (6, 8)
    assert len(G.tris(G.ellipsoid_surface(0.05, 0.08, 0.03, 0.08))) == 32


@SIDES
def test_identify_boundary_faces_order(G):
    """VolumeToSurfaceMeshTest.IdentifyBoundaryFaces (volume_to_surface_mesh_test.cc:19-58): two tetrahedra sharing a
    face; the reference expects the six boundary triangles in exactly this order with exactly these windings."""
    verts = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1]]
    mesh = G.from_arrays(verts, [[0, 1, 2, 3], [0, 2, 1, 4]])
    s = mesh.to_surface()
    # All five vertices are on the boundary, so the surface keeps the volume mesh's numbering.
    np.testing.assert_array_equal(G.verts(s), np.array(verts, float))
    assert [list(map(int, t)) for t in G.tris(s)] == [[1, 3, 0], [4, 1, 0], [3, 2, 0], [2, 4, 0], [1, 2, 3], [2, 1, 4]]


@SIDES
def test_volume_to_surface_mesh(G):
    """TestVolumeToSurfaceMesh (volume_to_surface_mesh_test.cc:78-118): the surface's vertices are the volume's boundary
    vertices, none duplicated, and every face normal points outwards."""
    volume = G.box_grid(0.2, 0.4, 0.6, 0.1)
    V = G.verts(volume)
    on_boundary = (np.abs(V[:, 0]) == 0.1) | (np.abs(V[:, 1]) == 0.2) | (np.abs(V[:, 2]) == 0.3)
    want = {tuple(v) for v in V[on_boundary]}
    s = volume.to_surface()
    SV, ST = G.verts(s), G.tris(s)
    got = {tuple(v) for v in SV}
    assert got == want and len(got) == len(SV)
    a, b, c = SV[ST[:, 0]], SV[ST[:, 1]], SV[ST[:, 2]]
    n = np.cross(b - a, c - a)
    assert np.all(np.einsum("ij,ij->i", n, a) > 0)
    assert {tuple(sorted(map(int, t))) for t in ST} == {
        tuple(sorted(int(np.nonzero(np.all(SV == V[v], axis=1))[0][0]) for v in f)) for f in _boundary_faces(G.tets(volume))}
