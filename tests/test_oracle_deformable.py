"""Pins oracle/ho_deformable.py (deformable x rigid contact data) with the reference's own fixtures:
geometry/proximity/test/deformable_mesh_intersection_test.cc (NoContact :33-104, OnePolygon :106-178,
OnlyRelativePoseMatters :180-262) and the contract of RotationMatrix::MakeFromOneUnitVector
(math/rotation_matrix.h, math/test/rotation_matrix_test.cc: column `axis` is the given unit vector, the matrix is a
right-handed orthonormal basis, the element of the other two columns at the index of u's smallest component is 0 / the
largest of its column). CPU only."""
import numpy as np

from oracle import ho
from oracle import ho_deformable as hd

EX, EY, EZ, O = [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0.0, 0, 0]


def test_one_polygon_fixture():
    """deformable_mesh_intersection_test.cc:106-178: a single deformable tetrahedron whose bottom face is cut by a
    single rigid tetrahedron with constant pressure 123 posed at (1/6, 1/6, -1/2): one contact point at (1/3, 1/3, 0)
    with barycentric coordinates (1/3, 1/3, 1/3) in the bottom triangle, whose volume-mesh vertices are {2, 1, 0}."""
    deformable_volume = ho.VolumeMesh.from_arrays([O, EX, EY, EZ], [[0, 1, 2, 3]])
    surface = deformable_volume.to_surface()
    rigid = ho.VolumeMesh.from_arrays([O, EX, EY, EZ], [[0, 1, 2, 3]])
    pressure = ho.VolumeField.from_values(rigid, [123.0] * 4)
    X_WR = ho.xform(None, [1.0 / 6.0, 1.0 / 6.0, -0.5])
    c = hd.deformable_rigid_contact(rigid, pressure, surface, X_WR)
    assert c is not None and len(c.pressures) == 1
    assert abs(c.pressures[0] - 123.0) <= np.finfo(float).eps * 123
    assert np.allclose(c.contact_points_W[0], [1.0 / 3.0, 1.0 / 3.0, 0.0], rtol=0, atol=1e-14)
    assert np.allclose(c.barycentric_centroids[0], [1.0 / 3, 1.0 / 3, 1.0 / 3], rtol=0, atol=1e-14)
    # All four vertices of a single tetrahedron are boundary vertices: surface index = volume index.
    tri_vol, participating = hd.participating_vertices(c.triangle_index, surface.tris, np.arange(4))
    assert list(tri_vol[0]) == [2, 1, 0] and list(participating) == [0, 1, 2]
    # Rigid into deformable: the bottom face's outward normal is -z, so nhat = +z; R_WC's z axis is -nhat.
    assert np.allclose(c.nhats_W[0], [0, 0, 1.0], atol=1e-15)
    assert np.allclose(c.R_WCs[0][:, 2], [0, 0, -1.0], atol=1e-15)
    # The constant field has a zero gradient: what the reference stores is the point transform of it, p_WR.
    assert np.allclose(c.pressure_gradients_W[0], [1.0 / 6.0, 1.0 / 6.0, -0.5], atol=1e-15)


def _octahedron_and_cube():
    # MakeSphereVolumeMesh(unit sphere, 10.0, kDenseInteriorVertices): the coarsest sphere, an octahedron of 8 tets.
    deformable_volume = ho.VolumeMesh.sphere(1.0, 10.0, dense=True)
    assert len(deformable_volume.tets) == 8
    rigid = ho.VolumeMesh.box_ma(2.0, 2.0, 2.0)
    return deformable_volume, rigid, ho.VolumeField.box(rigid, 2.0, 2.0, 2.0, 1e5)


def test_no_contact_fixture():
    """:33-104: the cube sits just outside the octahedron's face x + y + z = 1 (inside a tetrahedron's box): no
    contact surface."""
    deformable_volume, rigid, pressure = _octahedron_and_cube()
    t = 4.0 / 3 + 1e-10
    assert hd.deformable_rigid_contact(rigid, pressure, deformable_volume.to_surface(), ho.xform(None, [t, t, t])) is None
    # ... and a hair closer it does touch.
    t = 4.0 / 3 - 1e-3
    assert hd.deformable_rigid_contact(rigid, pressure, deformable_volume.to_surface(), ho.xform(None, [t, t, t])) is not None


def test_only_the_relative_pose_matters():
    """:180-262: the deformable vertices and the rigid pose moved by the same transform: as many contact points, the
    same pressures, the same barycentric coordinates; pressures are non-negative and not all zero."""
    deformable_volume, rigid, pressure = _octahedron_and_cube()
    surface = deformable_volume.to_surface()
    X_WR = ho.xform(None, [1.5, 0.0, 0.0])
    a = hd.deformable_rigid_contact(rigid, pressure, surface, X_WR)
    assert a is not None and len(a.pressures) > 0
    shift = np.array([1.2, 3.4, 5.6])
    surface.update_vertices(surface.verts + shift)
    b = hd.deformable_rigid_contact(rigid, pressure, surface, ho.xform(None, [1.5 + 1.2, 3.4, 5.6]))
    assert len(b.pressures) == len(a.pressures)
    order_a = np.lexsort(a.contact_points_W.round(9).T)
    order_b = np.lexsort((b.contact_points_W - shift).round(9).T)
    assert np.allclose(a.pressures[order_a], b.pressures[order_b], rtol=1e-9, atol=1e-6)
    assert np.allclose(a.barycentric_centroids[order_a], b.barycentric_centroids[order_b], atol=1e-9)
    assert a.pressures.min() >= 0.0 and a.pressures.max() > 0.0


def test_make_from_one_unit_vector():
    rng = np.random.default_rng(5)
    cases = [np.array(EX), np.array(EY), np.array(EZ), -np.array(EZ), np.array([1, 1, 1]) / np.sqrt(3.0)]
    cases += [v / np.linalg.norm(v) for v in rng.normal(size=(50, 3))]
    for u in cases:
        for axis in range(3):
            R = hd.make_from_one_unit_vector(u, axis)
            assert np.array_equal(R[:, axis], u)
            assert np.allclose(R.T @ R, np.eye(3), atol=4e-16 * 8) and abs(np.linalg.det(R) - 1) < 1e-14
            i = int(np.argmin(np.abs(u)))
            v, w = R[:, (axis + 1) % 3], R[:, (axis + 2) % 3]
            assert v[i] == 0 and w[i] == np.max(w)
    # u = +x as the z axis: the smallest component is y (the first of the two zeros), so v = -z, w = +y.
    assert np.array_equal(hd.make_from_one_unit_vector(EX, 2), np.array([[0, 0, 1.0], [0, 1.0, 0], [-1.0, 0, 0]]))


def test_calc_barycentric():
    v0, v1, v2 = np.array([0.5, 0, 0]), np.array([0, 2.0, 0.1]), np.array([-1.0, 0.3, 3.0])
    rng = np.random.default_rng(6)
    n = np.cross(v1 - v0, v2 - v0)
    for _ in range(50):
        b = rng.uniform(-0.5, 1.5, 3)
        b /= b.sum()
        p = b[0] * v0 + b[1] * v1 + b[2] * v2
        assert np.allclose(hd.calc_barycentric(p, v0, v1, v2), b, atol=1e-13)
        # A point off the plane gets the coordinates of its projection (least squares).
        assert np.allclose(hd.calc_barycentric(p + 0.7 * n, v0, v1, v2), b, atol=1e-12)


def _two_tets(shift_z):
    """deformable_field_intersection_test.cc:27-87: deform0 above the X-Y plane (signed distance 0 on the base, -1 at the
    apex), deform1 below it with the element written {1, 0, 2, 3}, translated along z."""
    d0 = ho.VolumeMesh.from_arrays([O, EX, EY, EZ], [[0, 1, 2, 3]])
    v1 = np.array([O, EX, EY, [0, 0, -1.0]]) + [0, 0, shift_z]
    d1 = ho.VolumeMesh.from_arrays(v1, [[1, 0, 2, 3]])
    return d0, d1


def test_deformable_deformable_fixture():
    """:99-112 (translated out of contact: nothing) and :114-173: translated 0.2 into contact, the call
    AddDeformableDeformableContactSurface(sdf1, deform1, id1, sdf0, deform0, id0) gives one contact point whose normal
    is +z exactly, whose signed distance is sdf1 evaluated there and whose barycentric coordinates are what the two
    meshes compute for the point."""
    sdf = [0.0, 0.0, 0.0, -1.0]
    d0, d1 = _two_tets(-0.2)
    assert hd.deformable_deformable_contact(d1, sdf, d0, sdf) is None
    d0, d1 = _two_tets(0.2)
    c = hd.deformable_deformable_contact(d1, sdf, d0, sdf)
    assert c is not None and len(c.signed_distances) == 1
    assert np.array_equal(c.nhats_W[0], [0.0, 0.0, 1.0])
    p = c.contact_points_W[0]
    assert np.array_equal(p, c.surface.centroids[0]) and abs(p[2] - 0.1) < 1e-15
    # sdf1 = z - 0.2 on deform1 (0 on its base at z = 0.2, -1 at its apex at z = -0.8).
    assert abs(c.signed_distances[0] - (p[2] - 0.2)) < 1e-15
    v0, t0 = d0.arrays()[:2]
    v1, t1 = d1.arrays()[:2]
    b_A, b_B = c.barycentric1[0], c.barycentric0[0]      # A = deform0 (smaller id) is the call's SECOND mesh
    assert np.allclose(b_A @ v0[t0[0]], p, atol=1e-15) and abs(b_A.sum() - 1) < 1e-15
    assert np.allclose(b_B @ v1[t1[0]], p, atol=1e-15) and abs(b_B.sum() - 1) < 1e-15
    assert list(t0[0]) == [0, 1, 2, 3] and list(t1[0]) == [1, 0, 2, 3]      # contact vertex indexes A / B (:162-170)
    assert np.allclose(c.R_WCs[0][:, 2], [0, 0, -1.0])
