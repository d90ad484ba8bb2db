"""The product-side mesh / field generators (drake_b200/csrc/hydro_shapes.cu) against the oracle's
independent restatement of the reference generators: identical arrays, bit for bit."""
import numpy as np
import pytest

import drake_b200 as db
from oracle import ho


def same_volume(m: db.Mesh, o: ho.VolumeMesh):
    assert (m.num_vertices, m.num_elements, m.vertices_per_element) == (o.num_vertices, o.num_elements, 4)
    np.testing.assert_array_equal(m.vertices, o.verts)
    np.testing.assert_array_equal(m.elements, o.tets)


def same_surface(m: db.Mesh, o: ho.TriMesh):
    assert (m.num_vertices, m.num_elements, m.vertices_per_element) == (o.num_vertices, o.num_elements, 3)
    np.testing.assert_array_equal(m.vertices, o.verts)
    np.testing.assert_array_equal(m.elements, o.tris)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("radius,hint", [(1.0, 3.0), (0.25, 0.25), (0.25, 0.1), (3.0, 1.0), (0.05, 0.015)])
def test_sphere_volume_and_field(radius, hint, dense):
    m = db.Mesh.sphere_volume(radius, hint, int(dense))
    o = ho.VolumeMesh.sphere(radius, hint, dense)
    same_volume(m, o)
    np.testing.assert_array_equal(m.set_sphere_pressure(radius, 1e5).pressure,
                                  ho.VolumeField.sphere(o, radius, 1e5).values)
    np.testing.assert_array_equal(m.set_sphere_pressure(radius, 3e4, 0.001).pressure,
                                  ho.VolumeField.sphere(o, radius, 3e4, 0.001).values)
    same_surface(m.to_surface(), o.to_surface())


def test_dense_sphere_level_4():
    same_volume(db.Mesh.sphere_volume(3.0, 0.5, 1), ho.VolumeMesh.sphere(3.0, 0.5, True))


@pytest.mark.parametrize("dense", [False, True])
def test_ellipsoid_volume_and_field(dense):
    for a, b, c, hint in [(3.01, 3.5, 4.0, 2.0), (0.02, 0.05, 0.03, 0.01), (0.1, 0.1, 0.01, 0.01)]:
        m = db.Mesh.ellipsoid_volume(a, b, c, hint, int(dense))
        o = ho.VolumeMesh.ellipsoid(a, b, c, hint, dense)
        same_volume(m, o)
        np.testing.assert_array_equal(m.set_ellipsoid_pressure(a, b, c, 1e5).pressure,
                                      ho.VolumeField.ellipsoid(o, a, b, c, 1e5).values)
    same_surface(db.Mesh.ellipsoid_surface(3.01, 3.5, 4.0, 1.0), ho.TriMesh.ellipsoid(3.01, 3.5, 4.0, 1.0))


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 2, 2), (2, 4, 2), (2, 2, 4), (2, 4, 6), (2, 6, 4), (4, 2, 6),
                                  (4, 6, 2), (6, 2, 4), (6, 4, 2), (0.3, 0.3, 0.05), (0.2, 0.2, 0.01)])
def test_box_with_medial_axis(dims):
    m = db.Mesh.box_volume_with_ma(*dims)
    o = ho.VolumeMesh.box_ma(*dims)
    same_volume(m, o)
    np.testing.assert_array_equal(m.set_box_pressure(*dims, 3e4).pressure, ho.VolumeField.box(o, *dims, 3e4).values)
    same_surface(db.Mesh.box_surface(*dims), ho.TriMesh.box(*dims))


def test_box_grid():
    for dims, hint in [((0.2, 0.4, 0.8), 0.1), ((0.07, 0.05, 0.09), 0.02)]:
        m = db.Mesh.box_volume(*dims, hint)
        o = ho.VolumeMesh.box_grid(*dims, hint)
        same_volume(m, o)
        np.testing.assert_array_equal(m.set_box_pressure(*dims, 1e5).pressure, ho.VolumeField.box(o, *dims, 1e5).values)


def test_bad_arguments_fail_loudly():
    with pytest.raises(db.HydroError):
        db.Mesh.sphere_volume(1.0, -1.0)
    with pytest.raises(db.HydroError):
        db.Mesh.sphere_volume(1.0, 0.5).set_sphere_pressure(1.0, -5.0)
