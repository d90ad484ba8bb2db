"""The known answers of the reference's hydroelastic-representation tests
(geometry/proximity/test/hydroelastic_internal_test.cc: HydroelasticRigidGeometryTest :578-766,
HydroelasticCompliantGeometryTest :1059-1318) for the meshes and pressure fields MakeRigidRepresentation /
MakeCompliantRepresentation build (hydroelastic_internal.cc:326-508). The C++ mirror's MakeAsset
(include/drake_b200_proximity_engine.h) issues exactly these generator calls; here they run through the oracle's
restatement and the product's host generators side by side. No GPU."""
import math

import numpy as np
import pytest

import drake_b200 as db
from oracle import ho

EPS = np.finfo(float).eps
E = 1e8          # AddCompliantHydroelasticProperties(edge_length, 1e8, ...) of the fixture (:1000-1004)


class _Oracle:
    name = "oracle"
    sphere = staticmethod(lambda r, h, dense=False: ho.VolumeMesh.sphere(r, h, dense))
    ellipsoid = staticmethod(lambda a, b, c, h, dense=False: ho.VolumeMesh.ellipsoid(a, b, c, h, dense))
    box = staticmethod(lambda x, y, z: ho.VolumeMesh.box_ma(x, y, z))
    cylinder = staticmethod(lambda r, l, h: ho.VolumeMesh.cylinder_ma(r, l, h))
    capsule = staticmethod(lambda r, l, h: ho.VolumeMesh.capsule(r, l, h))
    p_sphere = staticmethod(lambda m, r: ho.VolumeField.sphere(m, r, E).values)
    p_ellipsoid = staticmethod(lambda m, a, b, c: ho.VolumeField.ellipsoid(m, a, b, c, E).values)
    p_box = staticmethod(lambda m, x, y, z: ho.VolumeField.box(m, x, y, z, E).values)
    p_cylinder = staticmethod(lambda m, r, l: ho.VolumeField.cylinder(m, r, l, E).values)
    p_capsule = staticmethod(lambda m, r, l: ho.VolumeField.capsule(m, r, l, E).values)
    V = staticmethod(lambda m: m.verts)
    n = staticmethod(lambda m: m.num_elements)


class _Product:
    name = "product"
    sphere = staticmethod(lambda r, h, dense=False: db.Mesh.sphere_volume(r, h, int(dense)))
    ellipsoid = staticmethod(lambda a, b, c, h, dense=False: db.Mesh.ellipsoid_volume(a, b, c, h, int(dense)))
    box = staticmethod(lambda x, y, z: db.Mesh.box_volume_with_ma(x, y, z))
    cylinder = staticmethod(lambda r, l, h: db.Mesh.cylinder_volume_with_ma(r, l, h))
    capsule = staticmethod(lambda r, l, h: db.Mesh.capsule_volume(r, l, h))
    p_sphere = staticmethod(lambda m, r: np.array(m.set_sphere_pressure(r, E).pressure))
    p_ellipsoid = staticmethod(lambda m, a, b, c: np.array(m.set_ellipsoid_pressure(a, b, c, E).pressure))
    p_box = staticmethod(lambda m, x, y, z: np.array(m.set_box_pressure(x, y, z, E).pressure))
    p_cylinder = staticmethod(lambda m, r, l: np.array(m.set_cylinder_pressure(r, l, E).pressure))
    p_capsule = staticmethod(lambda m, r, l: np.array(m.set_capsule_pressure(r, l, E).pressure))
    V = staticmethod(lambda m: m.vertices)
    n = staticmethod(lambda m: m.num_elements)


SIDES = pytest.mark.parametrize("G", [_Oracle, _Product], ids=lambda g: g.name)


# ---- MakeRigidRepresentation: the boundary surface of the shape's volume mesh ----------------------------------------

@SIDES
def test_rigid_sphere(G):
    # :578-592 -- Sphere(0.5), resolution hint 0.5: every vertex on the sphere to 1e-15.
    s = G.sphere(0.5, 0.5).to_surface()
    assert np.all(np.abs(np.linalg.norm(G.V(s), axis=1) - 0.5) <= 1e-15)


@SIDES
def test_rigid_box(G):
    # :597-621 -- a 0.5 cube: 8 corners first, then the 6 face centres.
    s = G.box(0.5, 0.5, 0.5).to_surface()
    V = G.V(s)
    assert len(V) == 14
    assert np.all(np.abs(np.linalg.norm(V[:8], axis=1) - math.sqrt(3) * 0.5 / 2) <= 1e-15)
    assert np.all(np.abs(np.linalg.norm(V[8:], axis=1) - 0.25) <= 1e-15)


@SIDES
def test_rigid_cylinder(G):
    # :632-658 -- Cylinder(1, 2), hint 3: the coarsest mesh, 8 vertices / 12 triangles inside the cylinder's bounds.
    s = G.cylinder(1.0, 2.0, 3.0).to_surface()
    V = G.V(s)
    assert (len(V), G.n(s)) == (8, 12)
    assert np.all(V[:, 0] ** 2 + V[:, 1] ** 2 <= 1.0 + 1e-15) and np.all(V[:, 2] ** 2 <= 1.0 + 1e-15)


@SIDES
def test_rigid_capsule(G):
    # :663-698 -- Capsule(1, 2), hint 3: 8 vertices / 12 triangles, all ON the capsule; a finer hint refines.
    s = G.capsule(1.0, 2.0, 3.0).to_surface()
    V = G.V(s)
    assert (len(V), G.n(s)) == (8, 12)
    z = np.clip(V[:, 2], -1.0, 1.0)
    assert np.all(np.abs(np.sqrt(V[:, 0] ** 2 + V[:, 1] ** 2 + (V[:, 2] - z) ** 2) - 1.0) <= 1e-15)
    fine = G.capsule(1.0, 2.0, 2.0 * math.pi / 6.0).to_surface()
    assert len(G.V(fine)) > len(V) and G.n(fine) > G.n(s)


@SIDES
def test_rigid_ellipsoid(G):
    # :703-728 -- Ellipsoid(0.5, 0.8, 0.3), hint 1.6: the octahedron, on the ellipsoid to 1e-15.
    a, b, c = 0.5, 0.8, 0.3
    s = G.ellipsoid(a, b, c, 2.0 * b).to_surface()
    V = G.V(s)
    assert (len(V), G.n(s)) == (6, 8)
    assert np.all(np.abs((V[:, 0] / a) ** 2 + (V[:, 1] / b) ** 2 + (V[:, 2] / c) ** 2 - 1.0) <= 1e-15)


# ---- MakeCompliantRepresentation: volume mesh + pressure field ------------------------------------------------------------

@SIDES
def test_compliant_sphere(G):
    # :1059-1170
    R = 0.5
    s1, s2 = G.sphere(R, R), G.sphere(R, R / 2)
    assert G.n(s1) < G.n(s2)
    d = np.concatenate([np.linalg.norm(G.V(s1), axis=1), np.linalg.norm(G.V(s2), axis=1)])
    assert np.all(d <= R) and abs(d.max() - R) <= 1e-15
    p = G.p_sphere(s1, R)
    assert np.all(np.abs(p - E * (1.0 - np.linalg.norm(G.V(s1), axis=1) / R)) <= EPS * E)
    assert G.n(s1) * 4 == G.n(s2)                      # sparse: only the surface refines (:1117-1123)
    assert G.n(s1) < G.n(G.sphere(R, R, True))         # dense interior (:1125-1135)
    assert G.n(s1) == G.n(G.sphere(R, R, False))       # the default IS the sparse strategy (:1137-1147)


@SIDES
def test_compliant_box(G):
    # :1172-1187 -- Box(0.2, 0.4, 0.8): MakeBoxVolumeMeshWithMaAndSymmetricTriangles, 18 vertices.
    m = G.box(0.2, 0.4, 0.8)
    p = G.p_box(m, 0.2, 0.4, 0.8)
    assert len(G.V(m)) == 18 == len(p)
    assert np.all(p >= 0) and np.all(p <= E)


@SIDES
def test_compliant_cylinder(G):
    # :1192-1212 -- Cylinder(1, 2), hint 3: 9 vertices.
    m = G.cylinder(1.0, 2.0, 3.0)
    p = G.p_cylinder(m, 1.0, 2.0)
    assert len(G.V(m)) == 9 == len(p)
    assert np.all(p >= 0) and np.all(p <= E)


@SIDES
def test_compliant_capsule(G):
    # :1217-1239 -- Capsule(1, 2), hint 3: 10 vertices.
    m = G.capsule(1.0, 2.0, 3.0)
    p = G.p_capsule(m, 1.0, 2.0)
    assert len(G.V(m)) == 10 == len(p)
    assert np.all(p >= 0) and np.all(p <= E)


@SIDES
def test_compliant_ellipsoid(G):
    # :1244-1315 -- Ellipsoid(0.05, 0.08, 0.03): hint 0.16 gives 7 vertices; at 0.08 sparse (the default) < dense.
    a, b, c = 0.05, 0.08, 0.03
    m = G.ellipsoid(a, b, c, 0.16)
    p = G.p_ellipsoid(m, a, b, c)
    assert len(G.V(m)) == 7 == len(p)
    assert np.all(p >= 0) and np.all(p <= E)
    assert G.n(G.ellipsoid(a, b, c, 0.08)) == G.n(G.ellipsoid(a, b, c, 0.08, False)) < G.n(G.ellipsoid(a, b, c, 0.08, True))


def test_bad_tessellation_strategy_hints_and_sizes_fail_loudly():
    # BadResolutionHint / the "dense" string case (:978-988, :1149-1155, :1410-1422) and the shape constructors
    # (shape_specification.cc): an error, never a silent default.
    for make in (lambda: db.Mesh.sphere_volume(0.5, 0.5, 7), lambda: db.Mesh.ellipsoid_volume(0.05, 0.08, 0.03, 0.08, -1),
                 lambda: db.Mesh.sphere_volume(0.5, -0.2), lambda: db.Mesh.ellipsoid_volume(0.05, 0.08, 0.03, -0.2),
                 lambda: db.Mesh.cylinder_volume_with_ma(1.0, 2.0, -0.2), lambda: db.Mesh.capsule_volume(1.0, 2.0, -0.2),
                 lambda: db.Mesh.box_volume_with_ma(1.0, -1.0, 1.0), lambda: db.Mesh.box_volume(1.0, 0.0, 1.0, 0.5),
                 lambda: db.Mesh.ellipsoid_volume(0.05, -0.08, 0.03, 0.08),
                 # BadElasticModulus (:1424-1439) and the fields' DRAKE_DEMANDs (radius > margin, ...).
                 lambda: db.Mesh.sphere_volume(0.5, 0.5).set_sphere_pressure(0.5, -1e5),
                 lambda: db.Mesh.sphere_volume(0.5, 0.5).set_sphere_pressure(0.5, 1e5, 0.5),
                 lambda: db.Mesh.box_volume_with_ma(0.2, 0.4, 0.8).set_box_pressure(0.2, 0.4, 0.8, 1e5, 0.1)):
        with pytest.raises(db.HydroError):
            make()


# ---- HydroelasticVanished (:418-461): is_gradient_field_degenerate of the compliant primitives --------------------------

def _product_field_is_degenerate(mesh, pressure):
    """The product's registration-time derivation (derive_tet, drake_b200/csrc/hydro_derive.cuh, compiled for the host by
    tests/test_plane_math_cpu.py's harness) over every tetrahedron: what dbh_register_mesh answers
    DBH_ERR_DEGENERATE_FIELD on, and the C++ mirror turns into is_vanished() for a primitive."""
    import ctypes as C

    from tests.test_plane_math_cpu import harness
    L = harness()
    dp = C.POINTER(C.c_double)
    L.phc_derive_tet.argtypes = [dp] * 5
    V, T = np.ascontiguousarray(mesh.vertices, float), mesh.elements
    g, m, n = np.zeros(3), np.zeros(1), np.zeros(12)
    for t in T:
        v, u = np.ascontiguousarray(V[t]), np.ascontiguousarray(pressure[t], float)
        if not L.phc_derive_tet(v.ctypes.data_as(dp), u.ctypes.data_as(dp), g.ctypes.data_as(dp), m.ctypes.data_as(dp),
                                n.ctypes.data_as(dp)):
            return True
    return False


@pytest.mark.parametrize("shape,size,vanished", [
    ("sphere", 0.5, False), ("sphere", 1e-6, True), ("box", 0.5, False), ("box", 1e-6, True),
    ("capsule", 0.5, False), ("capsule", 1e-6, True), ("cylinder", 0.5, False), ("cylinder", 1e-6, True),
    ("ellipsoid", 0.5, False), ("ellipsoid", 1e-6, True)])
def test_vanished_primitives(shape, size, vanished):
    """DoTestVanished: AddCompliantHydroelasticProperties(resolution_hint 1.0, modulus 1e8); a primitive of size 1e-6 has a
    pressure field without gradient (|signed volume| <= eps, volume_mesh.h:395-446) and vanishes, one of size 0.5 does not."""
    s, hint = size, 1.0
    if shape == "sphere":
        o, p = ho.VolumeMesh.sphere(s, hint), db.Mesh.sphere_volume(s, hint)
        fo, fp = ho.VolumeField.sphere(o, s, E), p.set_sphere_pressure(s, E).pressure
    elif shape == "box":
        o, p = ho.VolumeMesh.box_ma(s, s, s), db.Mesh.box_volume_with_ma(s, s, s)
        fo, fp = ho.VolumeField.box(o, s, s, s, E), p.set_box_pressure(s, s, s, E).pressure
    elif shape == "capsule":
        o, p = ho.VolumeMesh.capsule(s, s, hint), db.Mesh.capsule_volume(s, s, hint)
        fo, fp = ho.VolumeField.capsule(o, s, s, E), p.set_capsule_pressure(s, s, E).pressure
    elif shape == "cylinder":
        o, p = ho.VolumeMesh.cylinder_ma(s, s, hint), db.Mesh.cylinder_volume_with_ma(s, s, hint)
        fo, fp = ho.VolumeField.cylinder(o, s, s, E), p.set_cylinder_pressure(s, s, E).pressure
    else:
        o, p = ho.VolumeMesh.ellipsoid(s, s, s, hint), db.Mesh.ellipsoid_volume(s, s, s, hint)
        fo, fp = ho.VolumeField.ellipsoid(o, s, s, s, E), p.set_ellipsoid_pressure(s, s, s, E).pressure
    assert bool(fo.arrays()[3]) == vanished
    assert _product_field_is_degenerate(p, np.array(fp)) == vanished
