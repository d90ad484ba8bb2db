"""The BASELINE configs built entirely on the oracle's side (test infrastructure; bench.py --impl reference).

drake_b200.workloads describes every config through a mesh factory (`Mesh=`): the product passes its own generators,
the reference arm of bench.py passes RefMesh below, whose constructors are the oracle's restatement of
MakeSphereVolumeMesh / MakeEllipsoidVolumeMesh / MakeBoxVolumeMesh* / ConvertVolumeToSurfaceMesh / Make*PressureField
(oracle/ho_mesh.h). The two generator families are bit-identical (tests/test_generators_cpu.py), so both arms see the
same meshes and -- the poses being a pure function of (config, scene) -- the same scenes; but this arm never maps the
product library (drake_b200/lib/libdrake_b200_hydro.so): only ctypes structure definitions are imported from there.
"""
from __future__ import annotations

import numpy as np

from . import ho


class RefMesh:
    """Same constructors and attributes as drake_b200.engine.Mesh, over oracle objects."""

    def __init__(self, mesh, vertices_per_element: int):
        self.mesh = mesh
        self.field = None
        self.vertices_per_element = vertices_per_element
        self.num_elements = mesh.num_elements
        self.num_vertices = mesh.num_vertices

    @staticmethod
    def sphere_volume(radius, resolution_hint, strategy=0):
        return RefMesh(ho.VolumeMesh.sphere(radius, resolution_hint, dense=bool(strategy)), 4)

    @staticmethod
    def ellipsoid_volume(a, b, c, resolution_hint, strategy=0):
        return RefMesh(ho.VolumeMesh.ellipsoid(a, b, c, resolution_hint, dense=bool(strategy)), 4)

    @staticmethod
    def box_volume_with_ma(x, y, z):
        return RefMesh(ho.VolumeMesh.box_ma(x, y, z), 4)

    @staticmethod
    def box_volume(x, y, z, resolution_hint):
        return RefMesh(ho.VolumeMesh.box_grid(x, y, z, resolution_hint), 4)

    @staticmethod
    def sphere_surface(radius, resolution_hint):
        return RefMesh(ho.TriMesh.sphere(radius, resolution_hint), 3)

    @staticmethod
    def ellipsoid_surface(a, b, c, resolution_hint):
        return RefMesh(ho.TriMesh.ellipsoid(a, b, c, resolution_hint), 3)

    @staticmethod
    def box_surface(x, y, z):
        return RefMesh(ho.TriMesh.box(x, y, z), 3)

    def set_sphere_pressure(self, radius, modulus, margin=0.0):
        self.field = ho.VolumeField.sphere(self.mesh, radius, modulus, margin)
        return self

    def set_ellipsoid_pressure(self, a, b, c, modulus, margin=0.0):
        self.field = ho.VolumeField.ellipsoid(self.mesh, a, b, c, modulus, margin)
        return self

    def set_box_pressure(self, x, y, z, modulus, margin=0.0):
        self.field = ho.VolumeField.box(self.mesh, x, y, z, modulus, margin)
        return self


def scene_for(workload):
    """ho.Scene of a workload built with Mesh=RefMesh (one template scene; poses differ per scene). Returns
    (scene, order): `order` sorts the template geometries by ascending id, the order the oracle wants poses in."""
    sc = ho.Scene()
    for g, mi in enumerate(workload.geom_mesh):
        m = workload.meshes[mi]
        gid, anchored = workload.geom_id[g], bool(workload.geom_anchored[g])
        if m.vertices_per_element == 0:   # drake_b200.workloads.HalfSpaceShape
            sc.add_half_space(gid, m.compliant, m.pressure_scale, m.margin, anchored)
        elif m.vertices_per_element == 4:
            sc.add_compliant(gid, m.mesh, m.field, anchored)
        else:
            sc.add_rigid(gid, m.mesh, anchored)
    if workload.excluded_local is not None:
        for a, b in np.asarray(workload.excluded_local).reshape(-1, 2):
            sc.exclude(int(workload.geom_id[a]), int(workload.geom_id[b]))
    if not getattr(workload, "normal_filter", True):
        sc.set_normal_filter(False)
    return sc, np.argsort(np.asarray(workload.geom_id))
