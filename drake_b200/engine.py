"""Python host mirror of the C-ABI in include/drake_b200_hydro.h.

Everything here is plumbing over libdrake_b200_hydro.so (ctypes): there is no
Python or CPU implementation of the contact query. Importing works anywhere
the shared library loads; creating a Context requires an sm_100 CUDA device and
raises HydroError otherwise.

Names follow the reference's domain: geometries, scenes, poses X_WG, contact
surfaces with ids (id_M, id_N), polygon / triangle representation
(geometry/proximity_engine.h:309-326, query_results/contact_surface.h).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import build as _build

TRIANGLE, POLYGON = 0, 1
UNDEFINED, RIGID, COMPLIANT = 0, 1, 2
STRICT, WITH_FALLBACK = 0, 1

OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_DEGENERATE_FIELD, ERR_CAPACITY = 1, 2, 3, 4, 5
BVH_HOST_MEDIAN_SPLIT, BVH_GPU_LBVH = 0, 1
INTERSECTOR_VOLUME_BVH, INTERSECTOR_SURFACE_BFS = 0, 1
ERR_RIGID_RIGID, ERR_UNSUPPORTED, ERR_NOT_READY, ERR_PLANE_NORMAL, ERR_POLYGON_SIZE = 6, 7, 8, 9, 10
ERR_HALF_SPACE_HALF_SPACE, ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH, ERR_REFERENCE_DEMAND, ERR_HOST = 11, 12, 13, 14


class HydroError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"dbh status {status}: {message}")
        self.status = status


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class AssetInfo(C.Structure):
    _fields_ = [("type", C.c_int32), ("num_vertices", C.c_int32), ("num_elements", C.c_int32),
                ("num_bvh_nodes", C.c_int32), ("bvh_height", C.c_int32), ("is_half_space", C.c_int32),
                ("reserved", C.c_int32 * 2),
                ("root_center", C.c_double * 3), ("root_half_width", C.c_double * 3)]


class BatchInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("num_scenes", "num_geometries", "num_geometry_pairs", "num_soft_rigid_pairs", "num_soft_soft_pairs",
                 "num_rigid_rigid_pairs", "num_unsupported_pairs", "max_traversal_levels", "num_half_space_pairs",
                 "num_rejected_half_space_pairs")]


class ResultSizes(C.Structure):
    _fields_ = [("num_surfaces", C.c_int64), ("num_faces", C.c_int64), ("num_vertices", C.c_int64),
                ("num_polygons", C.c_int64), ("face_data_ints", C.c_int64), ("representation", C.c_int32),
                ("reserved", C.c_int32)]


class ResultBuffers(C.Structure):
    _fields_ = [("surf_scene", _ip), ("surf_id_M", _lp), ("surf_id_N", _lp), ("surf_face_offset", _lp),
                ("surf_vertex_offset", _lp), ("surf_total_area", _dp), ("surf_centroid", _dp), ("surf_flags", _bp),
                ("vertices", _dp), ("pressures", _dp), ("face_data", _ip), ("areas", _dp), ("normals", _dp),
                ("centroids", _dp), ("pressure_at_centroid", _dp), ("grad_eM", _dp), ("grad_eN", _dp),
                ("elem_M", _ip), ("elem_N", _ip), ("vertex_origin", _ip)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("geometry_pairs_tested", "geometry_pairs_overlapping", "bv_pair_tests", "tet_tri_pairs",
                 "tet_tet_pairs", "polygons", "faces", "vertices", "surfaces", "algorithmic_bytes")] + \
               [("total_area", C.c_double), ("tet_tri_full_clips", C.c_int64), ("tet_tet_full_clips", C.c_int64),
                ("leaf_pairs", C.c_int64), ("cull_violations", C.c_int64), ("exceptional_clips", C.c_int64),
                ("leaf_exact_tests", C.c_int64), ("half_space_elements", C.c_int64), ("bv_pair_tests_tet_tet", C.c_int64),
                ("leaf_pairs_tet_tet", C.c_int64), ("discrete_pairs", C.c_int64),
                ("fallback_pairs", C.c_int64), ("bfs_pairs", C.c_int64), ("bfs_waves", C.c_int64)]


_lib = None


def lib():
    """Loads (building in-tree if stale) libdrake_b200_hydro.so. Fails loudly if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale():
        try:
            path = _build.build()
        except Exception as e:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(_build.LIB_PATH):
                raise HydroError(ERR_NO_DEVICE, f"libdrake_b200_hydro.so is missing and cannot be built: {e}")
            path = _build.LIB_PATH
    L = C.CDLL(path)
    L.dbh_last_error.restype = C.c_char_p
    L.dbh_batch_stream.restype = C.c_void_p
    vp = C.c_void_p
    L.dbh_context_create.argtypes = [C.c_int32, C.POINTER(vp)]
    L.dbh_context_destroy.argtypes = [vp]
    L.dbh_register_compliant_mesh.argtypes = [vp, C.c_int32, _dp, C.c_int32, _ip, _dp, _ip]
    L.dbh_register_rigid_mesh.argtypes = [vp, C.c_int32, _dp, C.c_int32, _ip, _ip]
    L.dbh_register_half_space.argtypes = [vp, C.c_int32, C.c_double, C.c_double, _ip]
    L.dbh_asset_get_info.argtypes = [vp, C.c_int32, C.POINTER(AssetInfo)]
    L.dbh_context_update_vertices.argtypes = [vp, C.c_int32, _ip, _dp]
    L.dbh_context_update_vertices_device.argtypes = [vp, C.c_int32, _ip, vp]
    L.dbh_asset_get_derived.argtypes = [vp, C.c_int32, _dp, _dp, _dp]
    L.dbh_asset_copy_bvh.argtypes = [vp, C.c_int32, _dp, _ip]
    L.dbh_asset_copy_topology.argtypes = [vp, C.c_int32, _ip, _ip]
    L.dbh_context_set_bvh_builder.argtypes = [vp, C.c_int32]
    L.dbh_context_set_soft_soft_intersector.argtypes = [vp, C.c_int32]
    L.dbh_batch_create.argtypes = [vp, C.c_int32, _ip, _ip, _lp, _bp, C.c_int32, _ip, C.POINTER(vp)]
    L.dbh_batch_destroy.argtypes = [vp]
    L.dbh_batch_set_capacity.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.dbh_batch_grow_capacity.argtypes = [vp]
    L.dbh_batch_set_chunks.argtypes = [vp, C.c_int32]
    L.dbh_batch_get_chunks.argtypes = [vp]
    L.dbh_batch_last_launch_count.argtypes = [vp]
    L.dbh_batch_get_info.argtypes = [vp, C.POINTER(BatchInfo)]
    L.dbh_batch_set_poses_host.argtypes = [vp, _dp]
    L.dbh_batch_set_poses_device.argtypes = [vp, vp]
    L.dbh_batch_query.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.dbh_batch_sync.argtypes = [vp]
    L.dbh_batch_stream.argtypes = [vp]
    L.dbh_batch_result_sizes.argtypes = [vp, C.POINTER(ResultSizes)]
    L.dbh_batch_copy_results.argtypes = [vp, C.POINTER(ResultBuffers)]
    L.dbh_batch_copy_results_async.argtypes = [vp, C.POINTER(ResultBuffers)]
    L.dbh_batch_stats.argtypes = [vp, C.POINTER(Stats)]
    L.dbh_batch_copy_candidates.argtypes = [vp, C.c_int32, _ip, _lp]
    L.dbh_batch_copy_geometry_pairs.argtypes = [vp, _ip]
    L.dbh_batch_set_body_kinematics.argtypes = [vp, _dp, _dp]
    L.dbh_batch_set_contact_properties.argtypes = [vp, _dp, _dp, _dp]
    L.dbh_batch_compute_spatial_forces.argtypes = [vp, C.c_double]
    L.dbh_batch_copy_spatial_forces.argtypes = [vp, _dp, _dp, _dp]
    L.dbh_batch_copy_spatial_forces_async.argtypes = [vp, vp, _lp]
    L.dbh_batch_compute_discrete_pairs.argtypes = [vp]
    L.dbh_batch_copy_discrete_pairs_async.argtypes = [vp, vp, _lp]
    L.dbh_batch_compute_deformable_contact.argtypes = [vp]
    L.dbh_batch_copy_deformable_contact_async.argtypes = [vp, vp, vp, vp, _lp]
    L.dbh_batch_enable_timing.argtypes = [vp, C.c_int32]
    L.dbh_batch_set_verify.argtypes = [vp, C.c_int32]
    L.dbh_batch_set_normal_filter.argtypes = [vp, C.c_int32]
    L.dbh_batch_set_small_pair_path.argtypes = [vp, C.c_int32]
    L.dbh_batch_set_geometry_bounds.argtypes = [vp, _dp]
    L.dbh_batch_copy_fallback_pairs.argtypes = [vp, _ip, _lp]
    L.dbh_batch_set_record_candidates.argtypes = [vp, C.c_int32]
    L.dbh_batch_kernel_times_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.dbh_batch_stage_times_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int32]
    L.dbh_mesh_last_error.restype = C.c_char_p
    dbl = C.c_double
    L.dbh_mesh_make_sphere_volume.argtypes = [dbl, dbl, C.c_int32, C.POINTER(vp)]
    L.dbh_mesh_make_ellipsoid_volume.argtypes = [dbl, dbl, dbl, dbl, C.c_int32, C.POINTER(vp)]
    L.dbh_mesh_make_box_volume_with_ma.argtypes = [dbl, dbl, dbl, C.POINTER(vp)]
    L.dbh_mesh_make_box_volume.argtypes = [dbl, dbl, dbl, dbl, C.POINTER(vp)]
    L.dbh_mesh_make_cylinder_volume_with_ma.argtypes = [dbl, dbl, dbl, C.POINTER(vp)]
    L.dbh_mesh_make_capsule_volume.argtypes = [dbl, dbl, dbl, C.POINTER(vp)]
    L.dbh_mesh_convert_volume_to_surface.argtypes = [vp, C.POINTER(vp)]
    L.dbh_mesh_set_sphere_pressure.argtypes = [vp, dbl, dbl, dbl]
    L.dbh_mesh_set_ellipsoid_pressure.argtypes = [vp, dbl, dbl, dbl, dbl, dbl]
    L.dbh_mesh_set_box_pressure.argtypes = [vp, dbl, dbl, dbl, dbl, dbl]
    L.dbh_mesh_set_cylinder_pressure.argtypes = [vp, dbl, dbl, dbl, dbl]
    L.dbh_mesh_set_capsule_pressure.argtypes = [vp, dbl, dbl, dbl, dbl]
    L.dbh_mesh_from_arrays.argtypes = [C.c_int32, _dp, C.c_int32, _ip, C.c_int32, C.POINTER(vp)]
    L.dbh_mesh_make_convex_volume.argtypes = [vp, C.POINTER(vp)]
    L.dbh_mesh_set_volume_mesh_pressure.argtypes = [vp, dbl, dbl]
    L.dbh_mesh_get_sizes.argtypes = [vp, _ip, _ip, _ip, _ip]
    L.dbh_mesh_copy.argtypes = [vp, _dp, _ip, _dp]
    L.dbh_mesh_destroy.argtypes = [vp]
    L.dbh_register_mesh.argtypes = [vp, vp, _ip]
    _lib = L
    return L


def _check(status: int):
    if status != OK:
        raise HydroError(status, lib().dbh_last_error().decode())


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


SINGLE_INTERIOR_VERTEX, DENSE_INTERIOR_VERTICES = 0, 1


class Mesh:
    """A generated primitive-shape mesh (host memory, set-up time): VolumeMesh (+ pressure field) or
    TriangleSurfaceMesh. Static constructors mirror the reference generators."""

    def __init__(self, handle):
        self._h = handle
        nv, ne, k, hp = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _check_mesh(lib().dbh_mesh_get_sizes(self._h, C.byref(nv), C.byref(ne), C.byref(k), C.byref(hp)))
        self.num_vertices, self.num_elements, self.vertices_per_element = nv.value, ne.value, k.value
        self._cache = None

    def __del__(self):
        try:
            if self._h:
                lib().dbh_mesh_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def _make(fn, *args) -> "Mesh":
        h = C.c_void_p()
        _check_mesh(fn(*args, C.byref(h)))
        return Mesh(h)

    @staticmethod
    def sphere_volume(radius, resolution_hint, strategy=SINGLE_INTERIOR_VERTEX) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_sphere_volume, radius, resolution_hint, strategy)

    @staticmethod
    def ellipsoid_volume(a, b, c, resolution_hint, strategy=SINGLE_INTERIOR_VERTEX) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_ellipsoid_volume, a, b, c, resolution_hint, strategy)

    @staticmethod
    def box_volume_with_ma(x, y, z) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_box_volume_with_ma, x, y, z)

    @staticmethod
    def box_volume(x, y, z, resolution_hint) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_box_volume, x, y, z, resolution_hint)

    @staticmethod
    def cylinder_volume_with_ma(radius, length, resolution_hint) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_cylinder_volume_with_ma, radius, length, resolution_hint)

    @staticmethod
    def capsule_volume(radius, length, resolution_hint) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_make_capsule_volume, radius, length, resolution_hint)

    @staticmethod
    def from_arrays(vertices, elements) -> "Mesh":
        """A caller's tetrahedral ([ne,4]) or triangle ([ne,3]) mesh."""
        v, e = _d(vertices).reshape(-1, 3), _i(elements)
        if e.ndim != 2 or e.shape[1] not in (3, 4):
            raise ValueError("elements must be [ne, 4] (tetrahedra) or [ne, 3] (triangles)")
        return Mesh._make(lib().dbh_mesh_from_arrays, len(v), _ptr(v, _dp), len(e), _ptr(e, _ip), e.shape[1])

    def to_convex_volume(self) -> "Mesh":
        """MakeConvexVolumeMesh of a closed triangle surface (make_convex_mesh.cc:17-60)."""
        return Mesh._make(lib().dbh_mesh_make_convex_volume, self._h)

    def to_surface(self) -> "Mesh":
        return Mesh._make(lib().dbh_mesh_convert_volume_to_surface, self._h)

    @staticmethod
    def sphere_surface(radius, resolution_hint) -> "Mesh":
        return Mesh.sphere_volume(radius, resolution_hint, SINGLE_INTERIOR_VERTEX).to_surface()

    @staticmethod
    def ellipsoid_surface(a, b, c, resolution_hint) -> "Mesh":
        return Mesh.ellipsoid_volume(a, b, c, resolution_hint, SINGLE_INTERIOR_VERTEX).to_surface()

    @staticmethod
    def box_surface(x, y, z) -> "Mesh":
        return Mesh.box_volume_with_ma(x, y, z).to_surface()

    def set_sphere_pressure(self, radius, modulus, margin=0.0) -> "Mesh":
        _check_mesh(lib().dbh_mesh_set_sphere_pressure(self._h, radius, modulus, margin))
        self._cache = None
        return self

    def set_ellipsoid_pressure(self, a, b, c, modulus, margin=0.0) -> "Mesh":
        _check_mesh(lib().dbh_mesh_set_ellipsoid_pressure(self._h, a, b, c, modulus, margin))
        self._cache = None
        return self

    def set_box_pressure(self, x, y, z, modulus, margin=0.0) -> "Mesh":
        _check_mesh(lib().dbh_mesh_set_box_pressure(self._h, x, y, z, modulus, margin))
        self._cache = None
        return self

    def set_volume_mesh_pressure(self, modulus, margin=0.0) -> "Mesh":
        """MakeVolumeMeshPressureField (make_mesh_field.cc:45-96)."""
        _check_mesh(lib().dbh_mesh_set_volume_mesh_pressure(self._h, modulus, margin))
        self._cache = None
        return self

    def set_cylinder_pressure(self, radius, length, modulus, margin=0.0) -> "Mesh":
        _check_mesh(lib().dbh_mesh_set_cylinder_pressure(self._h, radius, length, modulus, margin))
        self._cache = None
        return self

    def set_capsule_pressure(self, radius, length, modulus, margin=0.0) -> "Mesh":
        _check_mesh(lib().dbh_mesh_set_capsule_pressure(self._h, radius, length, modulus, margin))
        self._cache = None
        return self

    def arrays(self):
        """(vertices [V,3], elements [E,k], pressure [V] or None)."""
        if self._cache is None:
            hp = C.c_int32()
            _check_mesh(lib().dbh_mesh_get_sizes(self._h, None, None, None, C.byref(hp)))
            v = np.zeros((self.num_vertices, 3))
            e = np.zeros((self.num_elements, self.vertices_per_element), np.int32)
            p = np.zeros(self.num_vertices) if hp.value else None
            _check_mesh(lib().dbh_mesh_copy(self._h, _ptr(v, _dp), _ptr(e, _ip), _ptr(p, _dp)))
            self._cache = (v, e, p)
        return self._cache

    @property
    def vertices(self):
        return self.arrays()[0]

    @property
    def elements(self):
        return self.arrays()[1]

    @property
    def pressure(self):
        return self.arrays()[2]


def _check_mesh(status: int):
    if status != OK:
        raise HydroError(status, lib().dbh_mesh_last_error().decode())


class Context:
    """One per device (one process per GPU)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().dbh_context_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().dbh_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def register_compliant_mesh(self, vertices, tets, pressure) -> int:
        """VolumeMesh + VolumeMeshFieldLinear -> asset handle (hydroelastic_internal.cc:94-108)."""
        v, t, p = _d(vertices).reshape(-1, 3), _i(tets).reshape(-1, 4), _d(pressure).reshape(-1)
        if len(p) != len(v):
            raise ValueError("one pressure value per vertex")
        out = C.c_int32()
        _check(lib().dbh_register_compliant_mesh(self._h, len(v), _ptr(v, _dp), len(t), _ptr(t, _ip), _ptr(p, _dp),
                                                 C.byref(out)))
        return out.value

    def register_rigid_mesh(self, vertices, tris) -> int:
        v, t = _d(vertices).reshape(-1, 3), _i(tris).reshape(-1, 3)
        out = C.c_int32()
        _check(lib().dbh_register_rigid_mesh(self._h, len(v), _ptr(v, _dp), len(t), _ptr(t, _ip), C.byref(out)))
        return out.value

    def update_vertices(self, assets, vertices):
        """Deformable reuse: new vertex positions for mesh assets (DeformableMeshWithBvh::UpdateVertexPositions): triangle
        meshes (deforming surfaces) or volume meshes (deformable bodies carrying a signed-distance field): per-element
        records derived again and the BVHs refitted on the device. vertices: one [nv,3] array per asset."""
        ids = _i(list(assets))
        flat = np.ascontiguousarray(np.concatenate([_d(v).reshape(-1) for v in vertices]))
        _check(lib().dbh_context_update_vertices(self._h, len(ids), _ptr(ids, _ip), _ptr(flat, _dp)))

    def update_vertices_device(self, assets, device_pointer: int):
        """update_vertices with the new positions already on this device: the address of one float64 device buffer that
        holds the listed assets' [num_vertices, 3] arrays one after the other (e.g. torch_tensor.data_ptr())."""
        ids = _i(list(assets))
        _check(lib().dbh_context_update_vertices_device(self._h, len(ids), _ptr(ids, _ip), C.c_void_p(int(device_pointer))))

    def register_half_space(self, compliant: bool = False, pressure_scale: float = 0.0, margin: float = 0.0) -> int:
        """HalfSpace shape (z <= 0 of its frame). compliant: pressure_scale = hydroelastic modulus / slab thickness
        (hydroelastic_internal.h SoftHalfSpace). Pairs with it are CalcRigidCompliant's special cases
        (hydroelastic_calculator.cc:25-41)."""
        out = C.c_int32()
        _check(lib().dbh_register_half_space(self._h, COMPLIANT if compliant else RIGID, float(pressure_scale),
                                             float(margin), C.byref(out)))
        return out.value

    def register_mesh(self, mesh: "Mesh") -> int:
        """Volume mesh with pressure -> compliant asset; surface mesh -> rigid asset."""
        out = C.c_int32()
        _check(lib().dbh_register_mesh(self._h, mesh._h, C.byref(out)))
        return out.value

    def asset_info(self, asset: int) -> AssetInfo:
        info = AssetInfo()
        _check(lib().dbh_asset_get_info(self._h, asset, C.byref(info)))
        return info

    def set_bvh_builder(self, builder: int):
        """BVH_HOST_MEDIAN_SPLIT (default) or BVH_GPU_LBVH, for the registrations that follow."""
        _check(lib().dbh_context_set_bvh_builder(self._h, builder))

    def set_soft_soft_intersector(self, intersector: int):
        """INTERSECTOR_VOLUME_BVH (default: what ComputeContactSurfaceFromCompliantVolumes runs) or
        INTERSECTOR_SURFACE_BFS (the reference's surface-BVH + BFS overload, field_intersection.cc:254-375), for the
        compliant-mesh registrations and the batches created afterwards."""
        _check(lib().dbh_context_set_soft_soft_intersector(self._h, intersector))

    def asset_topology(self, asset: int):
        """(neighbors [num_elements, 4], number of boundary triangles) of a compliant mesh registered for the surface
        intersector (VolumeMeshTopology, volume_mesh_topology.h)."""
        n = self.asset_info(asset).num_elements
        nb, ntri = np.zeros((n, 4), np.int32), C.c_int32()
        _check(lib().dbh_asset_copy_topology(self._h, asset, _ptr(nb, _ip), C.byref(ntri)))
        return nb, ntri.value

    def asset_bvh(self, asset: int):
        """(boxes [n,6] = centre, half width; left [n]) with n = num_bvh_nodes + 1 (node 0 unused, 1 = root)."""
        n = self.asset_info(asset).num_bvh_nodes + 1
        boxes, left = np.zeros((n, 6)), np.zeros(n, np.int32)
        _check(lib().dbh_asset_copy_bvh(self._h, asset, _ptr(boxes, _dp), _ptr(left, _ip)))
        return boxes, left

    def asset_derived(self, asset: int):
        """Compliant: (gradients [T,3], values_at_Mo [T], inward_normals [T,4,3]); rigid: normals [T,3]."""
        info = self.asset_info(asset)
        n = info.num_elements
        if info.type == COMPLIANT:
            g, m, nn = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 4, 3))
            _check(lib().dbh_asset_get_derived(self._h, asset, _ptr(g, _dp), _ptr(m, _dp), _ptr(nn, _dp)))
            return g, m, nn
        g = np.zeros((n, 3))
        _check(lib().dbh_asset_get_derived(self._h, asset, _ptr(g, _dp), None, None))
        return g


@dataclass
class ContactSurface:
    """Host copy of one ContactSurface<double> (query_results/contact_surface.h:160-415)."""
    scene: int
    id_M: int
    id_N: int
    representation: int
    vertices: np.ndarray            # [V,3] p_WV
    pressures: np.ndarray           # [V]
    face_data: np.ndarray           # polygon: |n|v..|...; triangle: [F,3]; indices local to the surface
    areas: np.ndarray
    normals: np.ndarray
    centroids: np.ndarray
    pressure_at_centroid: np.ndarray
    grad_eM: Optional[np.ndarray]
    grad_eN: Optional[np.ndarray]
    total_area: float
    centroid: np.ndarray
    elem_M: np.ndarray              # emitting element of the first-constructed body
    elem_N: np.ndarray
    swapped: bool
    vertex_origin: Optional[np.ndarray] = None   # half-space surfaces copied with vertex_origin=True (see weld())

    @property
    def num_faces(self) -> int:
        return len(self.areas)

    def weld(self, elements) -> "ContactSurface":
        """The same half-space surface with the vertices the reference shares between polygons merged
        (mesh_plane_intersection.cc:205-231 cut_edges; mesh_half_space_intersection.cc:52-101): every polygon vertex is a
        copy of a vertex of the pair's one mesh or a cut of one of its edges (vertex_origin), so vertices with the same
        mesh vertex / mesh edge become one -- the first occurrence in face order keeps its position and pressure, as in
        the reference's maps. elements: [E,4] tetrahedra or [E,3] triangles of that mesh. Face quantities are unchanged."""
        assert self.vertex_origin is not None, "copy the results with vertex_origin=True"
        elements = np.asarray(elements)
        elem = self.elem_M if np.all(self.elem_M >= 0) else self.elem_N
        tri = self.representation == TRIANGLE
        fd = np.asarray(self.face_data).reshape(-1)
        new_index, keep, key_to_new = np.zeros(len(self.vertices), np.int64), [], {}
        seen = np.zeros(len(self.vertices), bool)
        i = 0
        for f in range(self.num_faces):
            if tri:
                idx = fd[3 * f:3 * f + 3]
            else:
                n = fd[i]
                idx = fd[i + 1:i + 1 + n]
                i += n + 1
            for v in idx:
                if seen[v]:
                    continue
                seen[v] = True
                o = int(self.vertex_origin[v])
                if o < 0:
                    key = ("c", int(v))                       # the centroid of a triangle fan: never shared
                elif o < 16:
                    key = ("v", int(elements[elem[f]][o]))
                else:
                    a, b = (o - 16) // 4, (o - 16) % 4
                    va, vb = int(elements[elem[f]][a]), int(elements[elem[f]][b])
                    key = ("e", min(va, vb), max(va, vb))
                if key not in key_to_new:
                    key_to_new[key] = len(keep)
                    keep.append(int(v))
                new_index[v] = key_to_new[key]
        out = np.array(fd, copy=True)
        if tri:
            out = new_index[out].astype(np.int32).reshape(-1, 3)
        else:
            i = 0
            for f in range(self.num_faces):
                n = fd[i]
                out[i + 1:i + 1 + n] = new_index[fd[i + 1:i + 1 + n]]
                i += n + 1
        import dataclasses
        return dataclasses.replace(self, vertices=self.vertices[keep], pressures=self.pressures[keep], face_data=out,
                                   vertex_origin=self.vertex_origin[keep])

    def polygons(self) -> List[np.ndarray]:
        if self.representation == TRIANGLE:
            return [t for t in self.face_data]
        out, i, fd = [], 0, self.face_data
        while i < len(fd):
            n = fd[i]
            out.append(fd[i + 1:i + 1 + n])
            i += n + 1
        return out


class PackedResults:
    """The packed host buffers of one query (all surfaces of all scenes)."""

    _LAYOUT = [("surf_scene", np.int32, "s", 1), ("surf_id_M", np.int64, "s", 1), ("surf_id_N", np.int64, "s", 1),
               ("surf_face_offset", np.int64, "s+1", 1), ("surf_vertex_offset", np.int64, "s+1", 1),
               ("surf_total_area", np.float64, "s", 1), ("surf_centroid", np.float64, "s", 3),
               ("surf_flags", np.uint8, "s", 1), ("vertices", np.float64, "v", 3), ("pressures", np.float64, "v", 1),
               ("face_data", np.int32, "d", 1), ("areas", np.float64, "f", 1), ("normals", np.float64, "f", 3),
               ("centroids", np.float64, "f", 3), ("pressure_at_centroid", np.float64, "f", 1),
               ("grad_eM", np.float64, "f", 3), ("grad_eN", np.float64, "f", 3), ("elem_M", np.int32, "f", 1),
               ("elem_N", np.int32, "f", 1), ("vertex_origin", np.int32, "v", 1)]

    # What MultibodyPlant's discrete update reads from a ContactSurface: per-face area, normal,
    # centroid, pressure at centroid, both gradients; per-surface ids, offsets, total area and centroid.
    FACE_FIELDS = ("surf_scene", "surf_id_M", "surf_id_N", "surf_face_offset", "surf_vertex_offset", "surf_total_area",
                   "surf_centroid", "surf_flags", "areas", "normals", "centroids", "pressure_at_centroid", "grad_eM",
                   "grad_eN")

    def __init__(self, sizes: ResultSizes, allocator=None, capacity: Optional[ResultSizes] = None, fields=None,
                 vertex_origin: bool = False):
        """allocator(shape, dtype) -> ndarray lets callers supply pinned host memory; `capacity`
        (>= sizes) allocates larger buffers once so they can be reused across queries; `fields` (names of
        dbh_result_buffers members) restricts the copy-out to those arrays (the others stay NULL / empty)."""
        self.representation = sizes.representation
        self._alloc = allocator or (lambda shape, dtype: np.zeros(shape, dtype))
        self._fields = None if fields is None else set(fields)
        self._origin = vertex_origin   # only batches with half-space pairs have it (see ContactSurface.weld)
        cap = capacity or sizes
        self._full = {}
        counts = {"s": cap.num_surfaces, "s+1": cap.num_surfaces + 1, "v": cap.num_vertices, "f": cap.num_faces,
                  "d": cap.face_data_ints}
        for name, dtype, kind, width in self._LAYOUT:
            n = counts[kind] if self._wanted(name) else 0
            shape = (n,) if width == 1 else (n, width)
            self._full[name] = self._alloc(shape, dtype) if n else np.zeros(shape, dtype)
        self._cap = {k: int(v) for k, v in counts.items()}
        self.resize(sizes)

    def _wanted(self, name: str) -> bool:
        if name == "vertex_origin":
            return self._origin
        return self._fields is None or name in self._fields

    def fits(self, sizes: ResultSizes) -> bool:
        c = self._cap
        return (c["s"] >= sizes.num_surfaces and c["v"] >= sizes.num_vertices and c["f"] >= sizes.num_faces and
                c["d"] >= sizes.face_data_ints)

    def resize(self, sizes: ResultSizes):
        """Re-slices the views to the sizes of the latest query (no reallocation)."""
        self.representation = sizes.representation
        counts = {"s": sizes.num_surfaces, "s+1": sizes.num_surfaces + 1, "v": sizes.num_vertices,
                  "f": sizes.num_faces, "d": sizes.face_data_ints}
        for name, _, kind, _ in self._LAYOUT:
            setattr(self, name, self._full[name][:counts[kind] if self._wanted(name) else 0])

    def nbytes(self) -> int:
        return sum(getattr(self, name).nbytes for name, _, _, _ in self._LAYOUT)

    def c_struct(self) -> ResultBuffers:
        b = ResultBuffers()
        for name, t in ResultBuffers._fields_:
            if self._wanted(name):
                setattr(b, name, _ptr(getattr(self, name), t))
        return b

    @property
    def num_surfaces(self) -> int:
        return len(self.surf_face_offset) - 1 if self._wanted("surf_face_offset") else len(self.surf_scene)

    def surfaces(self, scenes=None) -> List[ContactSurface]:
        """Per-surface views of the packed arrays, in result order; `scenes`: only the surfaces of these scenes."""
        out = []
        poly = self.representation == POLYGON
        which = range(self.num_surfaces)
        if scenes is not None:
            which = np.nonzero(np.isin(self.surf_scene, np.asarray(list(scenes))))[0]
        for s in which:
            f0, f1 = self.surf_face_offset[s], self.surf_face_offset[s + 1]
            v0, v1 = self.surf_vertex_offset[s], self.surf_vertex_offset[s + 1]
            flags = int(self.surf_flags[s])
            if poly:
                # |n|v0..vn-1| per face, faces and vertices of a surface contiguous and every polygon with its own
                # vertices: the surface's slice starts after f0 counts and v0 indices.
                fd = self.face_data[f0 + v0:f1 + v1]
            else:
                fd = self.face_data[3 * f0:3 * f1].reshape(-1, 3)
            out.append(ContactSurface(
                int(self.surf_scene[s]), int(self.surf_id_M[s]), int(self.surf_id_N[s]), self.representation,
                self.vertices[v0:v1], self.pressures[v0:v1], fd, self.areas[f0:f1], self.normals[f0:f1],
                self.centroids[f0:f1], self.pressure_at_centroid[f0:f1],
                self.grad_eM[f0:f1] if flags & 1 else None, self.grad_eN[f0:f1] if flags & 2 else None,
                float(self.surf_total_area[s]), self.surf_centroid[s], self.elem_M[f0:f1], self.elem_N[f0:f1],
                bool(flags & 4), self.vertex_origin[v0:v1] if self._origin else None))
        return out


class SceneBatch:
    """Registration state of many independent scenes + the batched query.

    Mirrors the ProximityEngine calls on this path: geometries are added with
    an asset (their hydroelastic representation), a GeometryId and an anchored
    flag (AddDynamicGeometry / AddAnchoredGeometry, proximity_engine.h:122-133);
    `excluded` lists filtered pairs (collision_filter.h:158).
    """

    # Test hook: (small-pair path mode, record candidates) applied to every batch created while it is set
    # (tests/test_gpu_parity.py runs its cases through both mid-phase implementations).
    test_defaults = None

    def __init__(self, ctx: Context, scene_offsets, geom_asset, geom_id, geom_anchored=None, excluded=None):
        self.ctx = ctx
        so, ga, gi = _i(scene_offsets), _i(geom_asset), np.ascontiguousarray(geom_id, np.int64)
        anch = np.ascontiguousarray(geom_anchored if geom_anchored is not None else np.zeros(len(ga)), np.uint8)
        ex = _i(excluded if excluded is not None else np.zeros((0, 2))).reshape(-1, 2)
        self._h = C.c_void_p()
        _check(lib().dbh_batch_create(ctx._h, len(so) - 1, _ptr(so, _ip), _ptr(ga, _ip), _ptr(gi, _lp),
                                      _ptr(anch, _bp), len(ex), _ptr(ex, _ip), C.byref(self._h)))
        self.num_scenes = len(so) - 1
        self.num_geometries = len(ga)
        self.scene_offsets, self.geom_asset, self.geom_id = so, ga, gi
        self._keep = None
        self._grow = 0
        if SceneBatch.test_defaults is not None:
            self.set_small_pair_path(SceneBatch.test_defaults[0])
            self.set_record_candidates(SceneBatch.test_defaults[1])

    @staticmethod
    def replicated(ctx: Context, num_scenes: int, assets: Sequence[int], ids: Sequence[int], anchored=None,
                   excluded_local=None) -> "SceneBatch":
        """num_scenes instances of one scene template (geometries listed once)."""
        g = len(assets)
        so = np.arange(num_scenes + 1, dtype=np.int32) * g
        ga = np.tile(_i(assets), num_scenes)
        gi = np.tile(np.asarray(ids, np.int64), num_scenes)
        an = np.tile(np.asarray(anchored if anchored is not None else np.zeros(g), np.uint8), num_scenes)
        ex = None
        if excluded_local is not None and len(excluded_local):
            e = _i(excluded_local).reshape(-1, 2)
            ex = (e[None, :, :] + (np.arange(num_scenes, dtype=np.int32) * g)[:, None, None]).reshape(-1, 2)
        return SceneBatch(ctx, so, ga, gi, an, ex)

    def close(self):
        if self._h:
            lib().dbh_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> BatchInfo:
        out = BatchInfo()
        _check(lib().dbh_batch_get_info(self._h, C.byref(out)))
        return out

    def set_capacity(self, bv_pairs: int = 0, element_pairs: int = 0, faces: int = 0):
        _check(lib().dbh_batch_set_capacity(self._h, bv_pairs, element_pairs, faces))

    def set_chunks(self, num_chunks: int):
        """Process the batch in `num_chunks` chunks of scenes that reuse the device work lists."""
        _check(lib().dbh_batch_set_chunks(self._h, num_chunks))

    @property
    def num_chunks(self) -> int:
        return int(lib().dbh_batch_get_chunks(self._h))

    @property
    def last_launch_count(self) -> int:
        return int(lib().dbh_batch_last_launch_count(self._h))

    def set_verify(self, on: bool = True):
        """Self-check mode: exact clipping for every candidate; stats()['cull_violations'] must stay 0."""
        _check(lib().dbh_batch_set_verify(self._h, int(on)))

    def set_geometry_bounds(self, boxes):
        """[G,6] centre + half widths, in the geometry frame, of the geometries registered with asset -1 (no
        hydroelastic representation); NaN rows = no bound. Pairs that produce no surface only count when the world boxes
        overlap (the reference's FindCollisionCandidates)."""
        B = _d(boxes).reshape(-1, 6)
        if len(B) != self.num_geometries:
            raise ValueError("one box per geometry")
        _check(lib().dbh_batch_set_geometry_bounds(self._h, _ptr(B, _dp)))

    def fallback_pairs(self) -> np.ndarray:
        """After a WITH_FALLBACK query: [n,4] = scene, flat geometry index (smaller id), flat geometry index (larger id),
        kind, of the pairs the reference hands to the point-pair narrow phase."""
        n = C.c_int64(0)
        _check(lib().dbh_batch_copy_fallback_pairs(self._h, None, C.byref(n)))
        out = np.zeros((n.value, 4), np.int32)
        if n.value:
            _check(lib().dbh_batch_copy_fallback_pairs(self._h, _ptr(out, _ip), C.byref(n)))
        return out

    def set_normal_filter(self, on: bool = True):
        """Rigid-compliant mesh pairs with (default, ComputeContactSurfaces) or without the face-normal filter (the
        reference's mesh_intersection_benchmark)."""
        _check(lib().dbh_batch_set_normal_filter(self._h, int(on)))

    def set_small_pair_path(self, mode: int):
        """Mid phase of small geometry pairs on chip, one warp per pair (k_small_pairs): 0 never, 1 every eligible
        pair, 2 (default) when a chunk holds enough of them to fill the device."""
        _check(lib().dbh_batch_set_small_pair_path(self._h, int(mode)))

    def set_record_candidates(self, on: bool = True):
        """Keep the candidate element-pair lists of the following queries for candidates()."""
        _check(lib().dbh_batch_set_record_candidates(self._h, int(on)))

    def enable_timing(self, on: bool = True):
        _check(lib().dbh_batch_enable_timing(self._h, int(on)))

    def set_poses(self, X_WGs):
        """UpdateWorldPoses: [G,12] host array (R column-major, then p)."""
        P = _d(X_WGs).reshape(-1, 12)
        if len(P) != self.num_geometries:
            raise ValueError(f"expected {self.num_geometries} poses, got {len(P)}")
        self._keep = P  # the H2D copy is asynchronous
        _check(lib().dbh_batch_set_poses_host(self._h, _ptr(P, _dp)))

    def set_poses_host_ptr(self, host_ptr: int):
        """Same as set_poses for a caller-owned (ideally pinned) host buffer of G*12 doubles."""
        _check(lib().dbh_batch_set_poses_host(self._h, C.cast(C.c_void_p(host_ptr), _dp)))

    def set_poses_device(self, device_ptr: int):
        _check(lib().dbh_batch_set_poses_device(self._h, C.c_void_p(device_ptr)))

    def query(self, representation: int = POLYGON, kind: int = STRICT, scene_stats_device_ptr: int = 0):
        _check(lib().dbh_batch_query(self._h, representation, kind,
                                     C.c_void_p(scene_stats_device_ptr) if scene_stats_device_ptr else None))

    def sync(self):
        _check(lib().dbh_batch_sync(self._h))

    def run(self, representation: int = POLYGON, kind: int = STRICT, scene_stats_device_ptr: int = 0,
            max_grow: int = 48):
        """query + sync, doubling device work-list capacities on DBH_ERR_CAPACITY."""
        for _ in range(max_grow + 1):
            self.query(representation, kind, scene_stats_device_ptr)
            try:
                self.sync()
                return
            except HydroError as e:
                if e.status != ERR_CAPACITY:
                    raise
                if os.environ.get("DBH_DEBUG"):
                    print("[drake_b200] growing work lists:", e, file=sys.stderr, flush=True)
                _check(lib().dbh_batch_grow_capacity(self._h))
        raise HydroError(ERR_CAPACITY, "device work lists still overflow after growing")

    def sizes(self) -> ResultSizes:
        out = ResultSizes()
        _check(lib().dbh_batch_result_sizes(self._h, C.byref(out)))
        return out

    def stats(self) -> Dict[str, float]:
        s = Stats()
        _check(lib().dbh_batch_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_ if n != "reserved"}

    def copy_results(self, into: Optional[PackedResults] = None, wait: bool = True,
                     vertex_origin: bool = False) -> PackedResults:
        """D2H of the packed results. wait=False only enqueues the copies on the batch stream (use pinned
        destination buffers); sync() completes them. vertex_origin: also the per-vertex origin codes of the half-space
        surfaces (ContactSurface.weld)."""
        sizes = self.sizes()
        if into is not None and into.fits(sizes):
            res = into
            res.resize(sizes)
        else:
            res = PackedResults(sizes, vertex_origin=vertex_origin)
        bufs = res.c_struct()
        if wait:
            _check(lib().dbh_batch_copy_results(self._h, C.byref(bufs)))
        else:
            _check(lib().dbh_batch_copy_results_async(self._h, C.byref(bufs)))
        return res

    def compute_contact_surfaces(self, X_WGs, representation: int = POLYGON,
                                 query_kind: int = STRICT) -> List[List[ContactSurface]]:
        """ComputeContactSurfaces (or, with WITH_FALLBACK, the surfaces of ComputeContactSurfacesWithFallback) for
        every scene: per scene a list sorted by (id_M, id_N)."""
        self.set_poses(X_WGs)
        self.run(representation, query_kind)
        out: List[List[ContactSurface]] = [[] for _ in range(self.num_scenes)]
        for s in self.copy_results().surfaces():
            out[s.scene].append(s)
        return out

    # -- device-side consumer: one spatial force per surface ------------------------------
    def set_body_kinematics(self, p_WBo=None, V_WB=None):
        """p_WBo [G,3]: body-frame origins (None: the geometry frames' origins); V_WB [G,6] = (w, v) (None: at rest)."""
        p = _d(p_WBo).reshape(-1, 3) if p_WBo is not None else None
        v = _d(V_WB).reshape(-1, 6) if V_WB is not None else None
        for a in (p, v):
            if a is not None and len(a) != self.num_geometries:
                raise ValueError("one row per geometry")
        _check(lib().dbh_batch_set_body_kinematics(self._h, _ptr(p, _dp) if p is not None else None,
                                                   _ptr(v, _dp) if v is not None else None))
        self._keep_kin = (p, v)

    def set_contact_properties(self, hydroelastic_modulus, hunt_crossley_dissipation, dynamic_friction):
        """Per geometry [G]; modulus = inf for a rigid geometry (hydroelastic_engine.cc:83-104)."""
        E, d, mu = _d(hydroelastic_modulus), _d(hunt_crossley_dissipation), _d(dynamic_friction)
        if not (len(E) == len(d) == len(mu) == self.num_geometries):
            raise ValueError("one value per geometry")
        _check(lib().dbh_batch_set_contact_properties(self._h, _ptr(E, _dp), _ptr(d, _dp), _ptr(mu, _dp)))

    def spatial_forces(self, vslip_regularizer: float = 1e-6):
        """HydroelasticTractionCalculator over the last query's surfaces, on the device: (F_Ac_W, F_Ao_W, F_Bo_W),
        each [num_surfaces, 6] = (tau, f), surfaces in result order."""
        _check(lib().dbh_batch_compute_spatial_forces(self._h, float(vslip_regularizer)))
        self.sync()
        n = int(self.sizes().num_surfaces)
        out = [np.zeros((n, 6)) for _ in range(3)]
        _check(lib().dbh_batch_copy_spatial_forces(self._h, *[_ptr(o, _dp) for o in out]))
        return tuple(out)

    def compute_discrete_pairs(self):
        """Enqueues the per-face discrete contact data of the last query's surfaces (discrete_update_manager.cc:896-1036)."""
        _check(lib().dbh_batch_compute_discrete_pairs(self._h))

    def copy_discrete_pairs(self, out=None, wait: bool = True):
        """[num_faces, 9] = p_WC, nhat_BA_W, phi0, fn0, stiffness (0: a face the reference skips). `out`: a (pinned)
        buffer of at least that size; wait=False leaves the copy in flight (complete with sync())."""
        n = C.c_int64(0)
        _check(lib().dbh_batch_copy_discrete_pairs_async(self._h, None, C.byref(n)))
        if out is None:
            out = np.zeros((n.value, 9))
        view = out.reshape(-1)[: n.value * 9].reshape(n.value, 9)
        if n.value:
            _check(lib().dbh_batch_copy_discrete_pairs_async(self._h, view.ctypes.data, C.byref(n)))
        if wait:
            self.sync()
        return view

    def compute_deformable_contact(self):
        """Enqueues the deformable x rigid contact data of the last query's faces (deformable_mesh_intersection.cc:85-183)."""
        _check(lib().dbh_batch_compute_deformable_contact(self._h))

    def copy_deformable_contact(self, wait: bool = True):
        """([num_faces, 24], element_a [num_faces], element_b [num_faces]). Tet-triangle faces (deformable surface x
        rigid pressure field): p_WC, nhat_W, R_WC (column-major), pressure, grad_W (the reference's point-transformed
        gradient), barycentric centroid in the triangle element_b, 0, 0. Tet-tet faces (two deformable volumes): p_WC,
        nhat_W, R_WC, signed distance, barycentric[4] in tetrahedron element_a of M, barycentric[4] in element_b of N.
        NaN rows for half-space faces."""
        n = C.c_int64(0)
        _check(lib().dbh_batch_copy_deformable_contact_async(self._h, None, None, None, C.byref(n)))
        out = np.zeros((n.value, 24))
        ea = np.zeros(n.value, dtype=np.int32)
        eb = np.zeros(n.value, dtype=np.int32)
        if n.value:
            _check(lib().dbh_batch_copy_deformable_contact_async(self._h, out.ctypes.data, ea.ctypes.data, eb.ctypes.data,
                                                                 C.byref(n)))
        if wait:
            self.sync()
        return out, ea, eb

    def compute_spatial_forces(self, vslip_regularizer: float = 1e-6):
        """Enqueues k_wrench over the last query's surfaces (no host sync)."""
        _check(lib().dbh_batch_compute_spatial_forces(self._h, float(vslip_regularizer)))

    def copy_spatial_forces(self, out, wait: bool = True) -> int:
        """After sync(): one D2H copy of [num_surfaces, 18] = F_Ac_W, F_Ao_W, F_Bo_W per surface into `out` (a pinned
        buffer of at least that size); returns num_surfaces. wait=False leaves the copy in flight."""
        n = C.c_int64(0)
        _check(lib().dbh_batch_copy_spatial_forces_async(self._h, None, C.byref(n)))
        if n.value:
            _check(lib().dbh_batch_copy_spatial_forces_async(self._h, out.ctypes.data, C.byref(n)))
        if wait:
            self.sync()
        return int(n.value)

    def candidates(self, kind: int) -> np.ndarray:
        """(geometry pair, element of first body, element of second body) triples; kind 0 tet-tri, 1 tet-tet."""
        n = C.c_int64()
        _check(lib().dbh_batch_copy_candidates(self._h, kind, None, C.byref(n)))
        out = np.zeros((n.value, 3), np.int32)
        if n.value:
            _check(lib().dbh_batch_copy_candidates(self._h, kind, _ptr(out, _ip), C.byref(n)))
        return out

    def geometry_pairs(self) -> np.ndarray:
        """[P,4] = (scene, flat geom of first body, flat geom of second body, kind)."""
        out = np.zeros((self.info().num_geometry_pairs, 4), np.int32)
        if len(out):
            _check(lib().dbh_batch_copy_geometry_pairs(self._h, _ptr(out, _ip)))
        return out

    def kernel_times_ms(self) -> Dict[str, float]:
        t = (C.c_float * 8)()
        _check(lib().dbh_batch_kernel_times_ms(self._h, t))
        keys = ["pose_root", "traverse", "clip_tet_tri", "clip_tet_tet", "scan", "pack", "surface", "total"]
        return dict(zip(keys, [float(x) for x in t]))

    STAGES = ("pose_root", "traverse", "cull_tet_tri", "cull_tet_tet", "clip_tet_tri", "clip_tet_tet", "half_space",
              "scan", "pack", "surface", "total")

    def stage_times_ms(self) -> Dict[str, float]:
        """Device time of the last query per stage (enable_timing): cull = k_filter + k_small_pairs, clip = k_clip."""
        t = (C.c_float * len(self.STAGES))()
        _check(lib().dbh_batch_stage_times_ms(self._h, t, len(self.STAGES)))
        return dict(zip(self.STAGES, [float(x) for x in t]))

    def stream(self) -> int:
        return int(lib().dbh_batch_stream(self._h) or 0)
