"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libhydro_oracle.so (the CPU restatement of the
reference's hydroelastic contact-surface path). Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (drake_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhydro_oracle.so")

KTRIANGLE, KPOLYGON = 0, 1
MODE_BRUTE, MODE_TREE, MODE_REFERENCE_BVH, MODE_REFERENCE_OBB = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compiles the oracle with the recipe in oracle/Makefile."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ho_last_error.restype = C.c_char_p
        _lib.ho_volume_total_volume.restype = C.c_double
        _lib.ho_distance_to_surface.restype = C.c_double
        for name, args in {
            "ho_make_sphere_volume": [C.c_double, C.c_double, C.c_int],
            "ho_make_ellipsoid_volume": [C.c_double] * 4 + [C.c_int],
            "ho_make_box_volume_ma": [C.c_double] * 3,
            "ho_make_box_volume_grid": [C.c_double] * 4,
            "ho_make_cylinder_volume_ma": [C.c_double] * 3,
            "ho_make_capsule_volume": [C.c_double] * 3,
            "ho_cylinder_field": [C.c_int] + [C.c_double] * 4,
            "ho_capsule_field": [C.c_int] + [C.c_double] * 4,
            "ho_mesh_field": [C.c_int, C.c_double, C.c_double],
            "ho_make_convex_volume": [C.c_int],
            "ho_distance_to_surface": [C.c_int, _dp],
            "ho_sphere_field": [C.c_int, C.c_double, C.c_double, C.c_double],
            "ho_ellipsoid_field": [C.c_int] + [C.c_double] * 5,
            "ho_box_field": [C.c_int] + [C.c_double] * 5,
            "ho_volume_total_volume": [C.c_int],
        }.items():
            getattr(_lib, name).argtypes = args
        _lib.ho_soft_rigid.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                       _ip, C.c_int, C.c_int]
        _lib.ho_soft_soft.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int64, C.c_int64, C.c_int,
                                      C.c_int, _ip, C.c_int]
        _lib.ho_surface_bfs_pairs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        _lib.ho_volume_neighbors.argtypes = [C.c_int, _ip]
        _lib.ho_soft_mesh_rigid_half_space.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int64, C.c_int64, C.c_int]
        _lib.ho_soft_half_space_rigid_mesh.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp, C.c_int64, C.c_int64,
                                                       C.c_int, C.c_int]
        _lib.ho_result_from_arrays.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, _ip, _dp, C.c_int64, C.c_int64]
        _lib.ho_traction.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        for name, n in (("ho_atan_x_over_x_from_x_squared", 1), ("ho_combined_dissipation", 4),
                        ("ho_combined_friction", 2)):
            getattr(_lib, name).argtypes = [C.c_double] * n
            getattr(_lib, name).restype = C.c_double
        _lib.ho_scene_add_half_space.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.ho_scene_add_geometry.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.ho_scene_add_undefined.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int]
        _lib.ho_scene_fallback_pairs.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]
        _lib.ho_scene_exclude_pair.argtypes = [C.c_int, C.c_int64, C.c_int64]
    return _lib


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pi(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_ip) if a is not None else None


def _check(h: int, what: str) -> int:
    if h < 0:
        raise RuntimeError(f"oracle {what} failed: {lib().ho_last_error().decode()}")
    return h


IDENTITY = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def xform(R: Optional[np.ndarray] = None, p: Sequence[float] = (0, 0, 0)) -> np.ndarray:
    """RigidTransform as 12 doubles: R column-major then p (math/rigid_transform.h)."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    return np.concatenate([R.T.reshape(-1), np.asarray(p, dtype=np.float64)])


def xform_parts(X: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    X = np.asarray(X, dtype=np.float64)
    return X[:9].reshape(3, 3).T.copy(), X[9:12].copy()


def rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """RollPitchYaw -> rotation matrix: R = Rz(yaw) Ry(pitch) Rx(roll) (math/roll_pitch_yaw.h)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def angle_axis_matrix(angle: float, axis: Sequence[float]) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def compose(X_AB: np.ndarray, X_BC: np.ndarray) -> np.ndarray:
    out = np.zeros(12)
    lib().ho_compose(_pd(_d(X_AB)), _pd(_d(X_BC)), _pd(out))
    return out


def invert_and_compose(X_BA: np.ndarray, X_BC: np.ndarray) -> np.ndarray:
    out = np.zeros(12)
    lib().ho_invert_and_compose(_pd(_d(X_BA)), _pd(_d(X_BC)), _pd(out))
    return out


class VolumeMesh:
    def __init__(self, handle: int):
        self.h = _check(handle, "volume mesh")
        nv, nt = C.c_int(), C.c_int()
        lib().ho_volume_counts(self.h, C.byref(nv), C.byref(nt))
        self.num_vertices, self.num_elements = nv.value, nt.value
        self._cache = None

    @staticmethod
    def sphere(r: float, hint: float, dense: bool = False) -> "VolumeMesh":
        return VolumeMesh(lib().ho_make_sphere_volume(r, hint, int(dense)))

    @staticmethod
    def ellipsoid(a: float, b: float, c: float, hint: float, dense: bool = False) -> "VolumeMesh":
        return VolumeMesh(lib().ho_make_ellipsoid_volume(a, b, c, hint, int(dense)))

    @staticmethod
    def box_ma(x: float, y: float, z: float) -> "VolumeMesh":
        return VolumeMesh(lib().ho_make_box_volume_ma(x, y, z))

    @staticmethod
    def box_grid(x: float, y: float, z: float, hint: float) -> "VolumeMesh":
        return VolumeMesh(lib().ho_make_box_volume_grid(x, y, z, hint))

    @staticmethod
    def cylinder_ma(radius: float, length: float, hint: float) -> "VolumeMesh":
        """MakeCylinderVolumeMeshWithMa (make_cylinder_mesh.cc:325-425)."""
        return VolumeMesh(lib().ho_make_cylinder_volume_ma(radius, length, hint))

    @staticmethod
    def capsule(radius: float, length: float, hint: float) -> "VolumeMesh":
        """MakeCapsuleVolumeMesh (make_capsule_mesh.cc:13-180)."""
        return VolumeMesh(lib().ho_make_capsule_volume(radius, length, hint))

    @staticmethod
    def from_arrays(verts, tets) -> "VolumeMesh":
        v, t = _d(verts).reshape(-1, 3), _i(tets).reshape(-1, 4)
        return VolumeMesh(lib().ho_make_volume(len(v), _pd(v), len(t), _pi(t)))

    def arrays(self):
        """(verts [V,3], tets [T,4], inward_normals [T,4,3])."""
        if self._cache is None:
            v = np.zeros((self.num_vertices, 3))
            t = np.zeros((self.num_elements, 4), dtype=np.int32)
            n = np.zeros((self.num_elements, 4, 3))
            lib().ho_volume_get(self.h, _pd(v), _pi(t), _pd(n))
            self._cache = (v, t, n)
        return self._cache

    @property
    def verts(self):
        return self.arrays()[0]

    @property
    def tets(self):
        return self.arrays()[1]

    def total_volume(self) -> float:
        return lib().ho_volume_total_volume(self.h)

    def neighbors(self) -> np.ndarray:
        """VolumeMeshTopology (volume_mesh_topology.cc:13-68): [e, f] = the tetrahedron across face f of e, -1 outside."""
        out = np.zeros((len(self.tets), 4), dtype=np.int32)
        _check(lib().ho_volume_neighbors(self.h, _pi(out)), "volume_neighbors")
        return out

    def to_surface(self) -> "TriMesh":
        return TriMesh(lib().ho_volume_to_surface(self.h, None))

    def element_boxes(self):
        c, hw = np.zeros((self.num_elements, 3)), np.zeros((self.num_elements, 3))
        lib().ho_element_boxes(0, self.h, _pd(c), _pd(hw))
        return c, hw

    def bvh_stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().ho_bvh_stats(0, self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class TriMesh:
    def __init__(self, handle: int):
        self.h = _check(handle, "triangle mesh")
        nv, nt = C.c_int(), C.c_int()
        lib().ho_tri_counts(self.h, C.byref(nv), C.byref(nt))
        self.num_vertices, self.num_elements = nv.value, nt.value
        self._cache = None

    @staticmethod
    def from_arrays(verts, tris) -> "TriMesh":
        v, t = _d(verts).reshape(-1, 3), _i(tris).reshape(-1, 3)
        return TriMesh(lib().ho_make_tri(len(v), _pd(v), len(t), _pi(t)))

    def to_convex_volume(self) -> "VolumeMesh":
        """MakeConvexVolumeMesh (make_convex_mesh.cc:17-60): this closed surface's vertices + the centroid of its volume."""
        return VolumeMesh(lib().ho_make_convex_volume(self.h))

    def distance_to(self, p) -> float:
        """CalcDistanceToSurfaceMesh (calc_distance_to_surface_mesh.cc:30-200)."""
        return float(lib().ho_distance_to_surface(self.h, _pd(_d(p))))

    def update_vertices(self, verts):
        """DeformableMeshWithBvh::UpdateVertexPositions (deformable_mesh_with_bvh.h:86-93): new vertex positions, derived
        quantities recomputed, BVHs refitted with their topology kept (BvhUpdater::Update, bvh_updater.h:45-103)."""
        v = _d(verts).reshape(-1, 3)
        assert len(v) == self.num_vertices
        _check(lib().ho_tri_update_vertices(self.h, _pd(v)), "tri_update_vertices")
        self._cache = None

    def bvh_is_fitted(self) -> bool:
        """Every box of the reference-shaped BVH contains its triangles and the root is the mesh's own box."""
        return bool(lib().ho_tri_bvh_is_fitted(self.h))

    @staticmethod
    def sphere(r: float, hint: float) -> "TriMesh":
        return VolumeMesh.sphere(r, hint, dense=False).to_surface()

    @staticmethod
    def ellipsoid(a: float, b: float, c: float, hint: float) -> "TriMesh":
        return VolumeMesh.ellipsoid(a, b, c, hint, dense=False).to_surface()

    @staticmethod
    def box(x: float, y: float, z: float) -> "TriMesh":
        return VolumeMesh.box_ma(x, y, z).to_surface()

    def arrays(self):
        """(verts [V,3], tris [T,3], normals [T,3], areas [T], (total_area, centroid[3]))."""
        if self._cache is None:
            v = np.zeros((self.num_vertices, 3))
            t = np.zeros((self.num_elements, 3), dtype=np.int32)
            n = np.zeros((self.num_elements, 3))
            a = np.zeros(self.num_elements)
            tot = np.zeros(4)
            lib().ho_tri_get(self.h, _pd(v), _pi(t), _pd(n), _pd(a), _pd(tot))
            self._cache = (v, t, n, a, tot)
        return self._cache

    @property
    def verts(self):
        return self.arrays()[0]

    @property
    def tris(self):
        return self.arrays()[1]

    def element_boxes(self):
        c, hw = np.zeros((self.num_elements, 3)), np.zeros((self.num_elements, 3))
        lib().ho_element_boxes(1, self.h, _pd(c), _pd(hw))
        return c, hw

    def bvh_stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().ho_bvh_stats(1, self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class VolumeField:
    def __init__(self, handle: int, mesh: VolumeMesh):
        self.h = _check(handle, "field")
        self.mesh = mesh
        self._cache = None

    @staticmethod
    def sphere(mesh: VolumeMesh, r: float, modulus: float, margin: float = 0.0) -> "VolumeField":
        return VolumeField(lib().ho_sphere_field(mesh.h, r, modulus, margin), mesh)

    @staticmethod
    def ellipsoid(mesh: VolumeMesh, a: float, b: float, c: float, modulus: float, margin: float = 0.0):
        return VolumeField(lib().ho_ellipsoid_field(mesh.h, a, b, c, modulus, margin), mesh)

    @staticmethod
    def box(mesh: VolumeMesh, x: float, y: float, z: float, modulus: float, margin: float = 0.0):
        return VolumeField(lib().ho_box_field(mesh.h, x, y, z, modulus, margin), mesh)

    @staticmethod
    def cylinder(mesh: VolumeMesh, radius: float, length: float, modulus: float, margin: float = 0.0):
        """MakeCylinderPressureField (make_cylinder_field.cc:18-75)."""
        return VolumeField(lib().ho_cylinder_field(mesh.h, radius, length, modulus, margin), mesh)

    @staticmethod
    def capsule(mesh: VolumeMesh, radius: float, length: float, modulus: float, margin: float = 0.0):
        """MakeCapsulePressureField (make_capsule_field.h:55-90)."""
        return VolumeField(lib().ho_capsule_field(mesh.h, radius, length, modulus, margin), mesh)

    @staticmethod
    def from_mesh(mesh: VolumeMesh, modulus: float, margin: float = 0.0):
        """MakeVolumeMeshPressureField (make_mesh_field.cc:45-96)."""
        return VolumeField(lib().ho_mesh_field(mesh.h, modulus, margin), mesh)

    @staticmethod
    def from_values(mesh: VolumeMesh, values) -> "VolumeField":
        v = _d(values)
        assert len(v) == mesh.num_vertices
        return VolumeField(lib().ho_make_field(mesh.h, _pd(v)), mesh)

    def arrays(self):
        """(values [V], gradients [T,3], values_at_Mo [T], degenerate)."""
        if self._cache is None:
            v = np.zeros(self.mesh.num_vertices)
            g = np.zeros((self.mesh.num_elements, 3))
            m = np.zeros(self.mesh.num_elements)
            deg = lib().ho_field_get(self.h, _pd(v), _pd(g), _pd(m))
            self._cache = (v, g, m, bool(deg))
        return self._cache

    @property
    def values(self):
        return self.arrays()[0]


@dataclass
class Surface:
    """Flattened ContactSurface<double> (query_results/contact_surface.h)."""
    id_M: int
    id_N: int
    rep: int
    vertices: np.ndarray
    pressures: np.ndarray
    face_data: np.ndarray  # polygon: |n|v..| ... ; triangle: [F,3]
    areas: np.ndarray
    normals: np.ndarray
    centroids: np.ndarray
    e_centroid: np.ndarray
    field_grad: Optional[np.ndarray]
    field_at_Wo: Optional[np.ndarray]
    grad_eM: Optional[np.ndarray]
    grad_eN: Optional[np.ndarray]
    total_area: float
    centroid: np.ndarray
    elem_first: np.ndarray
    elem_second: np.ndarray
    handle: Optional[int] = None  # oracle-side result kept alive (inside `with keep_handles():`) for traction()

    @property
    def num_faces(self) -> int:
        return len(self.areas)

    def polygons(self) -> List[np.ndarray]:
        """Per-face vertex index arrays."""
        if self.rep == KTRIANGLE:
            return [t for t in self.face_data]
        out, i = [], 0
        fd = self.face_data
        while i < len(fd):
            n = fd[i]
            out.append(fd[i + 1:i + 1 + n])
            i += n + 1
        return out


_KEEP_HANDLES = False


class keep_handles:
    """Context manager: surfaces fetched inside keep their oracle-side result (Surface.handle) so that
    traction() can run on exactly the arrays the oracle computed."""

    def __enter__(self):
        global _KEEP_HANDLES
        self._old, _KEEP_HANDLES = _KEEP_HANDLES, True
        return self

    def __exit__(self, *exc):
        global _KEEP_HANDLES
        _KEEP_HANDLES = self._old
        return False


def _fetch_result_keep(h: int) -> "Surface":
    with keep_handles():
        return _fetch_result(h)


def _fetch_result(h: int) -> Surface:
    L = lib()
    ids = (C.c_int64 * 2)()
    info = (C.c_int * 8)()
    L.ho_result_info(h, ids, info)
    rep, nf, nv, nfd, hm, hn = info[0], info[1], info[2], info[3], info[4], info[5]
    verts = np.zeros((nv, 3))
    pres = np.zeros(nv)
    fd = np.zeros(nfd, dtype=np.int32)
    areas = np.zeros(nf)
    normals, cents = np.zeros((nf, 3)), np.zeros((nf, 3))
    ec = np.zeros(nf)
    fg = np.zeros((nf, 3)) if rep == KPOLYGON else None
    fo = np.zeros(nf) if rep == KPOLYGON else None
    gm = np.zeros((nf, 3)) if hm else None
    gn = np.zeros((nf, 3)) if hn else None
    tot = np.zeros(4)
    e0 = np.zeros(nf, dtype=np.int32)
    e1 = np.zeros(nf, dtype=np.int32)
    L.ho_result_get(h, _pd(verts), _pd(pres), _pi(fd), _pd(areas), _pd(normals), _pd(cents), _pd(ec), _pd(fg), _pd(fo),
                    _pd(gm), _pd(gn), _pd(tot), _pi(e0), _pi(e1))
    if not _KEEP_HANDLES:
        L.ho_result_free(h)
    if rep == KTRIANGLE:
        fd = fd.reshape(-1, 3)
    return Surface(ids[0], ids[1], rep, verts, pres, fd, areas, normals, cents, ec, fg, fo, gm, gn, float(tot[0]),
                   tot[1:4].copy(), e0, e1, h if _KEEP_HANDLES else None)


def soft_rigid(mesh_S: VolumeMesh, field_S: VolumeField, mesh_R: TriMesh, X_WS, X_WR, id_S: int = 1, id_R: int = 2,
               rep: int = KPOLYGON, mode: int = MODE_BRUTE, pairs=None,
               filter_face_normal: bool = True) -> Optional[Surface]:
    """ComputeContactSurfaceFromCompliantVolumeRigidSurface (mesh_intersection.cc:384-440).

    filter_face_normal=False reproduces SampleVolumeFieldOnSurface(...,
    filter_face_normal_along_field_gradient=false) (mesh_intersection.h:155-169)."""
    p = _i(pairs).reshape(-1, 2) if pairs is not None else None
    h = lib().ho_soft_rigid(mesh_S.h, field_S.h, mesh_R.h, _pd(_d(X_WS)), _pd(_d(X_WR)), id_S, id_R, rep, mode,
                            _pi(p), 0 if p is None else len(p), int(filter_face_normal))
    if h == -1:
        return None
    return _fetch_result(_check(h, "soft_rigid"))


def soft_soft(mesh_M: VolumeMesh, field_M: VolumeField, mesh_N: VolumeMesh, field_N: VolumeField, X_WM, X_WN,
              id_M: int = 1, id_N: int = 2, rep: int = KPOLYGON, mode: int = MODE_BRUTE, pairs=None) -> Optional[Surface]:
    """ComputeContactSurfaceFromCompliantVolumes (field_intersection.cc:500-520)."""
    p = _i(pairs).reshape(-1, 2) if pairs is not None else None
    h = lib().ho_soft_soft(mesh_M.h, field_M.h, mesh_N.h, field_N.h, _pd(_d(X_WM)), _pd(_d(X_WN)), id_M, id_N, rep,
                           mode, _pi(p), 0 if p is None else len(p))
    if h == -1:
        return None
    return _fetch_result(_check(h, "soft_soft"))


def soft_mesh_rigid_half_space(mesh_S: VolumeMesh, field_S: VolumeField, X_WS, X_WR, id_S: int = 1, id_R: int = 2,
                               rep: int = KPOLYGON) -> Optional[Surface]:
    """ComputeContactSurfaceFromCompliantVolumeRigidHalfSpace (mesh_plane_intersection.cc:251-342): the half
    space is z <= 0 of frame R. Faces carry (tet, -1) as their emitting element pair."""
    h = lib().ho_soft_mesh_rigid_half_space(mesh_S.h, field_S.h, _pd(_d(X_WS)), _pd(_d(X_WR)), id_S, id_R, rep)
    if h == -1:
        return None
    return _fetch_result(_check(h, "soft_mesh_rigid_half_space"))


def soft_half_space_rigid_mesh(mesh_R: TriMesh, X_WS, pressure_scale: float, X_WR, id_S: int = 1, id_R: int = 2,
                               rep: int = KPOLYGON, margin: float = 0.0, filter_normals: bool = True) -> Optional[Surface]:
    """ComputeContactSurfaceFromCompliantHalfSpaceRigidMesh (mesh_half_space_intersection.cc:338-418): the
    compliant half space is z <= 0 of frame S. Faces carry (-1, triangle) as their emitting element pair.
    filter_normals=False skips the BVH callback's face-normal filter (ComputeContactSurface over every triangle,
    :264-336, which is what the reference's MeshHalfSpaceValueTest cases exercise)."""
    h = lib().ho_soft_half_space_rigid_mesh(mesh_R.h, _pd(_d(X_WS)), pressure_scale, margin, _pd(_d(X_WR)), id_S, id_R,
                                            rep, int(filter_normals))
    if h == -1:
        return None
    return _fetch_result(_check(h, "soft_half_space_rigid_mesh"))


# ---- the contact-surface consumer of SURVEY 8(f) rank 3 (hydroelastic_traction_calculator.cc) -------------

def surface_from_arrays(rep: int, verts, pressures, face_data, grads=None, id_M: int = 1, id_N: int = 2):
    """A ContactSurface from explicit World-frame arrays (polygon: face_data |n|v..| + one gradient per face;
    triangle: [F,3] indices). The returned Surface keeps its handle for traction()."""
    v, p, fd = _d(verts).reshape(-1, 3), _d(pressures).reshape(-1), _i(face_data).reshape(-1)
    g = _d(grads).reshape(-1, 3) if grads is not None else None
    h = _check(lib().ho_result_from_arrays(rep, len(v), _pd(v), _pd(p), len(fd), _pi(fd), _pd(g), id_M, id_N),
               "result_from_arrays")
    return _fetch_result_keep(h)


def traction(surface: Surface, p_WAo, p_WBo, V_WA, V_WB, dissipation: float, mu_coulomb: float,
             vslip_regularizer: float = 1e-6):
    """HydroelasticTractionCalculator: (F_Ac_W, F_Ao_W, F_Bo_W), each [6] = (tau, f); V = (w, v); A / B are the
    bodies of geometries M / N. `surface` must have been fetched inside `with keep_handles():`."""
    assert surface.handle is not None, "fetch the surface inside `with ho.keep_handles():`"
    out = np.zeros(18)
    rc = lib().ho_traction(surface.handle, _pd(_d(p_WAo)), _pd(_d(p_WBo)), _pd(_d(V_WA)), _pd(_d(V_WB)), dissipation,
                           mu_coulomb, vslip_regularizer, _pd(out))
    _check(rc, "traction")
    return out[0:6], out[6:12], out[12:18]


def discrete_pairs(surface: Surface) -> np.ndarray:
    """The geometry-only half of DiscreteUpdateManager::AppendDiscreteContactPairsForHydroelasticContact
    (multibody/plant/discrete_update_manager.cc:896-1036), face by face (plain loops: test sizes only): [F,9] = p_WC,
    nhat_BA_W, phi0, fn0, stiffness; a face the reference skips (:909-919, :941-949) keeps stiffness 0."""
    out = np.zeros((surface.num_faces, 9))
    for f in range(surface.num_faces):
        Ae = float(surface.areas[f])
        nhat = surface.normals[f]
        out[f, 0:3] = surface.centroids[f]
        out[f, 3:6] = nhat
        if not Ae > 1.0e-14:
            continue
        gM = float(np.dot(surface.grad_eM[f], nhat)) if surface.grad_eM is not None else math.inf
        gN = -float(np.dot(surface.grad_eN[f], nhat)) if surface.grad_eN is not None else math.inf
        if gM < 1.0e-14 or gN < 1.0e-14:
            continue
        g = 1.0 / (1.0 / gM + 1.0 / gN)
        p0 = float(surface.e_centroid[f])
        out[f, 6] = -p0 / g
        out[f, 7] = Ae * p0
        out[f, 8] = Ae * g
    return out


def atan_x_over_x_from_x_squared(x2: float) -> float:
    return lib().ho_atan_x_over_x_from_x_squared(x2)


def combined_dissipation(stiffness_A, stiffness_B, dA, dB) -> float:
    return lib().ho_combined_dissipation(stiffness_A, stiffness_B, dA, dB)


def combined_friction(mu_A, mu_B) -> float:
    return lib().ho_combined_friction(mu_A, mu_B)


def candidates(mesh_A: VolumeMesh, mesh_B, X_AB, mode: int = MODE_BRUTE) -> Tuple[np.ndarray, int]:
    """Candidate element pairs [(a, b)] and the number of BV tests (mode 2 only)."""
    kind = 0 if isinstance(mesh_B, VolumeMesh) else 1
    tests = C.c_int64()
    h = _check(lib().ho_candidates(mesh_A.h, kind, mesh_B.h, _pd(_d(X_AB)), mode, C.byref(tests)), "candidates")
    n = lib().ho_pairs_count(h)
    out = np.zeros((n, 2), dtype=np.int32)
    lib().ho_pairs_get(h, _pi(out))
    lib().ho_pairs_free(h)
    return out, tests.value


def surface_bfs_pairs(mesh_M: VolumeMesh, field_M: VolumeField, mesh_N: VolumeMesh, field_N: VolumeField,
                      X_MN) -> np.ndarray:
    """The (tet of M, tet of N) pairs the surface-BVH + BFS overload of VolumeIntersector::IntersectFields hands to
    CalcContactPolygon, in visit order (field_intersection.cc:254-375). soft_soft(..., pairs=these) is its surface."""
    h = _check(lib().ho_surface_bfs_pairs(mesh_M.h, field_M.h, mesh_N.h, field_N.h, _pd(_d(X_MN))), "surface_bfs_pairs")
    n = lib().ho_pairs_count(h)
    out = np.zeros((n, 2), dtype=np.int32)
    lib().ho_pairs_get(h, _pi(out))
    lib().ho_pairs_free(h)
    return out


RIGID, COMPLIANT = 1, 2


class Scene:
    """ProximityEngine-shaped scene for the oracle (proximity_engine.cc:905-934)."""

    def __init__(self):
        self.h = lib().ho_scene_create()
        self._keep = []

    def add_rigid(self, gid: int, mesh: TriMesh, anchored: bool = False):
        self._keep.append(mesh)
        lib().ho_scene_add_geometry(self.h, gid, RIGID, -1, -1, mesh.h, int(anchored))

    def add_compliant(self, gid: int, mesh: VolumeMesh, fld: VolumeField, anchored: bool = False):
        self._keep += [mesh, fld]
        lib().ho_scene_add_geometry(self.h, gid, COMPLIANT, mesh.h, fld.h, -1, int(anchored))

    def add_half_space(self, gid: int, compliant: bool = False, pressure_scale: float = 0.0, margin: float = 0.0,
                       anchored: bool = False):
        """HalfSpace shape (z <= 0 of its frame); compliant: pressure_scale = modulus / slab thickness."""
        lib().ho_scene_add_half_space(self.h, gid, COMPLIANT if compliant else RIGID, pressure_scale, margin,
                                      int(anchored))

    def add_undefined(self, gid: int, box=None, anchored: bool = False):
        """A geometry with no hydroelastic representation; box = (centre, half widths) of its shape in its frame, or
        None when unknown (then the geometry-level broad phase reports every pair with it)."""
        b = None if box is None else _pd(_d(box).reshape(6))
        lib().ho_scene_add_undefined(self.h, gid, b, int(anchored))

    def fallback_pairs(self, poses) -> np.ndarray:
        """[n,2] id pairs a WithFallback query hands to the point-pair narrow phase (proximity_engine.cc:958-970)."""
        P = _d(poses).reshape(-1, 12)
        assert len(P) == lib().ho_scene_num_geometries(self.h)
        cap = len(P) * (len(P) - 1) // 2 + 1
        out = np.zeros((cap, 2), np.int64)
        n = lib().ho_scene_fallback_pairs(self.h, _pd(P), out.ctypes.data_as(_lp), cap)
        _check(n, "scene_fallback_pairs")
        return out[:n]

    def set_normal_filter(self, on: bool = True):
        """Rigid-compliant mesh pairs with (default) or without IsFaceNormalAlongPressureGradient
        (SampleVolumeFieldOnSurface's filter_face_normal_along_field_gradient, mesh_intersection.h)."""
        lib().ho_scene_set_normal_filter(self.h, int(on))

    def exclude(self, a: int, b: int):
        lib().ho_scene_exclude_pair(self.h, a, b)

    def ids(self) -> np.ndarray:
        n = lib().ho_scene_num_geometries(self.h)
        out = np.zeros(n, dtype=np.int64)
        lib().ho_scene_geometry_ids(self.h, out.ctypes.data_as(_lp))
        return out

    def compute(self, poses, rep: int = KPOLYGON, mode: int = MODE_BRUTE, with_fallback: bool = False):
        """poses: [G,12] ordered by ascending geometry id. Returns (surfaces, stats dict)."""
        P = _d(poses).reshape(-1, 12)
        assert len(P) == lib().ho_scene_num_geometries(self.h)
        cap = len(P) * (len(P) - 1) // 2 + 1
        handles = (C.c_int * cap)()
        stats = (C.c_int64 * 5)()
        n = lib().ho_scene_compute(self.h, _pd(P), rep, mode, int(with_fallback), handles, cap, stats)
        if n == -3:
            raise RuntimeError("rigid-rigid contact is not supported by the hydroelastic model")
        if n == -2:
            raise RuntimeError("geometry without a hydroelastic representation")
        if n == -4:
            raise RuntimeError("contact between two half spaces is not allowed")
        if n == -5:
            raise RuntimeError("contact between two compliant geometries, one of which is a half space, is not allowed")
        _check(n, "scene_compute")
        surfaces = [_fetch_result(handles[k]) for k in range(n)]
        keys = ["geometry_pairs", "bv_tests", "element_pairs", "faces", "surfaces"]
        return surfaces, dict(zip(keys, [int(s) for s in stats]))

    def batch_time(self, poses, rep: int = KPOLYGON, mode: int = MODE_REFERENCE_BVH, nthreads: int = 1):
        """Times nscenes instances; poses [S,G,12]. Returns (seconds, stats dict)."""
        P = _d(poses)
        G = lib().ho_scene_num_geometries(self.h)
        P = P.reshape(-1, G, 12)
        sec = C.c_double()
        stats = (C.c_int64 * 5)()
        area = C.c_double()
        rc = lib().ho_scene_batch_time(self.h, len(P), _pd(P), rep, mode, nthreads, C.byref(sec), stats, C.byref(area))
        if rc != 0:
            raise RuntimeError(f"oracle batch failed rc={rc}: {lib().ho_last_error().decode()}")
        keys = ["geometry_pairs", "bv_tests", "element_pairs", "faces", "surfaces"]
        d = dict(zip(keys, [int(s) for s in stats]))
        d["total_area"] = area.value
        return sec.value, d


# ---- low-level functions for the known-answer tests -----------------------

def clip_polygon_by_half_space(verts, normal, point) -> np.ndarray:
    v = _d(verts).reshape(-1, 3)
    out = np.zeros((2 * len(v) + 2, 3))
    n = lib().ho_clip_polygon_by_half_space(len(v), _pd(v), _pd(_d(normal)), _pd(_d(point)), _pd(out))
    return out[:n]


def calc_intersection(A, B, normal, point) -> np.ndarray:
    out = np.zeros(3)
    lib().ho_calc_intersection(_pd(_d(A)), _pd(_d(B)), _pd(_d(normal)), _pd(_d(point)), _pd(out))
    return out


def remove_nearly_duplicate_vertices(verts) -> np.ndarray:
    v = _d(verts).reshape(-1, 3)
    out = np.zeros((len(v), 3))
    n = lib().ho_remove_nearly_duplicate_vertices(len(v), _pd(v), _pd(out))
    return out[:n]


def clip_triangle_by_tetrahedron(vol: VolumeMesh, tet: int, tri_mesh: TriMesh, tri: int, X_MN) -> np.ndarray:
    out = np.zeros((8, 3))
    n = _check(lib().ho_clip_triangle_by_tetrahedron(vol.h, tet, tri_mesh.h, tri, _pd(_d(X_MN)), _pd(out)), "clip")
    return out[:n]


def is_face_normal_along_pressure_gradient(fld: VolumeField, tri_mesh: TriMesh, X_MN, tet: int, tri: int) -> bool:
    return bool(lib().ho_is_face_normal_along_pressure_gradient(fld.h, tri_mesh.h, _pd(_d(X_MN)), tet, tri))


def calc_equilibrium_plane(f0: VolumeField, e0: int, f1: VolumeField, e1: int, X_MN):
    out = np.zeros(4)
    ok = lib().ho_calc_equilibrium_plane(f0.h, e0, f1.h, e1, _pd(_d(X_MN)), _pd(out))
    return (out[:3].copy(), float(out[3])) if ok else None


def intersect_tetrahedra(m0: VolumeMesh, e0: int, m1: VolumeMesh, e1: int, X_MN, nhat, displacement):
    pl = _d(list(nhat) + [displacement])
    verts = np.zeros((16, 3))
    faces = np.zeros(16, dtype=np.int32)
    n = _check(lib().ho_intersect_tetrahedra(m0.h, e0, m1.h, e1, _pd(_d(X_MN)), _pd(pl), _pd(verts), _pi(faces)),
               "intersect_tetrahedra")
    return verts[:n], faces[:n]


def is_plane_normal_along_pressure_gradient(nhat, tet: int, fld: VolumeField) -> bool:
    return bool(lib().ho_is_plane_normal_along_pressure_gradient(_pd(_d(nhat)), tet, fld.h))


def calc_polygon_centroid(polygon, normal, verts) -> np.ndarray:
    p = _i(polygon)
    v = _d(verts).reshape(-1, 3)
    out = np.zeros(3)
    lib().ho_calc_polygon_centroid(len(p), _pi(p), _pd(_d(normal)), len(v), _pd(v), _pd(out))
    return out


def boxes_overlap(half_a, half_b, X_AB) -> bool:
    return bool(lib().ho_boxes_overlap(_pd(_d(half_a)), _pd(_d(half_b)), _pd(_d(X_AB))))
