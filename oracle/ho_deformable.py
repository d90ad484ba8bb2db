"""CPU restatement of the deformable x rigid contact data the reference derives from a tet-triangle contact surface.
TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing of the product).

In the reference a deformable body meets a rigid one through the SAME intersector as a hydroelastic soft-rigid pair
(SurfaceVolumeIntersector, mesh_intersection.cc), with the roles the other way round: the volume mesh + pressure field
belong to the RIGID geometry (frame R) and the triangle mesh is the DEFORMABLE body's surface (frame D = World, its
vertices move every step and its Bvh<Aabb> is refit, bvh_updater.h):

  AddDeformableDeformableContactSurface       geometry/proximity/deformable_field_intersection.cc:46-124 (two deforming
                                              volume meshes with signed-distance fields through VolumeIntersector)
  VolumeMesh::CalcBarycentric                 geometry/proximity/volume_mesh.cc:24-52
  AddDeformableRigidContactSurface           geometry/proximity/deformable_mesh_intersection.cc:85-183
  DeformableSurfaceVolumeIntersector          :18-83 (triangle index + barycentric centroid per polygon)
  DeformableContactSurface ctor (rigid)       geometry/query_results/deformable_contact.cc:35-66
  RotationMatrix::MakeFromOneUnitVector       math/rotation_matrix.cc:133-242
  TriangleSurfaceMesh::CalcBarycentric        geometry/proximity/triangle_surface_mesh.cc:53-87

Pinned by the reference's own fixture (geometry/proximity/test/deformable_mesh_intersection_test.cc:106-178: one polygon,
pressure 123, contact point (1/3, 1/3, 0), barycentric (1/3, 1/3, 1/3), vertices {2, 1, 0}; :33-104: BVHs collide, no
contact; :180-262: only the relative pose matters), geometry/proximity/test/deformable_field_intersection_test.cc
(:99-112 no contact; :114-173 one contact point, normal +z, signed distance and barycentric coordinates as the mesh and
field compute them) and math/test/rotation_matrix_test.cc's properties of MakeFromOneUnitVector --
tests/test_oracle_deformable.py.

CalcBarycentric solves its 3 x 2 least-squares system with Eigen's colPivHouseholderQr; here numpy's lstsq. The two
agree to rounding, not bit for bit: values from this file are compared with a tolerance (1e-10 of the triangle's scale),
never bitwise.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from oracle import ho


def make_from_one_unit_vector(u, axis_index: int) -> np.ndarray:
    """math/rotation_matrix.cc:133-242: a right-handed basis whose column `axis_index` is u."""
    u = np.asarray(u, float)
    R = np.zeros((3, 3))
    R[:, axis_index] = u
    i = int(np.argmin(np.abs(u)))          # minCoeff: the first of equal minima, as numpy
    j, k = (i + 1) % 3, (i + 2) % 3
    mag_a_x_u = np.sqrt(1 - u[i] * u[i])
    r = 1 / mag_a_x_u
    s = -r * u[i]
    v, w = np.zeros(3), np.zeros(3)
    v[i], v[j], v[k] = 0, -r * u[k], r * u[j]
    w[i], w[j], w[k] = mag_a_x_u, s * u[j], s * u[k]
    R[:, (axis_index + 1) % 3] = v
    R[:, (axis_index + 2) % 3] = w
    return R


def calc_barycentric(p, v0, v1, v2) -> np.ndarray:
    """triangle_surface_mesh.cc:53-87: least-squares (b1, b2) of [v1 - v0, v2 - v0] b = p - v0; b0 = 1 - b1 - b2."""
    A = np.stack([np.asarray(v1, float) - v0, np.asarray(v2, float) - v0], axis=1)
    sol = np.linalg.lstsq(A, np.asarray(p, float) - v0, rcond=None)[0]
    return np.array([1.0 - sol[0] - sol[1], sol[0], sol[1]])


@dataclass
class DeformableRigidContact:
    """DeformableContactSurface<double> of a deformable (A) x rigid (B) pair, per contact point = per polygon."""
    contact_points_W: np.ndarray        # [F,3] polygon centroids (contact_mesh_W.element_centroid)
    nhats_W: np.ndarray                 # [F,3] from the rigid body into the deformable one: -face_normal
    R_WCs: np.ndarray                   # [F,3,3] MakeFromOneUnitVector(-nhat, z)
    pressures: np.ndarray               # [F] rigid pressure field at the centroid
    pressure_gradients_W: np.ndarray    # [F,3] as the reference computes them: X_WR * grad_R (translation included)
    triangle_index: np.ndarray          # [F] deformable surface triangle of each polygon
    barycentric_centroids: np.ndarray   # [F,3] of the centroid in that triangle
    surface: ho.Surface                 # the contact mesh itself (vertices, face_data, areas, ...), World


def deformable_rigid_contact(rigid_volume: ho.VolumeMesh, rigid_pressure: ho.VolumeField, deformable_surface: ho.TriMesh,
                             X_WR, X_WD=None) -> Optional[DeformableRigidContact]:
    """AddDeformableRigidContactSurface (deformable_mesh_intersection.cc:85-183) on the deformable surface's CURRENT
    vertices. X_WD is the identity in the reference (deformable quantities live in World); None when the intersection
    is empty."""
    X_WD = ho.IDENTITY if X_WD is None else X_WD
    # The intersector's own orientation: M = the volume (rigid body), N = the surface; ids chosen so nothing is swapped.
    surf = ho.soft_rigid(rigid_volume, rigid_pressure, deformable_surface, X_WR, X_WD, id_S=1, id_R=2, rep=ho.KPOLYGON,
                         filter_face_normal=True)
    if surf is None or surf.num_faces == 0:
        return None
    verts_D, tris, normals_D = deformable_surface.arrays()[:3]
    R_WR, p_WR = ho.xform_parts(X_WR)
    R_WD, p_WD = ho.xform_parts(X_WD)
    F = surf.num_faces
    polys = surf.polygons()
    out = DeformableRigidContact(surf.centroids.copy(), -surf.normals, np.zeros((F, 3, 3)), surf.e_centroid.copy(),
                                 np.zeros((F, 3)), surf.elem_second.copy(), np.zeros((F, 3)), surf)
    for f in range(F):
        tri = int(surf.elem_second[f])
        # n_R = R_RD * -face_normal_D(tri); p_RC = CalcPolygonCentroid(polygon, n_R, vertices_R); p_WC = X_DR * p_RC
        # (deformable_mesh_intersection.cc:69-78) -- evaluated here on the World vertices with the World normal: the
        # same point up to the rounding of one rigid transform.
        t = verts_D[tris[tri]]
        n_W = R_WD @ -normals_D[tri]
        p_WC = ho.calc_polygon_centroid(polys[f], n_W, surf.vertices)
        p_DC = R_WD.T @ (p_WC - p_WD)
        out.barycentric_centroids[f] = calc_barycentric(p_DC, t[0], t[1], t[2])
        out.R_WCs[f] = make_from_one_unit_vector(surf.normals[f], 2)      # -nhat = the mesh's face normal
        # pressure_gradient_W[e] = X_WR * pressure_gradient_R[e]: RigidTransform * Vector3 is the POINT transform, so the
        # reference adds p_WR to the rotated gradient (:172-174). Reproduced.
        out.pressure_gradients_W[f] = surf.grad_eM[f] + p_WR
    return out


def participating_vertices(triangle_index, surface_tris, surface_index_to_volume_index) -> List[np.ndarray]:
    """contact_tet_vertex_indices + the participating vertex set (deformable_mesh_intersection.cc:147-170): host-side
    index mapping, one Vector3<int> of VOLUME mesh vertex indices per polygon."""
    tri_vol = np.asarray(surface_index_to_volume_index)[np.asarray(surface_tris)[np.asarray(triangle_index)]]
    return [tri_vol, np.unique(tri_vol)]


def calc_barycentric_tet(p, v) -> np.ndarray:
    """VolumeMesh::CalcBarycentric (volume_mesh.cc:24-52): [1 1 1 1; v0 v1 v2 v3] b = [1; p] (there by partial-pivot LU)."""
    A = np.ones((4, 4))
    A[1:, :] = np.asarray(v, float).T
    return np.linalg.solve(A, np.concatenate([[1.0], np.asarray(p, float)]))


@dataclass
class DeformableDeformableContact:
    """The second DeformableContactSurface constructor's data (deformable_contact.cc:68-100), per polygon. Index 0 is
    the intersector's FIRST field (the reference passes the geometry with the larger id there and stores it as B), index
    1 the second (stored as A)."""
    contact_points_W: np.ndarray        # [F,3]
    nhats_W: np.ndarray                 # [F,3] contact_mesh_W.face_normal: out of field 0's body, into field 1's
    R_WCs: np.ndarray                   # [F,3,3] MakeFromOneUnitVector(-nhat, z)
    signed_distances: np.ndarray        # [F]
    tet0: np.ndarray                    # [F] tet0_of_polygon
    tet1: np.ndarray
    barycentric0: np.ndarray            # [F,4] of the contact point in tet0 of mesh 0
    barycentric1: np.ndarray
    surface: ho.Surface


def deformable_deformable_contact(mesh0: ho.VolumeMesh, sdf0, mesh1: ho.VolumeMesh, sdf1) -> Optional[DeformableDeformableContact]:
    """AddDeformableDeformableContactSurface (deformable_field_intersection.cc:46-124) on the meshes' CURRENT vertices:
    (mesh0, sdf0) is the first argument pair of the reference's function (its geometry has the larger id). sdf*: the
    per-vertex signed distances; the fields are rebuilt on the current meshes as CalcSignedDistanceField does
    (deformable_contact_geometries.cc:87-92)."""
    f0 = ho.VolumeField.from_values(mesh0, sdf0)
    f1 = ho.VolumeField.from_values(mesh1, sdf1)
    surf = ho.soft_soft(mesh0, f0, mesh1, f1, ho.IDENTITY, ho.IDENTITY, 1, 2, ho.KPOLYGON)
    if surf is None or surf.num_faces == 0:
        return None
    F = surf.num_faces
    v0, t0 = mesh0.arrays()[:2]
    v1, t1 = mesh1.arrays()[:2]
    out = DeformableDeformableContact(surf.centroids.copy(), surf.normals.copy(), np.zeros((F, 3, 3)), surf.e_centroid.copy(),
                                      surf.elem_first.copy(), surf.elem_second.copy(), np.zeros((F, 4)), np.zeros((F, 4)), surf)
    for f in range(F):
        out.R_WCs[f] = make_from_one_unit_vector(-surf.normals[f], 2)
        out.barycentric0[f] = calc_barycentric_tet(surf.centroids[f], v0[t0[int(surf.elem_first[f])]])
        out.barycentric1[f] = calc_barycentric_tet(surf.centroids[f], v1[t1[int(surf.elem_second[f])]])
    return out
