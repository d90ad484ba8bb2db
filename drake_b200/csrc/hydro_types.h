// Plain-old-data records shared by host and device code of the B200
// hydroelastic back end. Layouts are part of the HBM data-layout contract
// described in DESIGN.md ("Data layout in HBM").
#pragma once
#include <cstdint>

namespace dbh {

// One tetrahedron of a compliant mesh, self-contained so a narrow-phase thread
// gathers exactly one 128-byte line per tet (no vertex-index indirection).
// Replaces VolumeMesh<double> vertices/elements (volume_mesh.h:344-356) +
// MeshFieldLinear gradients_/values_at_Mo_ (mesh_field_linear.h:513-533).
struct alignas(128) TetRec {
  double v[4][3];  // vertex positions in the mesh frame M
  double grad[3];  // grad of the linear pressure on this tet, in M
  double p_mo;     // value of that linear function at the origin of M
};
static_assert(sizeof(TetRec) == 128, "TetRec must be one 128 B line");

// Unit inward normals of the 4 faces (face f is opposite vertex f):
// VolumeMesh::inward_normals_M_ (volume_mesh.cc:113-136). Only the second
// tetrahedron of a tet-tet pair reads these.
struct alignas(32) TetNormals {
  double n[4][3];
};
static_assert(sizeof(TetNormals) == 96, "TetNormals is 96 B");

// One triangle of a rigid surface mesh: vertices + unit face normal
// (triangle_surface_mesh.h:452-487).
struct alignas(32) TriRec {
  double v[3][3];
  double n[3];
};
static_assert(sizeof(TriRec) == 96, "TriRec is 96 B");

// BVH node: axis-aligned box in the mesh frame (center + half width, the
// reference's Aabb, aabb.h:59-65) and two links. For an internal node
// left/right are node indices; for a leaf left = ~element index (negative) and
// right = -1.
struct alignas(64) BvhNode {
  double c[3];
  double h[3];
  int32_t left;
  int32_t right;
  int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode is 64 B");

// Device-side view of a registered mesh asset.
struct AssetDev {
  const TetRec* tets;          // compliant only
  const TetNormals* normals;   // compliant only
  const TriRec* tris;          // rigid only
  const BvhNode* nodes;
  int32_t num_elements;
  int32_t num_nodes;
  int32_t type;                // DBH_RIGID / DBH_COMPLIANT
  int32_t height;
};

enum PairKind : int32_t { kSoftRigid = 0, kSoftSoft = 1, kRigidRigid = 2, kUnsupported = 3 };

// Static description of one potential geometry pair (one entry of
// FindCollisionCandidates() after filtering, proximity_engine.cc:867-883, in
// the canonical order of MaybeMakeContactSurface, hydroelastic_calculator.cc:77-156).
struct PairRec {
  int32_t scene;
  int32_t geom_a;    // flat geometry index of the first-constructed body (S or M)
  int32_t geom_b;    // second body (R or N)
  int32_t asset_a;
  int32_t asset_b;
  int32_t kind;      // PairKind
  int32_t swapped;   // soft-rigid with id_R < id_S: ContactSurface swaps M and N (contact_surface.cc:39-41)
  int32_t pad;
  int64_t id_m;      // final ids, id_m < id_n
  int64_t id_n;
};
static_assert(sizeof(PairRec) == 48, "PairRec is 48 B");

// Rigid transform as 12 doubles: R column-major then p (math/rigid_transform.h:1004-1007).
struct Xf {
  double m[12];
};

// Device counters of one query (zeroed by the first kernel).
struct Counters {
  unsigned long long frontier[64];  // entries per traversal level (== BV-pair tests per level)
  unsigned long long cand_tri;      // tet-tri candidate pairs
  unsigned long long cand_tet;      // tet-tet candidate pairs
  unsigned long long polys;         // contact polygons in the staging soup
  unsigned long long poly_verts;    // sum of polygon vertex counts
  unsigned long long root_overlaps; // geometry pairs whose root boxes overlap
  unsigned long long surfaces;
  unsigned long long faces;
  unsigned long long vertices;
  int32_t error;                    // first device-side dbh_status raised
  int32_t error_detail;
  double total_area;
};

constexpr int kMaxPolyVerts = 8;    // field_intersection.cc:100-104; 7 for tet-tri (mesh_intersection.cc:261-262)

}  // namespace dbh
