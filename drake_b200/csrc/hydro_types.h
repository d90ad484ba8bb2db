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

// Pose-independent derived data of one tetrahedron, computed once at
// registration with the reference's own expressions so the per-query kernels
// reuse the bits the reference would recompute for every element pair:
//  - plane_nh/plane_d: the four face half spaces of ClipTriangleByTetrahedron
//    (mesh_intersection.cc:225-238: n = (B-A)x(C-A) over the face table
//    {1,2,3},{0,3,2},{0,1,3},{0,2,1}; Plane ctor normalises, plane.cc:16-29;
//    d = nhat.A). plane_d[f] is NaN when |n| < 1e-10 (the Plane ctor throws).
//  - inward: unit inward normals, face f opposite vertex f
//    (VolumeMesh::inward_normals_M_, volume_mesh.cc:113-136); only the second
//    tetrahedron of a tet-tet pair reads these.
//  - ngrad: normalized(grad) used by both gradient filters
//    (mesh_intersection.cc:286, field_intersection.cc:204-205).
//  - amp: shape factor that scales the safety margin of the conservative culls
//    (not part of the arithmetic contract).
struct alignas(128) TetAux {
  double plane_nh[4][3];
  double plane_d[4];
  double inward[4][3];
  double ngrad[3];
  double amp;   // longest edge / inradius (>= ~4.9; huge for slivers): how far, in units of a plane
                // offset, the tetrahedron bounded by offset planes can reach beyond this one
};
static_assert(sizeof(TetAux) == 256, "TetAux is two 128 B lines");

// fp32 copies for the conservative cull stage of the narrow phase (never used
// for results): one 64 B line of vertices + field, one of outward unit face
// normals (face f opposite vertex f; its plane passes through vertex (f+1)%4).
struct alignas(64) TetCullA {
  float v[4][3];
  float grad[3];
  float p_mo;
};
struct alignas(64) TetCullB {
  float nh[4][3];
  float inv_grad_len;  // 1 / |grad| (0 for a zero gradient)
  float margin;        // relative safety margin of the hull culls: max(2e-5, 1e-11 (1 + amp))
  float pad[2];
};
struct alignas(16) TriCull {
  float v[3][3];
  float n[3];
};
static_assert(sizeof(TetCullA) == 64 && sizeof(TetCullB) == 64 && sizeof(TriCull) == 48, "cull records");

// One triangle of a rigid surface mesh: vertices + unit face normal
// (triangle_surface_mesh.h:452-487).
struct alignas(32) TriRec {
  double v[3][3];
  double n[3];
};
static_assert(sizeof(TriRec) == 96, "TriRec is 96 B");

// BVH node, exact copy: axis-aligned box in the mesh frame (center + half
// width, the reference's Aabb, aabb.h:59-65). All assets' nodes live in one
// device pool. Node 0 of an asset is unused, node 1 is the root, and the two
// children of an internal node are adjacent (left = first child, right =
// left + 1, both POOL indices); for a leaf left = ~(pool element index) and
// right = -1. Only the exact leaf predicate and the root test read these.
struct alignas(64) BvhNode {
  double c[3];
  double h[3];
  int32_t left;
  int32_t right;
  int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode is 64 B");

// Exact axis-aligned box of ONE element in its mesh frame (the leaf's Aabb, aabb.cc:102-117),
// indexed by pool element: the exact leaf predicate reads these when fp32 cannot decide.
struct alignas(16) ElemBox {
  double c[3];
  double h[3];
};
static_assert(sizeof(ElemBox) == 48, "ElemBox is 48 B");

// The same node for the conservative fp32 descent: the box is rounded outward
// (h32 >= h + |c - c32|, so it contains the exact box), 32 B, siblings share one
// 64 B line. child >= 0: pool index of the first child; child < 0: ~(pool element).
struct alignas(32) Node32 {
  float c[3];
  float h[3];
  int32_t child;
  float l1;   // |c|_1 + |h|_1: the scale of this box's contribution to the descent's safety pad
};
static_assert(sizeof(Node32) == 32, "Node32 is 32 B");

// Per geometry pair, fp32 copy of X_MN for the conservative descent.
struct alignas(32) PairF32 {
  float R[9];    // column-major
  float AR[9];   // |R| + 1e-6 (row-major: AR[i*3+j] = |R(i,j)| + eps, as in boxes_overlap.cc)
  float p[3];
  float S;       // |p|_1
  float pad[2];
};
static_assert(sizeof(PairF32) == 96, "PairF32 is 96 B");

// The element / node pools of a context: every registered asset occupies a
// contiguous range of each, so work-list entries carry pool indices and the
// kernels never chase pair -> asset -> array pointers.
struct Pools {
  const TetRec* tets;
  const TetAux* aux;
  const TriRec* tris;
  const BvhNode* nodes;
  const Node32* nodes32;
  const TetCullA* cull_a;
  const TetCullB* cull_b;
  const TriCull* cull_tri;
  const ElemBox* ebox_tet;   // per pool tetrahedron
  const ElemBox* ebox_tri;   // per pool triangle
};

// 0-3: mesh pairs and the failures of MaybeMakeContactSurface that need no shape knowledge. 4-8: pairs
// with a HalfSpace (hydroelastic_calculator.cc:25-41,113-123,145-147): the two supported special cases
// (k_plane_*), and the combinations the reference rejects (ContactSurfaceResult::kHalfSpaceHalfSpace,
// kCompliantHalfSpaceCompliantMesh, kRigidRigid with a half space: no root boxes to test).
enum PairKind : int32_t {
  kSoftRigid = 0,
  kSoftSoft = 1,
  kRigidRigid = 2,
  kUnsupported = 3,
  kSoftMeshRigidPlane = 4,   // compliant volume mesh (a) x rigid half space (b)
  kSoftPlaneRigidMesh = 5,   // compliant half space (a) x rigid surface mesh (b)
  kHalfSpaceHalfSpace = 6,
  kSoftPlaneSoftMesh = 7,
  kRigidRigidPlane = 8
};

// Static description of one potential geometry pair (one entry of
// FindCollisionCandidates() after filtering, proximity_engine.cc:867-883, in
// the canonical order of MaybeMakeContactSurface, hydroelastic_calculator.cc:77-156).
struct alignas(64) PairRec {
  int32_t scene;
  int32_t geom_a;    // flat geometry index of the first-constructed body (S or M)
  int32_t geom_b;    // second body (R or N)
  int32_t kind;      // PairKind
  int32_t swapped;   // soft-rigid with id_R < id_S: ContactSurface swaps M and N (contact_surface.cc:39-41)
  int32_t root_a;    // pool index of the BVH root of the first body (-1 when unsupported)
  int32_t root_b;
  int32_t elem_base_a;  // pool index of element 0 of the first body (tet pool)
  int32_t elem_base_b;  // second body (tri pool for soft-rigid, tet pool for soft-soft)
  int32_t root_leaf;    // bit0: root_a is a leaf, bit1: root_b is a leaf
  int32_t small;        // mesh pair whose two trees have at most 32767 nodes: eligible for k_small_pairs (2: a large one,
                        // more than kSmallSplitNodes nodes in the two trees together: its descent is split into work items)
  int32_t pad;          // half-space pairs: root_* / elem_base_* describe the one mesh, the other side's are -1
  int64_t id_m;      // final ids, id_m < id_n
  int64_t id_n;
};
static_assert(sizeof(PairRec) == 64, "PairRec is 64 B");

// Rigid transform as 12 doubles: R column-major then p (math/rigid_transform.h:1004-1007).
struct Xf {
  double m[12];
};

// Bound of one geometry in its own frame, for the geometry-level broad phase of the pairs that only matter when
// they touch (the combinations MaybeMakeContactSurface rejects and the point-pair fallback pairs): the reference
// calls the calculator only for FindCollisionCandidates() results (proximity_engine.cc:867-883,905-934), i.e. for
// pairs whose FCL world boxes overlap. kBox: the mesh's root box, or the box the caller gave for a geometry with no
// hydroelastic representation; kEverywhere: no bound known; kHalfSpace: z <= 0.
enum GeomBoundKind : int32_t { kBoundBox = 0, kBoundEverywhere = 1, kBoundHalfSpace = 2 };
struct alignas(64) GeomBound {
  double c[3];
  double h[3];
  int32_t kind;
  int32_t pad[3];
};
static_assert(sizeof(GeomBound) == 64, "GeomBound is 64 B");

// Device counters of one query (zeroed when the query starts). A query runs
// over one or more chunks of scenes that reuse the same work lists; the list
// cursors are chunk-local (k_next_chunk folds them into totals / maxima and
// zeroes them), everything else accumulates over the whole query.
struct Counters {
  // -- chunk-local list cursors (they keep counting past a full list, so after
  //    an overflow they say how many entries the failing stage wanted)
  unsigned long long frontier[64];  // overlapping node pairs to expand, per traversal level
  unsigned long long leaf_tri;      // leaf pairs that passed the conservative fp32 test (tet-tri)
  unsigned long long leaf_tet;      // same, tet-tet
  unsigned long long polys;         // contact polygons in the staging soup
  unsigned long long slow;          // candidates deferred to the exceptional (16-vertex) path
  unsigned long long surv_tri;      // tet-tri candidates that survived the conservative culls
  unsigned long long surv_tet;      // same, tet-tet
  unsigned long long big_surfaces;  // surfaces of this chunk summed by the whole grid
  unsigned long long work_tri;      // next unclaimed 1024-entry chunk of the tet-tri leaf list (filter kernel)
  unsigned long long work_tet;      // same, tet-tet
  unsigned long long plane_emit;    // half-space pairs: elements that emit a polygon (entries of the pack pass's list)
  unsigned long long fallback;      // WithFallback query: pairs handed to the point-pair narrow phase (whole query)
  unsigned long long small_pairs[2]; // overlapping small pairs of the chunk, [0] tet-tri, [1] tet-tet (entries of
                                     // k_small_pairs' work lists)
  unsigned long long small_work[2];  // next unclaimed entry of each list
  unsigned long long small_items[2];      // work items of the large small pairs (k_small_pairs, pass 1 -> pass 2)
  unsigned long long small_items_work[2]; // next unclaimed item
  unsigned long long undecided;      // leaf pairs of small geometry pairs whose fp32 leaf predicate could not decide: entries of
                                     // the list k_leaf_undecided settles in fp64
  // -- whole query
  unsigned long long bv_tests;      // node-pair box tests (root + every child pair tested)
  unsigned long long bv_tests_tet;  // of which: between two tetrahedron trees (the rest: tetrahedron x triangle trees,
                                    // plus one per half-space pair)
  unsigned long long cand_tri;      // tet-tri candidate pairs: leaf pairs that pass the exact predicate
  unsigned long long cand_tet;      // tet-tet candidate pairs
  unsigned long long poly_verts;    // sum of polygon vertex counts
  unsigned long long polygons;      // contact polygons (polys_total counts staging slots: a polygon with
                                    // more than 8 vertices takes two)
  unsigned long long root_overlaps; // geometry pairs whose root boxes overlap
  unsigned long long surfaces;
  unsigned long long faces;
  unsigned long long vertices;
  unsigned long long full_clips_tri; // tet-tri candidates that survived the conservative culls
  unsigned long long full_clips_tet; // tet-tet candidates that survived the conservative culls
  unsigned long long leaf_exact;     // leaf pairs the fp32 predicate could not decide (exact fp64 test ran)
  unsigned long long cull_violations; // verify mode: polygons from candidates the conservative cull rejected
  unsigned long long plane_tets;      // tetrahedra sliced by a rigid half space (k_plane, active pairs only)
  unsigned long long plane_tris;      // triangles clipped by a compliant half space
  unsigned long long discrete_pairs;  // k_discrete: faces that give a discrete contact pair (zeroed by its launch)
  unsigned long long small_leaf_tri;  // leaf pairs k_small_pairs filtered on chip (not recording: they never reach a list)
  unsigned long long small_leaf_tet;
  // The surface-seeded soft-soft intersector (hydro_bfs.cuh): the wave being processed is queue[bfs_begin, bfs_end) of
  // the tet-tet survivor list (chunk-local, set by k_bfs_advance); pairs handed to the exact stage and waves run (whole
  // query).
  unsigned long long bfs_begin, bfs_end;
  unsigned long long bfs_pairs, bfs_waves;
  // -- totals and per-chunk maxima of the chunk-local cursors
  unsigned long long leaf_tri_total, leaf_tet_total, polys_total, slow_total;
  unsigned long long frontier_max, leaf_tri_max, leaf_tet_max, polys_max, slow_max, undecided_max, small_items_max;
  int32_t error;                    // first device-side dbh_status raised
  int32_t error_detail;
  double total_area;
};

constexpr int kPlaneBlock = 128;   // leaves per block box of the half-space pairs' tree cut (= threads of k_plane's blocks)
constexpr int kMaxPolyVerts = 8;    // field_intersection.cc:100-104; 7 for tet-tri (mesh_intersection.cc:261-262)

}  // namespace dbh
