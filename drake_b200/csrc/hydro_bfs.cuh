// The surface-seeded soft-soft intersector: the second VolumeIntersector::IntersectFields overload
// (field_intersection.cc:254-375) re-designed for the device. The reference builds it next to the volume-BVH
// overload, proves the two equivalent on its fixtures (field_intersection_test.cc:397-476) and leaves
// ComputeContactSurfaceFromCompliantVolumes on the volume one (use_surfaces = false, :511-517). Here it is an option
// of the context (dbh_context_set_soft_soft_intersector) because its work is proportional to the contact SURFACE
// instead of the overlap VOLUME: two 2 M-tetrahedron spheres have ~10^7 candidate pairs per scene through their
// volume hierarchies and ~10^5 through this path.
//
//   reference (serial)                                   here (whole chunk of scenes at once)
//   --------------------------------------------------   ------------------------------------------------------------
//   Bvh::Collide of the two SURFACE meshes' BVHs,         every compliant asset carries a second tree over its boundary
//   pairs of triangles -> pairs of their tetrahedra       triangles whose leaves name the triangle's tetrahedron
//   (tri_to_tet), std::unordered_set visited_pairs        (tri_to_tet folded into the leaf); the soft-soft PairRecs point
//                                                         at these roots, so the unchanged mid phase (k_small_pairs /
//                                                         k_traverse + k_filter) delivers the seeds into the survivor
//                                                         list, which doubles as the BFS queue; k_bfs_seed enters them
//                                                         in a device hash set of (pair, tet, tet) keys and drops
//                                                         duplicates;
//   std::queue; per pair CalcContactPolygon, then for     k_bfs_wave: one thread per queued pair runs the exact stage
//   every intersected face the neighbour across it        with IntersectTetrahedra's face codes carried along (4 bits per
//   (VolumeMeshTopology) joins the queue unless           vertex), emits the polygon into the same staging soup k_clip
//   visited                                               fills, and appends the unvisited face neighbours (hash insert
//                                                         by atomicCAS, warp-aggregated append) for the next wave.
//
// The visit ORDER differs (waves instead of one FIFO) and the seeds are found through per-triangle AABBs instead of
// 3-triangle OBB leaves, so the set of visited pairs can differ from the reference's. What holds: every polygon is one
// the volume overload produces (same pair, same values); the polygons are those of the pairs the walk reaches from the
// rim of the contact patch. Like the reference's overload it can miss polygons: a body wholly inside another has no
// seed (the reference's own caveat, field_intersection_test.cc:296-344), and the walk stops at a pair that produces
// nothing (filters, slivers), which on dense clutter of grid meshes costs a few faces on a fifth of the surfaces --
// a faithful CPU restatement of the reference's overload loses them too (tests/test_zz_gpu_surface_bfs.py).
#pragma once
#include "hydro_clip.cuh"

namespace dbh {

// Face codes of a polygon's vertices (IntersectTetrahedra's face_buffer, field_intersection.cc:96-108): 0-3 a face of
// tet0, 4-7 a face of tet1; 4 bits per slot, up to 16 slots.
typedef unsigned long long FaceCodes;
__host__ __device__ __forceinline__ int face_code_at(FaceCodes c, int slot) { return (int)((c >> (4 * slot)) & 0xFull); }
__host__ __device__ __forceinline__ FaceCodes face_code_put(FaceCodes c, int slot, int code) {
  if (slot >= 16) return c;
  return (c & ~(0xFull << (4 * slot))) | ((FaceCodes)(code & 0xF) << (4 * slot));
}

// kMarchingTetsFaceTable (mesh_plane_intersection.cc:99-117): the face of the tetrahedron on which the polygon edge
// FROM vertex k lies, one per nibble in the order of kMarchingTetsEdgeTable's edges, 0xF = none.
__host__ __device__ __forceinline__ unsigned int marching_tets_face_row(int code) {
  const unsigned short kFaceTable[16] = {0xFFFF, 0xF312, 0xF203, 0x0312, 0xF013, 0x3012, 0x2013, 0xF201,
                                         0xF102, 0x2310, 0x3210, 0xF310, 0x0213, 0xF302, 0xF213, 0xFFFF};
  return kFaceTable[code];
}

// clip_by_half_space_edges (the clip loop of IntersectTetrahedra, field_intersection.cc:146-182) with the face codes
// carried along (:158-180): a kept vertex keeps its code, the crossing of an edge that leaves the half space gets the
// clipping face's code, the crossing of an edge that enters keeps the code of the edge's first vertex.
template <class Poly>
__host__ __device__ __forceinline__ int clip_by_half_space_edges_faces(Poly& P, int buf, int n, D3 nh, double d,
                                                                       int clip_code, FaceCodes in, FaceCodes* out) {
  int m = 0;
  FaceCodes o = 0;
  const D3 first = P.get(buf, 0);
  const double d_first = dot(nh, first) - d;
  D3 cur = first;
  double di = d_first;
  for (int k = 0; k < n; ++k) {
    D3 nxt = first;
    double dj = d_first;
    if (k + 1 < n) {
      nxt = P.get(buf, k + 1);
      dj = dot(nh, nxt) - d;
    }
    if (di <= 0) {
      if (m < Poly::kCap) P.set(1 - buf, m, cur);
      o = face_code_put(o, m, face_code_at(in, k));
      ++m;
      if (dj > 0) {
        const double wa = dj / (dj - di);
        const double wb = 1.0 - wa;
        if (m < Poly::kCap) P.set(1 - buf, m, wa * cur + wb * nxt);
        o = face_code_put(o, m, clip_code);
        ++m;
      }
    } else if (dj <= 0) {
      const double wa = dj / (dj - di);
      const double wb = 1.0 - wa;
      if (m < Poly::kCap) P.set(1 - buf, m, wa * cur + wb * nxt);
      o = face_code_put(o, m, face_code_at(in, k));
      ++m;
    }
    cur = nxt;
    di = dj;
  }
  *out = o;
  return m;
}

// What CalcContactPolygon (field_intersection.cc:377-428) gives the BFS for one pair of tetrahedra.
struct TetPairPolygon {
  int n;             // vertices of the polygon, in P's buffer `buf` (0: no polygon)
  int buf;
  unsigned faces;    // bit c: code c is among IntersectTetrahedra's returned faces (0 when n == 0)
  int status;        // 0; 9: a face normal of tet1 is degenerate (Plane ctor throws, plane.cc:16-29);
                     // 1: an intermediate polygon outgrew Poly::kCap vertices (run again on larger storage)
  D3 grad0, nhat_M;  // field0's gradient and the polygon's normal, frame M
  double f0_Mo;
  FaceCodes codes;   // per vertex of the polygon before the last RemoveNearlyDuplicateVertices (the returned vector)
  int codes_n;
};

// The exact stage of a tet-tet pair with the reference's operation order -- the same statements as tet_stage_b
// (hydro_kernels.cuh), kept apart so that the hot kernel's code is untouched -- plus the face codes. The first
// RemoveNearlyDuplicateVertices (field_intersection.cc:113) shortens the polygon but not face_buffer, so a vertex
// that survives it reads the code at its NEW index: reproduced by leaving the codes alone.
template <class Poly>
__host__ __device__ __forceinline__ void tet_pair_polygon_with_faces(const Xf& X, const TetRec& tet0, const TetRec& tet1,
                                                                     const TetAux& aux0, const TetAux& aux1, Poly& P,
                                                                     TetPairPolygon* r) {
  r->n = 0;
  r->buf = 0;
  r->faces = 0;
  r->status = 0;
  r->codes = 0;
  r->codes_n = 0;
  // CalcEquilibriumPlane (field_intersection.cc:30-83).
  const D3 grad0 = load3(tet0.grad);
  const double f0_Mo = tet0.p_mo;
  r->grad0 = grad0;
  r->f0_Mo = f0_Mo;
  r->nhat_M = make_d3(0, 0, 0);
  const D3 grad1_N = load3(tet1.grad);
  const D3 grad1_M = rot(X, grad1_N);
  const D3 p_NMo = -rot_t(X, make_d3(X.m[9], X.m[10], X.m[11]));
  const double f1_Mo = dot(grad1_N, p_NMo) + tet1.p_mo;
  const D3 n_M = grad0 - grad1_M;
  const double magnitude = norm(n_M);
  if (!(magnitude > 0.0)) return;
  const D3 nhat_M = n_M / magnitude;
  r->nhat_M = nhat_M;
  const D3 p_MQ = ((f1_Mo - f0_Mo) / magnitude) * nhat_M;
  const double displacement = dot(nhat_M, p_MQ);
  // IsPlaneNormalAlongPressureGradient x2 (field_intersection.cc:200-216,393-400).
  if (!(dot(nhat_M, load3(aux0.ngrad)) > kCosAlphaDevice)) return;
  const D3 rev_N = rot_t(X, -nhat_M);
  if (!(dot(rev_N, load3(aux1.ngrad)) > kCosAlphaDevice)) return;
  // SliceTetrahedronWithPlane (mesh_plane_intersection.cc:121-180).
  double dist[4];
  int code = 0;
  for (int k = 0; k < 4; ++k) {
    dist[k] = dot(nhat_M, load3(tet0.v[k])) - displacement;
    if (dist[k] > 0) code |= 1 << k;
  }
  const unsigned short kEdgeTable[16] = {0xFFFF, 0xF230, 0xF410, 0x1234, 0xF521, 0x1530, 0x4520, 0xF453,
                                         0xF543, 0x0254, 0x0351, 0xF251, 0x4321, 0xF140, 0xF320, 0xFFFF};
  const unsigned entry = kEdgeTable[code];
  const unsigned face_row = marching_tets_face_row(code);
  int n = 0;
  FaceCodes codes = 0;
  for (int k = 0; k < 4; ++k) {
    const int edge = (entry >> (4 * k)) & 0xF;
    if (edge != 0xF && n == k) {
      // kTetEdges = {0,1},{1,2},{2,0},{0,3},{1,3},{2,3}.
      const int e0 = edge < 3 ? edge : edge - 3;
      const int e1 = edge < 3 ? (edge == 2 ? 0 : edge + 1) : 3;
      const double d0 = e0 == 0 ? dist[0] : (e0 == 1 ? dist[1] : dist[2]);
      const double d1 = e1 == 1 ? dist[1] : (e1 == 2 ? dist[2] : (e1 == 0 ? dist[0] : dist[3]));
      const D3 V0 = load3(tet0.v[e0]);
      const D3 V1 = load3(tet0.v[e1]);
      const double t = d0 / (d0 - d1);
      P.set(0, n, V0 + t * (V1 - V0));
      codes = face_code_put(codes, n, (int)((face_row >> (4 * k)) & 0xF));
      ++n;
    }
  }
  n = remove_nearly_duplicate(P, 0, n);
  if (n < 3) return;
  int buf = 0;
  // Clip by the four faces of tet1 (field_intersection.cc:120-190).
  for (int face = 0; face < 4; ++face) {
    const D3 p_MA = xform_point(X, load3(tet1.v[(face + 1) % 4]));
    const D3 outward = rot(X, -load3(aux1.inward[face]));
    const double mag = norm(outward);
    if (mag < 1e-10) {
      r->status = 9;
      return;
    }
    const D3 nh = outward / mag;
    const double d = dot(nh, p_MA);
    FaceCodes next = 0;
    n = clip_by_half_space_edges_faces(P, buf, n, nh, d, face + 4, codes, &next);
    codes = next;
    if (n > Poly::kCap) {
      r->status = 1;
      return;
    }
    buf = 1 - buf;
    if (n == 0) return;
  }
  const int returned = n;
  n = remove_nearly_duplicate(P, buf, n);
  if (n < 3) return;
  r->n = n;
  r->buf = buf;
  r->codes = codes;
  r->codes_n = returned;
  unsigned faces = 0;
  for (int k = 0; k < returned && k < 16; ++k) faces |= 1u << face_code_at(codes, k);
  r->faces = faces;
}

// Key of one (geometry pair, tetrahedron of M, tetrahedron of N) in the visited set: 22 + 21 + 21 bits
// (dbh_batch_create refuses the surface intersector beyond 2^22 - 2 pairs or 2^21 tetrahedra per mesh).
constexpr int kBfsPairBits = 22, kBfsTetBits = 21;
constexpr unsigned long long kBfsEmpty = ~0ull;
__host__ __device__ __forceinline__ unsigned long long bfs_key(int pair, int tet_a, int tet_b) {
  return ((unsigned long long)(unsigned)pair << (2 * kBfsTetBits)) | ((unsigned long long)(unsigned)tet_a << kBfsTetBits) |
         (unsigned long long)(unsigned)tet_b;
}

}  // namespace dbh
