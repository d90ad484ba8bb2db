// Device kernels of the B200 hydroelastic contact-surface pipeline (sm_100a).
//
//   k_pairs      pose composition X_MN = X_WM^-1 X_WN per geometry pair, level-0 frontier
//   k_traverse   one level of the dual-BVH traversal (replaces Bvh::Collide, bvh.h:220-262)
//   k_clip<0>    tet-triangle clipping (SurfaceVolumeIntersector::CalcContactPolygon,
//                mesh_intersection.cc:321-382)
//   k_clip<1>    tet-tet equilibrium-plane clipping (VolumeIntersector::CalcContactPolygon,
//                field_intersection.cc:377-428)
//   k_scan_*     per-pair face/vertex offsets (CSR) and surface compaction
//   k_pack       per-polygon area / normal / centroid / pressure, world transform, id-order
//                swap, final packed layout (PolyMeshBuilder/TriMeshBuilder + PolygonSurfaceMesh /
//                TriangleSurfaceMesh derived quantities + TransformVertices + SwapMAndN)
//   k_surface    per-surface total area and centroid
//   k_wrench     one spatial force per contact surface (hydroelastic_traction_calculator.cc)
//   k_plane_*    the half-space special cases (mesh_plane_intersection.cc:182-342,
//                mesh_half_space_intersection.cc:17-430): one thread per element of the one mesh
//
// All arithmetic is fp64 with the reference's operation order (hydro_math.cuh).
#pragma once
#include <cuda_runtime.h>

#include "hydro_clip.cuh"
#include "hydro_math.cuh"
#include "hydro_plane.cuh"
#include "hydro_wrench.cuh"
#include "hydro_types.h"

#ifdef DBH_TRACE
#include <cstdio>
// Debug build only (-DDBH_TRACE): prints the exact stage's intermediate values for one tet-tet
// candidate {pair, pool tet0, pool tet1} chosen by DBH_TRACE_PAIR="pair,tet0,tet1".
#define DBH_TR(...) do { if (c.x == a.trace.x && c.y == a.trace.y && c.z == a.trace.z) printf(__VA_ARGS__); } while (0)
#else
#define DBH_TR(...) do { } while (0)
#endif

namespace dbh {


// ---------------------------------------------------------------------------
// Warp-aggregated append for lists where every lane contributes 0 or 1 entries (ballot + popc; one atomic per warp).
__device__ __forceinline__ unsigned long long warp_append_flag(unsigned long long* counter, bool flag) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0 && m) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}

// Block-aggregated appends: ONE atomic per list per block iteration. The list cursors are single
// addresses hit by every block of the grid, and same-address atomics serialise in L2 (about one per
// nanosecond): per-warp appends ran the traversal and filter kernels at that limit. Every thread of
// the block must call these (two barriers inside).
template <int kWarps>
struct BlockAppendSmem {
  unsigned int tot[kWarps][2];
  unsigned long long base[3];
};
struct Slots3 {
  unsigned long long s0, s1, s2;
};
// Three lists at once; n0, n1, n2 <= 63 per thread, block totals < 65536.
template <int kWarps>
__device__ __forceinline__ Slots3 block_append3(unsigned long long* c0, unsigned long long* c1, unsigned long long* c2,
                                                int n0, int n1, int n2, BlockAppendSmem<kWarps>& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int a = (unsigned int)n0 | ((unsigned int)n1 << 16), b = (unsigned int)n2;
  unsigned int ia = a, ib = b;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int ya = __shfl_up_sync(0xffffffffu, ia, d);
    const unsigned int yb = __shfl_up_sync(0xffffffffu, ib, d);
    if (lane >= d) {
      ia += ya;
      ib += yb;
    }
  }
  if (lane == 31) {
    sh.tot[warp][0] = ia;
    sh.tot[warp][1] = ib;
  }
  __syncthreads();
  unsigned int pa = 0, pb = 0, ta = 0, tb = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const unsigned int x = sh.tot[w][0], y = sh.tot[w][1];
    if (w < warp) {
      pa += x;
      pb += y;
    }
    ta += x;
    tb += y;
  }
  if (threadIdx.x == 0 && (ta & 0xffffu)) sh.base[0] = atomicAdd(c0, (unsigned long long)(ta & 0xffffu));
  if (threadIdx.x == 32 && (ta >> 16)) sh.base[1] = atomicAdd(c1, (unsigned long long)(ta >> 16));
  if (threadIdx.x == 64 && tb) sh.base[2] = atomicAdd(c2, (unsigned long long)tb);
  __syncthreads();
  const unsigned int ea = pa + ia - a, eb = pb + ib - b;
  return Slots3{sh.base[0] + (ea & 0xffffu), sh.base[1] + (ea >> 16), sh.base[2] + eb};
}
// One list, 0 or 1 entries per thread.
template <int kWarps>
__device__ __forceinline__ unsigned long long block_append_flag(unsigned long long* counter, bool flag,
                                                                BlockAppendSmem<kWarps>& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) sh.tot[warp][0] = __popc(m);
  __syncthreads();
  unsigned int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const unsigned int x = sh.tot[w][0];
    if (w < warp) before += x;
    total += x;
  }
  if (threadIdx.x == 0 && total) sh.base[0] = atomicAdd(counter, (unsigned long long)total);
  __syncthreads();
  return sh.base[0] + before + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void raise_error(Counters* c, int status, int detail) {
  if (atomicCAS(&c->error, 0, status) == 0) c->error_detail = detail;
}

// ---------------------------------------------------------------------------
// Work-list entries.
//  frontier: {pair, a, b, kind}: an overlapping node pair still to be
//    expanded. a >= 0: pool index of the first of two adjacent children to
//    expand; a < 0: ~(pool index) of a leaf node that is kept as is. Same for b.
//  leaf list: {pair, pool element a, pool element b, 0}: leaf pairs that passed
//    the conservative fp32 box test; the narrow phase applies the exact predicate.
__device__ __forceinline__ int frontier_code(int node_index, int child) { return child < 0 ? ~node_index : child; }

// ---------------------------------------------------------------------------
// k_pairs: one thread per potential geometry pair: X_MN, its fp32 copy, the
// root box test, level-0 frontier.
struct PairsArgs {
  const PairRec* pairs;
  int pair_begin;        // this chunk's range of the static pair list
  int pair_end;
  Pools pools;
  const Xf* poses;       // X_WG per flat geometry index
  Xf* pair_X;            // out: X_MN per pair
  PairF32* pair_f32;     // out
  int4* frontier0;       // out: level-0 frontier
  int4* leaf_tri;        // out: single-element meshes go straight to the leaf lists
  int4* leaf_tet;
  Counters* counters;
  int strict;            // DBH_QUERY_STRICT
  int* pair_polys;       // zeroed here
  int* pair_verts;
  unsigned long long* pair_cursor;  // hi 32: polygons placed, lo 32: polygon vertices placed
  int* small_list;       // out: overlapping tet-tri pairs that k_small_pairs descends on chip (null: every pair
                         // takes the generic path)
  int* small_list_tet;   // same, tet-tet pairs
  const GeomBound* geom_bounds;      // per flat geometry: its bound in its own frame (broad phase of the no-surface pairs)
  int* fallback_list;                // out, WithFallback query: no-surface pairs whose world boxes overlap
  unsigned long long fallback_capacity;
};

__global__ void k_pairs(PairsArgs a) {
  // Launched with whole warps covering the chunk's pair range rounded up to 32.
  const int p = a.pair_begin + blockIdx.x * blockDim.x + threadIdx.x;
  bool push = false, tested = false;
  PairRec pr;
  pr.root_a = pr.root_b = pr.kind = pr.root_leaf = pr.small = 0;
  int child_a = 0, child_b = 0;
  if (p < a.pair_end) {
    a.pair_polys[p] = 0;
    a.pair_verts[p] = 0;
    a.pair_cursor[p] = 0ull;
    pr = a.pairs[p];
    // The pairs that produce no surface -- the combinations MaybeMakeContactSurface rejects (hydroelastic_calculator.cc:
    // 77-156) -- only matter when the geometry-level broad phase reports them (proximity_engine.cc:867-883): the strict
    // query then fails with the reference's result code (proximity_engine.cc:1361-1390), the WithFallback query hands
    // them to the point-pair narrow phase (:958-970; here: lists them for the host, dbh_batch_copy_fallback_pairs).
    const bool no_surface = pr.kind == kUnsupported || pr.kind == kRigidRigid || pr.kind >= kHalfSpaceHalfSpace;
    if (no_surface) {
      if (world_boxes_overlap(a.geom_bounds[pr.geom_a], a.poses[pr.geom_a], a.geom_bounds[pr.geom_b], a.poses[pr.geom_b])) {
        if (a.strict) {
          const int code = pr.kind == kUnsupported ? 7 /*DBH_ERR_UNSUPPORTED*/
                           : pr.kind == kHalfSpaceHalfSpace ? 11 /*DBH_ERR_HALF_SPACE_HALF_SPACE*/
                           : pr.kind == kSoftPlaneSoftMesh ? 12 /*DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH*/
                                                           : 6 /*DBH_ERR_RIGID_RIGID*/;
          raise_error(a.counters, code, p);
        } else {
          const unsigned long long slot = atomicAdd(&a.counters->fallback, 1ull);
          if (slot < a.fallback_capacity) a.fallback_list[slot] = p;
        }
      }
    } else if (pr.kind <= kSoftSoft) {
      const Xf X_WA = a.poses[pr.geom_a];
      const Xf X_WB = a.poses[pr.geom_b];
      const Xf X_AB = invert_and_compose(X_WA, X_WB);
      a.pair_X[p] = X_AB;
      PairF32 f;
#pragma unroll
      for (int k = 0; k < 9; ++k) f.R[k] = (float)X_AB.m[k];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) f.AR[i * 3 + j] = fabsf(f.R[j * 3 + i]) + 1e-6f;
        f.p[i] = (float)X_AB.m[9 + i];
      }
      f.S = fabsf(f.p[0]) + fabsf(f.p[1]) + fabsf(f.p[2]);
      f.pad[0] = f.pad[1] = 0.f;
      a.pair_f32[p] = f;
      // Root boxes: the exact 15-axis test on boxes padded by 1e-12 (an ancestor must never hide a passing leaf
      // pair).
      const BvhNode& na = a.pools.nodes[pr.root_a];
      const BvhNode& nb = a.pools.nodes[pr.root_b];
      child_a = na.left;
      child_b = nb.left;
      double ha[3] = {na.h[0], na.h[1], na.h[2]};
      double hb[3] = {nb.h[0], nb.h[1], nb.h[2]};
      const D3 t = xform_point(X_AB, load3(nb.c)) - load3(na.c);
      const double s =
          1e-12 * (fabs(t.x) + fabs(t.y) + fabs(t.z) + ha[0] + ha[1] + ha[2] + hb[0] + hb[1] + hb[2]) + 1e-300;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ha[k] += s;
        hb[k] += s;
      }
      tested = true;
      push = boxes_overlap(ha, hb, X_AB, t);
    }
    // (kSoftMeshRigidPlane / kSoftPlaneRigidMesh: k_plane_setup poses the half space and tests the mesh's root box.)
  }
  const unsigned tests = __ballot_sync(0xffffffffu, tested);
  const unsigned tests_tet = __ballot_sync(0xffffffffu, tested && pr.kind == kSoftSoft);
  const unsigned hits = __ballot_sync(0xffffffffu, push);
  if ((threadIdx.x & 31) == 0) {
    if (tests) atomicAdd(&a.counters->bv_tests, (unsigned long long)__popc(tests));
    if (tests_tet) atomicAdd(&a.counters->bv_tests_tet, (unsigned long long)__popc(tests_tet));
    if (hits) atomicAdd(&a.counters->root_overlaps, (unsigned long long)__popc(hits));
  }
  // Small pairs (both trees fit 15-bit node indices) are descended by one warp each in k_small_pairs.
  const bool small = a.small_list != nullptr && pr.small != 0;
  const unsigned long long s0 = warp_append_flag(&a.counters->small_pairs[0], push && small && pr.kind == kSoftRigid);
  const unsigned long long s1 = warp_append_flag(&a.counters->small_pairs[1], push && small && pr.kind == kSoftSoft);
  if (push && small) (pr.kind == kSoftRigid ? a.small_list : a.small_list_tet)[pr.kind == kSoftRigid ? s0 : s1] = p;
  push = push && !small;
  const bool both_leaves = child_a < 0 && child_b < 0;
  const unsigned long long slot = warp_append_flag(&a.counters->frontier[0], push && !both_leaves);
  if (push && !both_leaves)
    a.frontier0[slot] = make_int4(p, frontier_code(pr.root_a, child_a), frontier_code(pr.root_b, child_b), pr.kind);
  const unsigned long long lt = warp_append_flag(&a.counters->leaf_tri, push && both_leaves && pr.kind == kSoftRigid);
  const unsigned long long lq = warp_append_flag(&a.counters->leaf_tet, push && both_leaves && pr.kind == kSoftSoft);
  if (push && both_leaves) {
    // Lists are at least as long as a chunk's pair range (host guarantees).
    if (pr.kind == kSoftRigid) a.leaf_tri[lt] = make_int4(p, ~child_a, ~child_b, 0);
    else a.leaf_tet[lq] = make_int4(p, ~child_a, ~child_b, 0);
  }
}

// ---------------------------------------------------------------------------
// k_traverse: one level of the dual-tree descent (replaces Bvh::Collide,
// bvh.h:220-262). One thread per overlapping node pair: it fetches the two
// children of each internal side (one 64 B line per side) and tests the up to
// four child pairs with a conservative fp32 6-axis box test (the face axes of
// Gottschalk's SAT; outward-rounded boxes plus a 4e-6 relative pad cover every
// fp32 rounding, so a pair that the exact fp64 predicate accepts at the leaves
// is never dropped on the way down). Overlapping child pairs go to the next
// level's frontier, overlapping leaf pairs to the leaf lists.
struct TraverseArgs {
  const int4* in;
  int4* out;
  int level;
  unsigned long long out_capacity;
  const PairF32* pair_f32;
  Pools pools;
  int4* leaf_tri;
  int4* leaf_tet;
  unsigned long long leaf_capacity_tri;
  unsigned long long leaf_capacity_tet;
  Counters* counters;
};

struct Box32 {
  float cx, cy, cz, hx, hy, hz, l1;
};

// Conservative: false only if the boxes are certainly disjoint under the
// exact predicate. tb = R cb + p (b's centre in a's frame). (Hoisting the one-box parts of the right-hand
// sides out of the four child tests saves a sixth of the instructions on paper and was measured 8 % slower:
// the extra live values spill.)
__device__ __forceinline__ bool overlap6_f32(const Box32& A, const Box32& B, float tbx, float tby, float tbz, float sb,
                                             const PairF32& f) {
  const float tx = tbx - A.cx, ty = tby - A.cy, tz = tbz - A.cz;
  const float S = A.l1 + sb;
  const float delta = 4e-6f * S;
  const float ax = A.hx + delta, ay = A.hy + delta, az = A.hz + delta;
  const float bx = B.hx + delta, by = B.hy + delta, bz = B.hz + delta;
  // a's face axes.
  if (fabsf(tx) > ax + fmaf(bx, f.AR[0], fmaf(by, f.AR[1], bz * f.AR[2]))) return false;
  if (fabsf(ty) > ay + fmaf(bx, f.AR[3], fmaf(by, f.AR[4], bz * f.AR[5]))) return false;
  if (fabsf(tz) > az + fmaf(bx, f.AR[6], fmaf(by, f.AR[7], bz * f.AR[8]))) return false;
  // b's face axes: t . R(:, i).
  if (fabsf(fmaf(tx, f.R[0], fmaf(ty, f.R[1], tz * f.R[2]))) > bx + fmaf(ax, f.AR[0], fmaf(ay, f.AR[3], az * f.AR[6])))
    return false;
  if (fabsf(fmaf(tx, f.R[3], fmaf(ty, f.R[4], tz * f.R[5]))) > by + fmaf(ax, f.AR[1], fmaf(ay, f.AR[4], az * f.AR[7])))
    return false;
  if (fabsf(fmaf(tx, f.R[6], fmaf(ty, f.R[7], tz * f.R[8]))) > bz + fmaf(ax, f.AR[2], fmaf(ay, f.AR[5], az * f.AR[8])))
    return false;
  return true;
}

__device__ __forceinline__ Box32 load_box32(const Node32& n) {
  return Box32{n.c[0], n.c[1], n.c[2], n.h[0], n.h[1], n.h[2], n.l1};
}

template <int kMinBlocks, int kThreads>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_traverse(TraverseArgs a) {
  // The query already failed (the host re-runs it): one block-uniform decision, because the appends below are
  // block-wide collectives and another block may raise the flag while this one starts.
  __shared__ BlockAppendSmem<kThreads / 32> sh;
  __shared__ int failed_already;
  if (threadIdx.x == 0) failed_already = a.counters->error;
  __syncthreads();
  if (failed_already != 0) return;
  // After an overflow the counter exceeds what was stored (the query fails
  // with DBH_ERR_CAPACITY); never read past the stored entries.
  const unsigned long long count = min(a.counters->frontier[a.level], a.out_capacity);
  const unsigned long long stride = (unsigned long long)gridDim.x * kThreads;
  unsigned int tests = 0, tests_tet = 0;
  // Whole blocks stay in the loop: the appends are block-wide. (A two-deep software pipeline with L1
  // prefetches, which pays in the filter kernel, was measured here and does not: the nodes of the batched
  // configs are L1-resident and the extra registers spill.)
  for (unsigned long long base = (unsigned long long)blockIdx.x * kThreads; base < count; base += stride) {
    const unsigned long long i = base + threadIdx.x;
    int4 e = make_int4(0, 0, 0, 0);
    // Results of the up to four child tests as bit masks (bit q = jb * 2 + ja), so nothing is indexed
    // dynamically and everything stays in registers; the entries are rebuilt from the masks when stored.
    unsigned pass = 0, leafm = 0;
    int ia = 0, ib = 0, child_a0 = 0, child_a1 = 0, child_b0 = 0, child_b1 = 0;
    if (i < count) {
      e = a.in[i];
      const int na = e.y >= 0 ? 2 : 1, nb = e.z >= 0 ? 2 : 1;
      ia = e.y >= 0 ? e.y : ~e.y;
      ib = e.z >= 0 ? e.z : ~e.z;
      const PairF32 f = a.pair_f32[e.x];
      Node32 A[2], B[2];
      A[0] = a.pools.nodes32[ia];
      B[0] = a.pools.nodes32[ib];
      A[1] = na == 2 ? a.pools.nodes32[ia + 1] : A[0];
      B[1] = nb == 2 ? a.pools.nodes32[ib + 1] : B[0];
      child_a0 = A[0].child;
      child_a1 = A[1].child;
      child_b0 = B[0].child;
      child_b1 = B[1].child;
      tests = tests + na * nb;
      if (e.w == kSoftSoft) tests_tet = tests_tet + na * nb;
#pragma unroll
      for (int jb = 0; jb < 2; ++jb) {
        if (jb < nb) {
          const Box32 bb = load_box32(B[jb]);
          const float tbx = fmaf(f.R[0], bb.cx, fmaf(f.R[3], bb.cy, fmaf(f.R[6], bb.cz, f.p[0])));
          const float tby = fmaf(f.R[1], bb.cx, fmaf(f.R[4], bb.cy, fmaf(f.R[7], bb.cz, f.p[1])));
          const float tbz = fmaf(f.R[2], bb.cx, fmaf(f.R[5], bb.cy, fmaf(f.R[8], bb.cz, f.p[2])));
          const float sb = bb.l1 + f.S;
#pragma unroll
          for (int ja = 0; ja < 2; ++ja) {
            if (ja < na) {
              if (overlap6_f32(load_box32(A[ja]), bb, tbx, tby, tbz, sb, f)) {
                pass |= 1u << (jb * 2 + ja);
                if (A[ja].child < 0 && B[jb].child < 0) leafm |= 1u << (jb * 2 + ja);
              }
            }
          }
        }
      }
    }
    const int n_leaf = __popc(leafm), n_next = __popc(pass) - n_leaf;
    const bool tri = e.w == kSoftRigid;
    const Slots3 slots = block_append3<kThreads / 32>(&a.counters->frontier[a.level + 1], &a.counters->leaf_tri, &a.counters->leaf_tet,
                                          n_next, tri ? n_leaf : 0, tri ? 0 : n_leaf, sh);
    unsigned long long s_next = slots.s0, s_leaf = tri ? slots.s1 : slots.s2;
    bool ok_next = true, ok_leaf = true;
    if (n_next && s_next + n_next > a.out_capacity) {
      raise_error(a.counters, 5 /*DBH_ERR_CAPACITY*/, 1);
      ok_next = false;
    }
    if (n_leaf && s_leaf + n_leaf > (tri ? a.leaf_capacity_tri : a.leaf_capacity_tet)) {
      raise_error(a.counters, 5, 2);
      ok_leaf = false;
    }
    int4* const leaf_list = tri ? a.leaf_tri : a.leaf_tet;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if ((pass >> q) & 1u) {
        const int ja = q & 1, jb = q >> 1;
        const int ca = ja ? child_a1 : child_a0, cb = jb ? child_b1 : child_b0;
        if ((leafm >> q) & 1u) {
          if (ok_leaf) leaf_list[s_leaf] = make_int4(e.x, ~ca, ~cb, 0);  // pool element indices
          ++s_leaf;
        } else {
          if (ok_next) a.out[s_next] = make_int4(e.x, frontier_code(ia + ja, ca), frontier_code(ib + jb, cb), e.w);
          ++s_next;
        }
      }
    }
  }
  // One atomic per warp for the test count.
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    tests += __shfl_xor_sync(0xffffffffu, tests, d);
    tests_tet += __shfl_xor_sync(0xffffffffu, tests_tet, d);
  }
  if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&a.counters->bv_tests, (unsigned long long)tests);
  if ((threadIdx.x & 31) == 0 && tests_tet) atomicAdd(&a.counters->bv_tests_tet, (unsigned long long)tests_tet);
}

// ---------------------------------------------------------------------------
// Narrow phase, two kernels per element kind:
//   k_filter (every leaf pair, one per thread): the leaf predicate, then stage
//     A: conservative fp32 culls that can only say "certainly no polygon" --
//     the gradient filters with a safety band and separating face / hull
//     planes with a relative margin. Anything uncertain survives and is
//     compacted into a dense survivor list.
//   k_clip (survivors only, dense warps): stage B, the reference's exact
//     operation sequence in fp64.
// So the emitted polygons are exactly those of running stage B on every
// candidate, but divergent lanes no longer idle through another lane's clip and
// the filter runs without the clip's shared-memory footprint.
//
// Per-thread polygon scratch in shared memory: [buffer][slot][coord][thread],
// so a warp's access to one (slot, coord) is a single conflict-free 256 B row
// whatever slot each lane is on.
// kThreads: the block size; kSlots: vertices per buffer (8: IntersectTetrahedra's bound; a triangle clipped by four
// half spaces has at most 7).
template <int kThreads, int kSlots>
struct PolyScratchT {
  static constexpr int kCap = kSlots;
  double* base;  // &smem[threadIdx.x]
  __device__ __forceinline__ double& at(int buf, int slot, int c) const {
    return base[((buf * kSlots + slot) * 3 + c) * kThreads];
  }
  __device__ __forceinline__ D3 get(int buf, int slot) const {
    return D3{at(buf, slot, 0), at(buf, slot, 1), at(buf, slot, 2)};
  }
  __device__ __forceinline__ void set(int buf, int slot, D3 v) const {
    at(buf, slot, 0) = v.x;
    at(buf, slot, 1) = v.y;
    at(buf, slot, 2) = v.z;
  }
};
constexpr int clip_smem_bytes(int threads, int slots) { return 2 * slots * 3 * threads * 8; }

// The staging soup of contact polygons (frame M of the first body).
struct PolySoup {
  int4* hdr;            // {pair, pool elem a, pool elem b, n}
  double* v;            // [(slot*4 + c) * capacity + i], c = x, y, z, pressure
  double* nhat;         // [c * capacity + i]: polygon normal in M (triangle rep needs it)
  unsigned long long capacity;
};

struct ClipArgs {
  // In: leaf pairs {pair, pool tet, pool tri | pool tet, 0} from the descent. The
  // kernel sets the flag of every entry: 1 iff the pair passes the exact leaf
  // predicate, i.e. is a candidate element pair (dbh_batch_copy_candidates
  // reads the flagged entries).
  int4* cand;
  unsigned long long cand_capacity;
  // The survivors of the filter stage {pair, pool tet, pool tri | pool tet, 1 (| 2 culled, verify
  // mode)}: the traversal's frontier buffers, free once the descent is done.
  int4* surv;
  unsigned long long surv_capacity;
  const Xf* pair_X;
  const PairF32* pair_f32;
  Pools pools;
  PolySoup soup;
  int* pair_polys;
  int* pair_verts;
  Counters* counters;
  int4* slow;           // exceptional candidates {pair, elem, elem, 0 tri | 1 tet (| 2: culled, verify mode)}
  unsigned int slow_capacity;
  int4* undecided;      // leaf pairs the fp32 leaf predicate cannot decide {pair, pool elem a, pool elem b, index + 1 of the
                        // entry in the leaf list | 0}: settled in fp64 by k_leaf_undecided (hydro_small.cuh)
  unsigned long long undecided_capacity;
  // Self-check mode (dbh_batch_set_verify): the exact stage runs for EVERY candidate and
  // Counters::cull_violations_* count polygons emitted by candidates stage A had rejected.
  int verify;
  // Tet-triangle pairs without IsFaceNormalAlongPressureGradient: SurfaceVolumeIntersector::SampleVolumeFieldOnSurface's
  // filter_face_normal_along_field_gradient = false (mesh_intersection.h), which is how the reference's own
  // mesh_intersection_benchmark runs it. ComputeContactSurfaces always filters (the default).
  int no_normal_filter;
  int tune;             // experiments (DBH_TUNE "stagea="): bit0 no triangle-edge planes, bit1 no triangle plane, bit2 no
                        // tet-tet face planes, bit3 no tet-tet equilibrium-plane cull (each only widens the survivor set)
  int4 trace;           // debug builds only
};

// A candidate whose intermediate polygon does not fit the fast path's 8 slots.
__device__ __forceinline__ void defer_to_slow_path(const ClipArgs& a, int4 c, int kind) {
  const unsigned long long slot = atomicAdd(&a.counters->slow, 1ull);
  if (slot < a.slow_capacity) {
    a.slow[slot] = make_int4(c.x, c.y, c.z, kind | (c.w & 2));
  } else {
    raise_error(a.counters, 5, 8);
  }
}

// Called by all 32 lanes of a warp (n == 0: nothing to emit for this lane).
template <class Poly>
__device__ __forceinline__ void emit_polygon(const ClipArgs& a, Poly& P, int buf, int n, int4 c, D3 grad, double p_mo,
                                             D3 nhat) {
  const unsigned long long slot = warp_append_flag(&a.counters->polys, n > 0);
  // Per-pair polygon / vertex counts (the scan's input). The polygons of a warp mostly belong to one
  // geometry pair: then one lane adds the warp's totals instead of 32 lanes hitting the same address.
  {
    const unsigned emitting = __ballot_sync(0xffffffffu, n > 0);
    if (emitting) {
      const int lane = threadIdx.x & 31;
      const int lead = __ffs(emitting) - 1;
      const int pair0 = __shfl_sync(0xffffffffu, c.x, lead);
      const bool uniform = __all_sync(0xffffffffu, n <= 0 || c.x == pair0);
      if (uniform) {
        int total = n > 0 ? n : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
        if (lane == lead) {
          atomicAdd(&a.pair_polys[pair0], __popc(emitting));
          atomicAdd(&a.pair_verts[pair0], total);
        }
      } else if (n > 0) {
        atomicAdd(&a.pair_polys[c.x], 1);
        atomicAdd(&a.pair_verts[c.x], n);
      }
    }
  }
  if (n <= 0) return;
  if (slot >= a.soup.capacity) {
    raise_error(a.counters, 5, 3);
    return;
  }
  if (c.w & 2) atomicAdd(&a.counters->cull_violations, 1ull);  // verify mode only
  a.soup.hdr[slot] = make_int4(c.x, c.y, c.z, n);
  for (int k = 0; k < n; ++k) {
    const D3 p = P.get(buf, k);
    a.soup.v[(k * 4 + 0) * a.soup.capacity + slot] = p.x;
    a.soup.v[(k * 4 + 1) * a.soup.capacity + slot] = p.y;
    a.soup.v[(k * 4 + 2) * a.soup.capacity + slot] = p.z;
    // MeshFieldLinear::EvaluateCartesian (mesh_field_linear.h:349).
    a.soup.v[(k * 4 + 3) * a.soup.capacity + slot] = dot(grad, p) + p_mo;
  }
  a.soup.nhat[0 * a.soup.capacity + slot] = nhat.x;
  a.soup.nhat[1 * a.soup.capacity + slot] = nhat.y;
  a.soup.nhat[2 * a.soup.capacity + slot] = nhat.z;
}

// The exceptional path's emit: per-lane atomics. A polygon with more than 8 vertices (only a clip
// plane coplanar with the polygon can produce one; the reference's vectors have no bound) takes two
// consecutive staging slots: vertices 8.. live in the second, whose header has n = 0.
template <class Poly>
__device__ __forceinline__ void emit_polygon_slow(const ClipArgs& a, Poly& P, int buf, int n, int4 c, D3 grad,
                                                  double p_mo, D3 nhat) {
  if (n <= 0) return;
  const int slots = n > kMaxPolyVerts ? 2 : 1;
  const unsigned long long slot = atomicAdd(&a.counters->polys, (unsigned long long)slots);
  if (slot + slots > a.soup.capacity) {
    raise_error(a.counters, 5, 3);
    return;
  }
  if (c.w & 2) atomicAdd(&a.counters->cull_violations, 1ull);  // verify mode only
  a.soup.hdr[slot] = make_int4(c.x, c.y, c.z, n);
  if (slots == 2) a.soup.hdr[slot + 1] = make_int4(c.x, c.y, c.z, 0);
  for (int k = 0; k < n; ++k) {
    const D3 p = P.get(buf, k);
    const unsigned long long s = slot + (k >> 3);
    const int kk = k & 7;
    a.soup.v[(kk * 4 + 0) * a.soup.capacity + s] = p.x;
    a.soup.v[(kk * 4 + 1) * a.soup.capacity + s] = p.y;
    a.soup.v[(kk * 4 + 2) * a.soup.capacity + s] = p.z;
    a.soup.v[(kk * 4 + 3) * a.soup.capacity + s] = dot(grad, p) + p_mo;
  }
  a.soup.nhat[0 * a.soup.capacity + slot] = nhat.x;
  a.soup.nhat[1 * a.soup.capacity + slot] = nhat.y;
  a.soup.nhat[2 * a.soup.capacity + slot] = nhat.z;
  atomicAdd(&a.pair_polys[c.x], 1);
  atomicAdd(&a.pair_verts[c.x], n);
}

// ---- tet-triangle ---------------------------------------------------------
// Why the culls below cannot change the result. Every vertex the reference's
// clipper creates is a convex combination (weights in [0,1]) of the three
// transformed triangle vertices, so it lies in hull(tri) up to rounding e ~
// 1e-15 L; every vertex that survives all four half spaces has height <= e in
// each of them, i.e. lies in the tetrahedron bounded by the face planes offset
// by e, which is the tetrahedron scaled about its incentre by (1 + e/inradius)
// and reaches at most e * amp beyond it (amp = longest edge / inradius, stored
// in TetAux). Hence
//  (R) three triangle vertices outside one face plane by a margin: empty;
//  (N) tetrahedron separated from hull(tri) -- by the triangle's plane or by
//      one of its three edge planes -- by more than 1e-11 (1 + amp) L: empty.
// Slivers get a huge amp and are simply never culled by (N).

// The culls run in fp32 on fp32 copies of the records (TetCullA/B, TriCull,
// PairF32): a cull never produces a value, it only has to be certain, so its
// margins are simply widened to 2e-5 relative (fp32 rounding through these
// short expressions stays below ~2e-6 of the same scale).
struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 f3(const float* p) { return F3{p[0], p[1], p[2]}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dotf(F3 a, F3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ F3 crossf(F3 a, F3 b) {
  return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float abs_sumf(F3 a) { return fabsf(a.x) + fabsf(a.y) + fabsf(a.z); }
__device__ __forceinline__ F3 rotf(const PairF32& f, F3 v) {
  return F3{fmaf(f.R[0], v.x, fmaf(f.R[3], v.y, f.R[6] * v.z)), fmaf(f.R[1], v.x, fmaf(f.R[4], v.y, f.R[7] * v.z)),
            fmaf(f.R[2], v.x, fmaf(f.R[5], v.y, f.R[8] * v.z))};
}
__device__ __forceinline__ F3 rot_tf(const PairF32& f, F3 v) {
  return F3{fmaf(f.R[0], v.x, fmaf(f.R[1], v.y, f.R[2] * v.z)), fmaf(f.R[3], v.x, fmaf(f.R[4], v.y, f.R[5] * v.z)),
            fmaf(f.R[6], v.x, fmaf(f.R[7], v.y, f.R[8] * v.z))};
}
__device__ __forceinline__ F3 xformf(const PairF32& f, F3 v) {
  const F3 r = rotf(f, v);
  return F3{r.x + f.p[0], r.y + f.p[1], r.z + f.p[2]};
}
constexpr float kCullMarginF = 2e-5f;
constexpr float kCosAlphaF = -0.38268343f;

// ---- the leaf predicate ------------------------------------------------------
// A leaf pair is a candidate iff the two elements' own boxes pass the scalar
// 15-axis BoxesOverlap in fp64 (aabb.cc:16-33, boxes_overlap.cc:197-255). The
// fp32 pass below evaluates the same 15 inequalities on boxes rebuilt from the
// fp32 vertex copies and answers only when every axis is decided by more than
// 1e-5 of the scale of the operands (fp32 rounding through these expressions
// stays below ~1e-6 of it); anything closer is settled by the exact predicate.
struct BoxF {
  float c[3], h[3];
  float l1;  // |c|_1 + |h|_1
};
template <int N>
__device__ __forceinline__ BoxF box_of(const float (*v)[3]) {
  BoxF b;
  b.l1 = 0.f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float mn = v[0][d], mx = v[0][d];
#pragma unroll
    for (int k = 1; k < N; ++k) {
      mn = fminf(mn, v[k][d]);
      mx = fmaxf(mx, v[k][d]);
    }
    b.c[d] = 0.5f * (mn + mx);
    b.h[d] = mx - b.c[d];
    b.l1 += fabsf(b.c[d]) + b.h[d];
  }
  return b;
}
// +1: certainly overlapping, -1: certainly separated, 0: too close to call in fp32.
// Error budget per axis (relative 1e-5, fp32 rounding stays below ~1e-6): the translation t carries an
// absolute error proportional to S = |ca|_1 + |ha|_1 + |cb|_1 + |hb|_1 + |p|_1, so face axes get
// 1e-5 S. An edge axis multiplies t by two entries of R, so its error scales with |R(i1,j)| + |R(i2,j)|
// -- which is what keeps the parallel-axes case decidable: with R(i1,j) = R(i2,j) = 0 both sides
// vanish except for the 1e-6 term of the right-hand side, and the margin shrinks with them.
__device__ __forceinline__ int boxes_overlap_f32(const BoxF& A, const BoxF& B, const PairF32& f) {
  const float S = A.l1 + B.l1 + f.S;
  const float delta = 1e-5f * S;
  float t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    t[i] = fmaf(f.R[i], B.c[0], fmaf(f.R[3 + i], B.c[1], fmaf(f.R[6 + i], B.c[2], f.p[i]))) - A.c[i];
  bool separated = false, unsure = false;
  auto axis = [&](float lhs, float rhs, float d) {
    const float g = lhs - rhs;
    separated |= g > d;
    unsure |= g > -d;
  };
#pragma unroll
  for (int i = 0; i < 3; ++i)
    axis(fabsf(t[i]), A.h[i] + fmaf(B.h[0], f.AR[i * 3 + 0], fmaf(B.h[1], f.AR[i * 3 + 1], B.h[2] * f.AR[i * 3 + 2])),
         delta);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    axis(fabsf(fmaf(t[0], f.R[i * 3 + 0], fmaf(t[1], f.R[i * 3 + 1], t[2] * f.R[i * 3 + 2]))),
         B.h[i] + fmaf(A.h[0], f.AR[0 * 3 + i], fmaf(A.h[1], f.AR[1 * 3 + i], A.h[2] * f.AR[2 * 3 + i])), delta);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float rhs = fmaf(A.h[i1], f.AR[i2 * 3 + j], A.h[i2] * f.AR[i1 * 3 + j]) +
                        fmaf(B.h[j1], f.AR[i * 3 + j2], B.h[j2] * f.AR[i * 3 + j1]);
      // AR = |R| + 1e-6: a slight over-estimate of |R(i1,j)| + |R(i2,j)|.
      axis(fabsf(t[i2] * f.R[j * 3 + i1] - t[i1] * f.R[j * 3 + i2]), rhs,
           1e-5f * fmaf(S, f.AR[i1 * 3 + j] + f.AR[i2 * 3 + j], rhs));
    }
  }
  return separated ? -1 : (unsure ? 0 : 1);
}
// The exact predicate (rare: only when fp32 cannot decide).
template <bool kTet>
__device__ __noinline__ bool leaf_predicate_exact(const ClipArgs& a, int4 c) {
  const ElemBox& ba = a.pools.ebox_tet[c.y];
  const ElemBox& bb = kTet ? a.pools.ebox_tet[c.z] : a.pools.ebox_tri[c.z];
  const Xf X = a.pair_X[c.x];
  const D3 t = xform_point(X, load3(bb.c)) - load3(ba.c);
  return boxes_overlap(ba.h, bb.h, X, t);
}

// Stage A. false = certainly no polygon.
__device__ __forceinline__ bool tri_stage_a(const ClipArgs& a, int4 c, const PairF32& X, const TetCullA& ta,
                                            const TriCull& tri) {
  const TetCullB& tb = a.pools.cull_b[c.y];
  // Gradient filter (certain rejections only).
  const F3 nhat = rotf(X, f3(tri.n));
  const bool unit = fabsf(dotf(nhat, nhat) - 1.0f) < 1e-4f;
  if (unit && !a.no_normal_filter && dotf(f3(ta.grad), nhat) * tb.inv_grad_len < kCosAlphaF - 1e-3f) return false;
  const F3 q0 = xformf(X, f3(tri.v[0])), q1 = xformf(X, f3(tri.v[1])), q2 = xformf(X, f3(tri.v[2]));
  F3 v[4];
  float L = fmaxf(fmaxf(abs_sumf(q0), abs_sumf(q1)), abs_sumf(q2));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = f3(ta.v[k]);
    L = fmaxf(L, abs_sumf(v[k]));
  }
  bool keep = true;
  // (R) tet face planes against the triangle's vertices.
  const float mr = kCullMarginF * L;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const F3 nh = f3(tb.nh[f]);
    const float d = dotf(nh, v[(f + 1) % 4]);
    if (fminf(fminf(dotf(nh, q0) - d, dotf(nh, q1) - d), dotf(nh, q2) - d) > mr) keep = false;
  }
  // (N) the triangle's plane and edge planes against the tet's vertices.
  if (unit) {
    const float m = tb.margin * L;  // inf for slivers: never culls
    const float w = dotf(nhat, q0);
    int above = 0, below = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float h = dotf(nhat, v[k]) - w;
      above += h > m;
      below += h < -m;
    }
    if (!(a.tune & 2) && (above == 4 || below == 4)) keep = false;
    if (a.tune & 1) return keep;
    // Edge e from qa to qb; u = nhat x (qb - qa) points into the triangle.
    const F3 qa[3] = {q0, q1, q2};
    const F3 qb[3] = {q1, q2, q0};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const F3 edge = qb[e] - qa[e];
      const F3 u = crossf(nhat, edge);
      // u.(v - qa) < thr as u.v < thr + u.qa: the cancellation error (~1e-7 |u| L) is far inside the margin.
      const float thr = dotf(u, qa[e]) - m * abs_sumf(edge);
      int outside = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) outside += dotf(u, v[k]) < thr;
      if (outside == 4) keep = false;
    }
  }
  return keep;
}

// Stage B: SurfaceVolumeIntersector::CalcContactPolygon (mesh_intersection.cc:321-382).
template <class Poly>
__device__ __forceinline__ void tri_stage_b(const ClipArgs& a, Poly& P, int4 c, bool active) {
  constexpr bool kSlow = Poly::kCap > kMaxPolyVerts;
  int n = 0;
  int buf = 0;
  D3 grad = make_d3(0, 0, 0), nhat_M = make_d3(0, 0, 0);
  double p_mo = 0;
  if (active) {
    const Xf X = a.pair_X[c.x];
    const TetRec& tet = a.pools.tets[c.y];
    const TetAux& aux = a.pools.aux[c.y];
    const TriRec& tri = a.pools.tris[c.z];
    grad = load3(tet.grad);
    p_mo = tet.p_mo;
    // IsFaceNormalAlongPressureGradient (mesh_intersection.cc:266-288).
    nhat_M = rot(X, load3(tri.n));
    if (a.no_normal_filter || face_normal_along_gradient(load3(aux.ngrad), nhat_M)) {
      // ClipTriangleByTetrahedron (mesh_intersection.cc:176-264) with the face
      // half spaces precomputed at registration (TetAux).
      n = 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) P.set(0, k, xform_point(X, load3(tri.v[k])));
      // The next face's half space is requested while the current one clips (the loop is bound by the latency of
      // these loads: 4 % of the kernel, profiles/README.md).
      D3 nh_next = load3(aux.plane_nh[0]);
      double d_next = aux.plane_d[0];
#pragma unroll 1
      for (int f = 0; f < 4; ++f) {
        const D3 nh = nh_next;
        const double d = d_next;
        if (f < 3 && n > 0) {
          nh_next = load3(aux.plane_nh[f + 1]);
          d_next = aux.plane_d[f + 1];
        }
        if (n > 0) {
          if (d != d) {  // Plane ctor throws (plane.cc:16-29)
            raise_error(a.counters, 9 /*DBH_ERR_PLANE_NORMAL*/, c.x);
            n = 0;
          } else {
            n = clip_by_half_space(P, buf, n, nh, d);
            if (n > Poly::kCap) {
              if (kSlow) raise_error(a.counters, 10 /*DBH_ERR_POLYGON_SIZE*/, c.x); else defer_to_slow_path(a, c, 0);
              n = 0;
            }
            buf = 1 - buf;
          }
        }
      }
      n = remove_nearly_duplicate(P, buf, n);
      if (n < 3) n = 0;
      if (n > 7) {  // DRAKE_DEMAND(polygon_M->size() <= 7), mesh_intersection.cc:261-262
        raise_error(a.counters, 10, c.x);
        n = 0;
      }
    }
  }
  if constexpr (kSlow) emit_polygon_slow(a, P, buf, n, c, grad, p_mo, nhat_M);
  else emit_polygon(a, P, buf, n, c, grad, p_mo, nhat_M);
}

// ---- tet-tet --------------------------------------------------------------
// Same argument: the sliced polygon's vertices lie on edges of tet0 (in
// hull(tet0) up to e) and within e of the equilibrium plane; vertices that
// survive tet1's four half spaces lie in tet1 inflated by e * amp1.
//  (R) tet0 entirely on one side of the equilibrium plane (marching-tets code
//      0 or 15), or entirely outside one face of tet1, by a margin: empty;
//  (N) tet1 entirely on one side of the equilibrium plane, or entirely outside
//      one face of tet0, by more than 1e-11 (1 + amp1) L: empty.
// Stage A. false = certainly no polygon.
__device__ __forceinline__ bool tet_stage_a(const ClipArgs& a, int4 c, const PairF32& X, const TetCullA& a0,
                                            const TetCullA& a1) {
  const TetCullB& b0 = a.pools.cull_b[c.y];
  const TetCullB& b1 = a.pools.cull_b[c.z];
  const F3 g0 = f3(a0.grad), g1 = f3(a1.grad);
  const F3 p_NMo = rot_tf(X, F3{-X.p[0], -X.p[1], -X.p[2]});
  const float f1_Mo = dotf(g1, p_NMo) + a1.p_mo;
  const F3 n = g0 - rotf(X, g1);
  const float s = dotf(n, n);
  const float gsum = abs_sumf(g0) + abs_sumf(g1);
  // n is a difference of gradients: when it nearly cancels fp32 cannot say
  // anything about the plane; leave the pair to the exact stage.
  if (!(s > 1e-6f * gsum * gsum)) return true;
  // Both gradient filters on the un-normalised normal: cos <= cos(5pi/8) < 0 is
  // certain when c < 0 and c^2 > cos^2 |n|^2 (1 + 2 %).
  const float kCos2 = kCosAlphaF * kCosAlphaF * 1.02f;
  const float c_0 = dotf(n, g0) * b0.inv_grad_len;
  if (c_0 < 0 && c_0 * c_0 > kCos2 * s) return false;
  const float c_1 = -dotf(rot_tf(X, n), g1) * b1.inv_grad_len;
  if (c_1 < 0 && c_1 * c_1 > kCos2 * s) return false;
  // Equilibrium plane: |n| * height(p) = n.p - (f1(Mo) - f0(Mo)).
  const float g = f1_Mo - a0.p_mo;
  F3 v0[4], v1[4];
  float L = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v0[k] = f3(a0.v[k]);
    v1[k] = xformf(X, f3(a1.v[k]));
    L = fmaxf(L, fmaxf(abs_sumf(v0[k]), abs_sumf(v1[k])));
  }
  // Scale of every term that enters a height, times the margin.
  const float scale = gsum * L + (fabsf(a0.p_mo) + fabsf(a1.p_mo) + abs_sumf(g1) * X.S);
  const float tol0 = kCullMarginF * scale;   // (R) tet0 against the plane
  const float tol1 = b1.margin * scale;      // (N) tet1 against the plane (margin >= 2e-5)
  int above0 = 0, below0 = 0, above1 = 0, below1 = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float h0 = dotf(n, v0[k]) - g;
    const float h1 = dotf(n, v1[k]) - g;
    above0 += h0 > tol0;
    below0 += h0 < -tol0;
    above1 += h1 > tol1;
    below1 += h1 < -tol1;
  }
  if (!(a.tune & 8) && (above0 == 4 || below0 == 4 || above1 == 4 || below1 == 4)) return false;
  bool keep = true;
  if (a.tune & 4) return keep;
  const float mr = kCullMarginF * L, mn = b1.margin * L;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    // (R) a face of tet1 (outward normal rotated into M) with all of tet0 outside.
    // (n.(v - A) > m evaluated as n.v > m + n.A: the cancellation error, ~1e-7 L, is far inside the margins.)
    const F3 o = rotf(X, f3(b1.nh[f]));
    const float da = dotf(o, v1[(f + 1) % 4]) + mr;
    int out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) out += dotf(o, v0[k]) > da;
    if (out == 4) keep = false;
    // (N) a face of tet0 with all of tet1 outside.
    const F3 nh = f3(b0.nh[f]);
    const float db = dotf(nh, v0[(f + 1) % 4]) + mn;
    out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) out += dotf(nh, v1[k]) > db;
    if (out == 4) keep = false;
  }
  return keep;
}

// Stage B: VolumeIntersector::CalcContactPolygon (field_intersection.cc:377-428).
template <class Poly>
__device__ __forceinline__ void tet_stage_b(const ClipArgs& a, Poly& P, int4 c, bool active) {
  constexpr bool kSlow = Poly::kCap > kMaxPolyVerts;
  int n = 0;
  int buf = 0;
  D3 grad0 = make_d3(0, 0, 0), nhat_M = make_d3(0, 0, 0);
  double f0_Mo = 0;
  if (active) {
    const Xf X = a.pair_X[c.x];
    const TetRec& tet0 = a.pools.tets[c.y];
    const TetRec& tet1 = a.pools.tets[c.z];
    const TetAux& aux0 = a.pools.aux[c.y];
    const TetAux& aux1 = a.pools.aux[c.z];
    // CalcEquilibriumPlane (field_intersection.cc:30-83).
    grad0 = load3(tet0.grad);
    f0_Mo = tet0.p_mo;
    const D3 grad1_N = load3(tet1.grad);
    const D3 grad1_M = rot(X, grad1_N);
    const D3 p_NMo = -rot_t(X, make_d3(X.m[9], X.m[10], X.m[11]));
    const double f1_Mo = dot(grad1_N, p_NMo) + tet1.p_mo;
    const D3 n_M = grad0 - grad1_M;
    const double magnitude = norm(n_M);
    bool ok = magnitude > 0.0;
    double displacement = 0;
    if (ok) {
      nhat_M = n_M / magnitude;
      const D3 p_MQ = ((f1_Mo - f0_Mo) / magnitude) * nhat_M;
      displacement = dot(nhat_M, p_MQ);
      // IsPlaneNormalAlongPressureGradient x2 (field_intersection.cc:200-216,393-400).
      ok = dot(nhat_M, load3(aux0.ngrad)) > kCosAlphaDevice;
      if (ok) {
        const D3 rev_N = rot_t(X, -nhat_M);
        ok = dot(rev_N, load3(aux1.ngrad)) > kCosAlphaDevice;
      }
    }
    DBH_TR("[trace] filters ok %d magnitude %.17g f1_Mo %.17g f0_Mo %.17g\n", (int)ok, magnitude, f1_Mo, f0_Mo);
    if (ok) {
      // SliceTetrahedronWithPlane (mesh_plane_intersection.cc:121-180).
      double dist[4];
      int code = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dist[k] = dot(nhat_M, load3(tet0.v[k])) - displacement;
        if (dist[k] > 0) code |= 1 << k;
      }
      // kMarchingTetsEdgeTable packed 4 bits per entry (0xF = none), edges kTetEdges.
      const unsigned short kEdgeTable[16] = {0xFFFF, 0xF230, 0xF410, 0x1234, 0xF521, 0x1530, 0x4520, 0xF453,
                                             0xF543, 0x0254, 0x0351, 0xF251, 0x4321, 0xF140, 0xF320, 0xFFFF};
      const unsigned entry = kEdgeTable[code];
      DBH_TR("[trace] mag %.17g nhat %.17g %.17g %.17g disp %.17g dist %.17g %.17g %.17g %.17g code %d\n", magnitude,
             nhat_M.x, nhat_M.y, nhat_M.z, displacement, dist[0], dist[1], dist[2], dist[3], code);
      n = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int edge = (entry >> (4 * k)) & 0xF;
        if (edge != 0xF && n == k) {
          // kTetEdges = {0,1},{1,2},{2,0},{0,3},{1,3},{2,3}.
          const int e0 = edge < 3 ? edge : edge - 3;
          const int e1 = edge < 3 ? (edge == 2 ? 0 : edge + 1) : 3;
          // dist[] is indexed dynamically; select to keep it in registers.
          const double d0 = e0 == 0 ? dist[0] : (e0 == 1 ? dist[1] : dist[2]);
          const double d1 = e1 == 1 ? dist[1] : (e1 == 2 ? dist[2] : (e1 == 0 ? dist[0] : dist[3]));
          const D3 V0 = load3(tet0.v[e0]);
          const D3 V1 = load3(tet0.v[e1]);
          const double t = d0 / (d0 - d1);
          P.set(0, n, V0 + t * (V1 - V0));
          ++n;
        }
      }
      DBH_TR("[trace] sliced n %d\n", n);
#ifdef DBH_TRACE
      for (int q = 0; q < 4; ++q) {
        const D3 pv = P.get(0, q);
        DBH_TR("[trace]   sliced slot%d %.17g %.17g %.17g\n", q, pv.x, pv.y, pv.z);
      }
#endif
      n = remove_nearly_duplicate(P, 0, n);
#ifdef DBH_TRACE
      for (int q = 0; q < 4; ++q) {
        const D3 pv = P.get(0, q);
        DBH_TR("[trace]   deduped slot%d %.17g %.17g %.17g\n", q, pv.x, pv.y, pv.z);
      }
#endif
      DBH_TR("[trace] deduped n %d\n", n);
      if (n < 3) n = 0;
      // Clip by the four faces of tet1 (field_intersection.cc:120-190).
#pragma unroll 1
      for (int face = 0; face < 4; ++face) {
        if (n > 0) {
          const D3 p_MA = xform_point(X, load3(tet1.v[(face + 1) % 4]));
          const D3 outward = rot(X, -load3(aux1.inward[face]));
          const double mag = norm(outward);
          if (mag < 1e-10) {
            raise_error(a.counters, 9, c.x);
            n = 0;
          } else {
            const D3 nh = outward / mag;
            const double d = dot(nh, p_MA);
#ifdef DBH_TRACE
            for (int q = 0; q < n; ++q) {
              const D3 pv = P.get(buf, q);
              DBH_TR("[trace]   in v%d %.17g %.17g %.17g h %.17g\n", q, pv.x, pv.y, pv.z, dot(nh, pv) - d);
            }
#endif
            n = clip_by_half_space_edges(P, buf, n, nh, d);
            DBH_TR("[trace] face %d nh %.17g %.17g %.17g d %.17g -> n %d\n", face, nh.x, nh.y, nh.z, d, n);
            if (n > Poly::kCap) {
              // More than 16 intermediate vertices: far outside the reference's defined behaviour
              // (its distances array holds 8, field_intersection.cc:98).
              if (kSlow) raise_error(a.counters, 10 /*DBH_ERR_POLYGON_SIZE*/, c.x); else defer_to_slow_path(a, c, 1);
              n = 0;
            }
            buf = 1 - buf;
          }
        }
      }
      if (n > 0) {
        n = remove_nearly_duplicate(P, buf, n);
        if (n < 3) n = 0;
        if (!kSlow && n > kMaxPolyVerts) {  // cannot happen: the fast path's scratch holds 8
          raise_error(a.counters, 10, c.x);
          n = 0;
        }
      }
    }
  }
  if constexpr (kSlow) emit_polygon_slow(a, P, buf, n, c, grad0, f0_Mo, nhat_M);
  else emit_polygon(a, P, buf, n, c, grad0, f0_Mo, nhat_M);
}

// ---- k_filter: the leaf predicate + stage A over every leaf pair ------------
// One thread per leaf pair, light enough on registers to keep the SM full. The
// kernel is bound by the latency of its gathers (leaf entry -> per-element
// records), so the loop is software-pipelined two deep: the entry of iteration
// k + 2 is loaded and the records of iteration k + 1 are prefetched into L1 while
// iteration k computes. Survivors are staged per warp in shared memory (ballot +
// popc, no atomics) and flushed 32 at a time: one global atomic and one coalesced
// 512 B store per 32 survivors, no block-wide barrier. Warps claim contiguous
// chunks of the leaf list dynamically (the cost per pair varies with the survivor
// density). Every entry's candidate flag is written back to the leaf list.
template <bool kTet>
__device__ __forceinline__ void prefetch_records(const ClipArgs& a, int4 c) {
  prefetch_l1(a.pair_f32 + c.x);
  prefetch_l1(a.pools.cull_a + c.y);
  prefetch_l1(a.pools.cull_b + c.y);
  if (kTet) {
    prefetch_l1(a.pools.cull_a + c.z);
    prefetch_l1(a.pools.cull_b + c.z);
  } else {
    prefetch_l1(a.pools.cull_tri + c.z);
  }
}
template <bool kTet, int kFilterThreads, int kMinBlocks>
__global__ void __launch_bounds__(kFilterThreads, kMinBlocks) k_filter(ClipArgs a) {
  __shared__ int4 stage[kFilterThreads / 32][64];
  if (a.counters->error != 0) return;
  const unsigned long long count = min(kTet ? a.counters->leaf_tet : a.counters->leaf_tri, a.cand_capacity);
  const unsigned long long rounded = (count + 31ull) & ~31ull;  // whole warps reach the ballots
  // Every warp claims contiguous chunks of the leaf list (one atomic per chunk: the cost per leaf pair varies
  // with the survivor density, so static ranges leave a tail) and walks them 32 entries at a time, so the
  // survivors it stages (and later the polygons they emit) stay in leaf-list order: long runs of the same
  // geometry pair, which is what keeps the pack kernel's writes contiguous.
  // Chunk: 1024 entries, smaller when the list is short (at least ~4 chunks per warp of the grid), never under 128.
  const unsigned long long num_warps = (unsigned long long)gridDim.x * (kFilterThreads / 32);
  const unsigned long long kChunk = min(1024ull, max(128ull, (count / (4 * num_warps)) & ~31ull));
  const unsigned long long stride = 32;
  const int lane = threadIdx.x & 31;
  int4* const my_stage = stage[threadIdx.x >> 5];
  int fill = 0;  // staged survivors of this warp (warp-uniform)
  unsigned int n_cand = 0, n_exact = 0;  // per warp (every lane holds the warp's count)
  unsigned long long* cursor = kTet ? &a.counters->surv_tet : &a.counters->surv_tri;
  unsigned long long* work = kTet ? &a.counters->work_tet : &a.counters->work_tri;
  auto flush = [&](int n) {  // the first n staged entries (n <= 32) go to the survivor list
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < n) {
      if (base + lane < a.surv_capacity) a.surv[base + lane] = my_stage[lane];
      else raise_error(a.counters, 5 /*DBH_ERR_CAPACITY*/, 1);
    }
  };
  const int4 kNone = make_int4(0, 0, 0, 0);
  for (;;) {
    unsigned long long warp_begin = 0;
    if (lane == 0) warp_begin = atomicAdd(work, 1ull) * kChunk;
    warp_begin = __shfl_sync(0xffffffffu, warp_begin, 0);
    if (warp_begin >= rounded) break;
    const unsigned long long warp_end = min(warp_begin + kChunk, rounded);
    unsigned long long i = warp_begin + lane;
    const unsigned long long last = min(warp_end, count);  // entries at or past `last` do not exist / are not ours
    int4 c_cur = i < last ? a.cand[i] : kNone;
    int4 c_next = i + stride < last ? a.cand[i + stride] : kNone;
    if (i < last) prefetch_records<kTet>(a, c_cur);
    for (; i < warp_end; i += stride) {
      const int4 c_next2 = i + 2 * stride < last ? a.cand[i + 2 * stride] : kNone;
      if (i + stride < last) prefetch_records<kTet>(a, c_next);
      bool keep = false, is_cand = false, undecided = false;
      int4 c = c_cur;
      if (i < count) {
        const PairF32 X = a.pair_f32[c.x];
        const TetCullA ta = a.pools.cull_a[c.y];
        int decided;
        if (kTet) {
          const TetCullA t1 = a.pools.cull_a[c.z];
          decided = boxes_overlap_f32(box_of<4>(ta.v), box_of<4>(t1.v), X);
          is_cand = decided > 0;
          if (is_cand) keep = tet_stage_a(a, c, X, ta, t1);
        } else {
          const TriCull tri = a.pools.cull_tri[c.z];
          decided = boxes_overlap_f32(box_of<4>(ta.v), box_of<3>(tri.v), X);
          is_cand = decided > 0;
          if (is_cand) keep = tri_stage_a(a, c, X, ta, tri);
        }
        undecided = decided == 0;
        c.w = is_cand ? 1 : 0;
        a.cand[i].w = c.w;
        if (is_cand && !keep && a.verify) {
          keep = true;
          c.w |= 2;
        }
      }
      n_cand += __popc(__ballot_sync(0xffffffffu, is_cand));
      // Undecided in fp32 (0.1 - 3 % of the leaf pairs): left to k_leaf_undecided, whose fp64 predicate would cost this
      // kernel a third of its resident warps in registers and stack.
      const unsigned um = __ballot_sync(0xffffffffu, undecided);
      if (um) {
        n_exact += __popc(um);
        unsigned long long ubase = 0;
        if (lane == 0) ubase = atomicAdd(&a.counters->undecided, (unsigned long long)__popc(um));
        ubase = __shfl_sync(0xffffffffu, ubase, 0);
        if (undecided) {
          const unsigned long long slot = ubase + __popc(um & ((1u << lane) - 1u));
          if (slot < a.undecided_capacity) a.undecided[slot] = make_int4(c.x, c.y, c.z, (int)i + 1);   // (a leaf list holds far fewer than 2^31 entries: 16 B each)
          else raise_error(a.counters, 5 /*DBH_ERR_CAPACITY*/, 10);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) my_stage[fill + __popc(m & ((1u << lane) - 1u))] = c;
      fill += __popc(m);
      __syncwarp();
      if (fill >= 32) {
        flush(32);
        __syncwarp();
        int4 moved = kNone;
        if (lane < fill - 32) moved = my_stage[32 + lane];
        __syncwarp();
        if (lane < fill - 32) my_stage[lane] = moved;
        fill -= 32;
        __syncwarp();
      }
      c_cur = c_next;
      c_next = c_next2;
    }
  }
  if (fill > 0) flush(fill);
  if (lane == 0 && n_cand) atomicAdd(kTet ? &a.counters->cand_tet : &a.counters->cand_tri, (unsigned long long)n_cand);
  if (lane == 0 && n_exact) atomicAdd(&a.counters->leaf_exact, (unsigned long long)n_exact);
}

// ---- k_clip: the exact stage over the survivors, one per thread, dense warps --
template <bool kTet, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_clip(ClipArgs a) {
  extern __shared__ double smem[];
  if (a.counters->error != 0) return;
  PolyScratchT<kThreads, kTet ? 8 : 7> P{smem + threadIdx.x};
  const unsigned long long count = min(kTet ? a.counters->surv_tet : a.counters->surv_tri, a.surv_capacity);
  // (Claiming blocks of survivors dynamically, which pays in the filter kernel, was measured 3 % slower here.)
  for (unsigned long long base = (unsigned long long)blockIdx.x * kThreads; base < count;
       base += (unsigned long long)gridDim.x * kThreads) {
    const unsigned long long i = base + threadIdx.x;
    const bool active = i < count;
    const int4 w = active ? a.surv[i] : make_int4(0, 0, 0, 0);
    if (kTet) tet_stage_b(a, P, w, active); else tri_stage_b(a, P, w, active);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && count)
    atomicAdd(kTet ? &a.counters->full_clips_tet : &a.counters->full_clips_tri, count);
}

// The exceptional path: the same stage B with 16-slot local-memory scratch.
__global__ void __launch_bounds__(64) k_clip_slow(ClipArgs a) {
  if (a.counters->error != 0) return;
  const unsigned long long count = min(a.counters->slow, (unsigned long long)a.slow_capacity);
  const unsigned long long rounded = (count + 31ull) & ~31ull;
  LocalPoly P;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < rounded;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const bool active = i < count;
    const int4 c = active ? a.slow[i] : make_int4(0, 0, 0, 0);
    tri_stage_b(a, P, c, active && (c.w & 1) == 0);
    tet_stage_b(a, P, c, active && (c.w & 1) == 1);
  }
}

// ---------------------------------------------------------------------------
// Exclusive scans over the geometry-pair list (three kernels: per-block
// partials, single-block scan of the partials, apply). Produces for every
// pair the polygon offset, the polygon-vertex offset and the surface index.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

struct ScanArgs {
  const int* pair_polys;
  const int* pair_verts;
  int pair_begin;        // this chunk's pair range; offsets continue from the previous chunk's
  int num_pairs;         // end of the range (exclusive)
  int first_chunk;       // the running offsets start at 0
  long long* poly_off;   // [num_pairs + 1]
  long long* vert_off;   // [num_pairs + 1]
  int* surf_index;       // [num_pairs + 1]
  long long* block_sums; // [3 * num_blocks]
  int num_blocks;
  int* surf_pair;        // [num_surfaces]
  Counters* counters;
};

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total, long long* sh /*[kScanThreads/32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) sh[w] = incl;
  __syncthreads();
  long long warp_base = 0, tot = 0;
  for (int k = 0; k < kScanThreads / 32; ++k) {
    if (k < w) warp_base += sh[k];
    tot += sh[k];
  }
  __syncthreads();
  *total = tot;
  return warp_base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_partials(ScanArgs a) {
  __shared__ long long sh[kScanThreads / 32];
  const int base = a.pair_begin + blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  long long s0 = 0, s1 = 0, s2 = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const int p = base + k;
    if (p < a.num_pairs) {
      const int np = a.pair_polys[p];
      s0 += np;
      s1 += a.pair_verts[p];
      s2 += np > 0;
    }
  }
  long long t0, t1, t2;
  block_exclusive_scan(s0, &t0, sh);
  block_exclusive_scan(s1, &t1, sh);
  block_exclusive_scan(s2, &t2, sh);
  if (threadIdx.x == 0) {
    a.block_sums[0 * a.num_blocks + blockIdx.x] = t0;
    a.block_sums[1 * a.num_blocks + blockIdx.x] = t1;
    a.block_sums[2 * a.num_blocks + blockIdx.x] = t2;
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_blocks(ScanArgs a) {
  // One block scans the per-block partials of the three sequences in turn.
  __shared__ long long sh[kScanThreads / 32];
  for (int seq = 0; seq < 3; ++seq) {
    // The previous chunk left its totals in the sentinel slot, which is this chunk's first.
    long long carry = 0;
    if (!a.first_chunk)
      carry = seq == 0 ? a.poly_off[a.pair_begin] : (seq == 1 ? a.vert_off[a.pair_begin] : a.surf_index[a.pair_begin]);
    for (int start = 0; start < a.num_blocks; start += kScanThreads) {
      const int b = start + threadIdx.x;
      const long long v = b < a.num_blocks ? a.block_sums[seq * a.num_blocks + b] : 0;
      long long tot;
      const long long ex = block_exclusive_scan(v, &tot, sh);
      if (b < a.num_blocks) a.block_sums[seq * a.num_blocks + b] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) {
      if (seq == 0) a.poly_off[a.num_pairs] = carry;
      if (seq == 1) a.vert_off[a.num_pairs] = carry;
      if (seq == 2) {
        a.surf_index[a.num_pairs] = (int)carry;
        a.counters->surfaces = (unsigned long long)carry;
      }
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(ScanArgs a) {
  __shared__ long long sh[kScanThreads / 32];
  const int base = a.pair_begin + blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int np[kScanItems], nv[kScanItems];
  long long s0 = 0, s1 = 0, s2 = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const int p = base + k;
    np[k] = p < a.num_pairs ? a.pair_polys[p] : 0;
    nv[k] = p < a.num_pairs ? a.pair_verts[p] : 0;
    s0 += np[k];
    s1 += nv[k];
    s2 += np[k] > 0;
  }
  long long t;
  long long e0 = block_exclusive_scan(s0, &t, sh) + a.block_sums[0 * a.num_blocks + blockIdx.x];
  long long e1 = block_exclusive_scan(s1, &t, sh) + a.block_sums[1 * a.num_blocks + blockIdx.x];
  long long e2 = block_exclusive_scan(s2, &t, sh) + a.block_sums[2 * a.num_blocks + blockIdx.x];
  for (int k = 0; k < kScanItems; ++k) {
    const int p = base + k;
    if (p < a.num_pairs) {
      a.poly_off[p] = e0;
      a.vert_off[p] = e1;
      a.surf_index[p] = (int)e2;
      if (np[k] > 0) a.surf_pair[e2] = p;
      e0 += np[k];
      e1 += nv[k];
      e2 += np[k] > 0;
    }
  }
}

// ---------------------------------------------------------------------------
// Final packed results, in the caller's layout (include/drake_b200_hydro.h
// dbh_result_buffers): 3-vectors are xyz-interleaved, faces and vertices of a
// surface are contiguous, surfaces ordered like the static pair list.
struct ResultArrays {
  double* vertices;      // [V*3]
  double* pressures;     // [V]
  int* face_data;        // polygon: |n|v..|; triangle: [F*3]; indices local to the surface
  double* areas;         // [F]
  double* normals;       // [F*3]
  double* centroids;     // [F*3]
  double* e_centroid;    // [F]
  double* grad_eM;       // [F*3]
  double* grad_eN;       // [F*3]
  int* elem_a;           // [F]
  int* elem_b;           // [F]
  int* face_pair;        // [F]
  int* vertex_origin;    // [V], half-space surfaces only (null when the batch has none): WorldPoly::origin per vertex,
                         // -1 for the centroid vertex of a triangle fan
  long long face_capacity;
  long long vert_capacity;
  long long face_data_capacity;
};

// k_pack: one thread per staged polygon (up to 16 vertices: the exceptional path's capacity).
struct PackArgs {
  PolySoup soup;
  const PairRec* pairs;
  Pools pools;
  const Xf* poses;
  const long long* poly_off;
  const long long* vert_off;
  unsigned long long* pair_cursor;
  int representation;  // DBH_TRIANGLE / DBH_POLYGON
  ResultArrays out;
  Counters* counters;
};

__device__ __forceinline__ void store3(double* base, long long i, D3 v) {
  base[3 * i] = v.x;
  base[3 * i + 1] = v.y;
  base[3 * i + 2] = v.z;
}

__device__ __forceinline__ void store_face(const PackArgs& a, long long f, int pair, int ea, int eb, double area,
                                           D3 n_W, D3 c_W, double e_c, D3 gA_W, D3 gB_W, bool has_b, bool swapped) {
  if (swapped) n_W = -n_W;  // ReverseFaceWinding negates normals (polygon_surface_mesh.cc:50-52)
  a.out.areas[f] = area;
  store3(a.out.normals, f, n_W);
  store3(a.out.centroids, f, c_W);
  a.out.e_centroid[f] = e_c;
  const D3 zero = make_d3(0, 0, 0);
  // swapped: grad_eM <-> grad_eN (contact_surface.cc:58).
  store3(a.out.grad_eM, f, swapped ? (has_b ? gB_W : zero) : gA_W);
  store3(a.out.grad_eN, f, swapped ? gA_W : (has_b ? gB_W : zero));
  a.out.elem_a[f] = ea;
  a.out.elem_b[f] = eb;
  a.out.face_pair[f] = pair;
}

// kRep: the representation (DBH_POLYGON = 1 / DBH_TRIANGLE = 0) as a template parameter: the two branches share
// little, and a kernel holding both is allocated the registers of the larger (polygon alone: 16 B of spills
// under the 128-register bound instead of 124 B).
template <int kMinBlocks, int kRep>
__global__ void __launch_bounds__(128, kMinBlocks) k_pack(PackArgs a) {
  if (a.counters->error != 0) return;
  const unsigned long long count = min(a.counters->polys, a.soup.capacity);
  const unsigned long long rounded = (count + 31ull) & ~31ull;  // whole warps reach the collectives
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride) {
    int4 h = make_int4(-1, 0, 0, 0);
    // The first three vertices of the staged polygon (every polygon has at least three) are requested together with
    // the header, before the dependent chain header -> pair record -> poses -> offsets starts: the kernel is bound by
    // the latency of these DRAM round trips, not by their number.
    D3 pv0 = make_d3(0, 0, 0), pv1 = pv0, pv2 = pv0;
    double pe0 = 0, pe1 = 0, pe2 = 0;
    if (i < count) {
      h = a.soup.hdr[i];
      const double* sv = a.soup.v + i;
      const unsigned long long cap = a.soup.capacity;
      pv0 = make_d3(sv[0 * cap], sv[1 * cap], sv[2 * cap]);
      pe0 = sv[3 * cap];
      pv1 = make_d3(sv[4 * cap], sv[5 * cap], sv[6 * cap]);
      pe1 = sv[7 * cap];
      pv2 = make_d3(sv[8 * cap], sv[9 * cap], sv[10 * cap]);
      pe2 = sv[11 * cap];
    }
    const int pair = h.x, n = h.w;
    // n == 0: the second staging slot of a polygon with more than 8 vertices.
    const bool valid = i < count && n != 0;
    // One atomic claims both the polygon slot and the vertex range, so within a pair polygon order
    // and vertex order agree. When the whole warp works on one geometry pair (large meshes: tens of
    // thousands of polygons per pair) a single atomic serves the warp; same-address atomics serialise.
    unsigned long long cur = 0;
    {
      const unsigned validmask = __ballot_sync(0xffffffffu, valid);
      const int lead = validmask ? __ffs(validmask) - 1 : 0;
      const int pair0 = __shfl_sync(0xffffffffu, pair, lead);
      const bool uniform = __all_sync(0xffffffffu, !valid || pair == pair0);
      const unsigned long long v = valid ? ((1ull << 32) | (unsigned long long)n) : 0ull;
      if (uniform) {
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long base = 0;
        if (lane == lead && total) base = atomicAdd(&a.pair_cursor[pair], total);
        base = __shfl_sync(0xffffffffu, base, lead);
        cur = base + incl - v;
      } else if (valid) {
        cur = atomicAdd(&a.pair_cursor[pair], v);
      }
    }
    if (!valid) continue;
    const PairRec pr = a.pairs[pair];
    const bool soft_soft = pr.kind == kSoftSoft;
    const bool swapped = pr.swapped != 0;
    const Xf X_WM = a.poses[pr.geom_a];
    // The polygon is streamed from the staging soup (coalesced SoA reads; a second pass hits L1 / L2):
    // no per-thread vertex arrays, so no local memory. Vertices 8.. of a large polygon are in the next slot.
    auto vert = [&](int k, double* e_k) {
      if (k < 3) {
        *e_k = k == 0 ? pe0 : (k == 1 ? pe1 : pe2);
        return k == 0 ? pv0 : (k == 1 ? pv1 : pv2);
      }
      const unsigned long long s = i + (k >> 3);
      const int kk = k & 7;
      *e_k = a.soup.v[(kk * 4 + 3) * a.soup.capacity + s];
      return make_d3(a.soup.v[(kk * 4 + 0) * a.soup.capacity + s], a.soup.v[(kk * 4 + 1) * a.soup.capacity + s],
                     a.soup.v[(kk * 4 + 2) * a.soup.capacity + s]);
    };
    const D3 gradA_M = load3(a.pools.tets[h.y].grad);
    const D3 gA_W = rot(X_WM, gradA_M);
    D3 gB_W = make_d3(0, 0, 0);
    if (soft_soft) {
      const Xf X_WN = a.poses[pr.geom_b];
      gB_W = rot(X_WN, load3(a.pools.tets[h.z].grad));
    }
    const int pj = (int)(cur >> 32);
    const int vj = (int)(cur & 0xffffffffull);
    double e0;
    const D3 v0 = vert(0, &e0);
    if constexpr (kRep == 1 /*DBH_POLYGON*/) {
      const long long f = a.poly_off[pair] + pj;
      const long long vb = a.vert_off[pair] + vj;
      if (f >= a.out.face_capacity || vb + n > a.out.vert_capacity || f + vb + n + 1 > a.out.face_data_capacity) {
        raise_error(a.counters, 5, 6);
        continue;
      }
      // PolygonSurfaceMesh::CalcAreaNormalAndCentroid (polygon_surface_mesh.cc:117-218): fan about vertex 0,
      // accumulated while the vertices stream by; vertices, pressures and face_data are written on the way
      // (TransformVertices, polygon_surface_mesh.cc:25-37; face_data |n|v0..vn-1| with surface-local indices,
      // reversed when M and N are swapped, ReverseFaceWinding :39-49).
      int* fd = a.out.face_data + (f + vb);
      fd[0] = n;
      const D3 r_MA = v0;
      D3 p_scaled = make_d3(0, 0, 0), normal_M = make_d3(0, 0, 0), vsum = v0;
      double double_area = 0;
      store3(a.out.vertices, vb, xform_point(X_WM, v0));
      a.out.pressures[vb] = e0;
      fd[1] = vj + (swapped ? (n - 1) : 0);
      D3 prev = v0;  // vertex k - 1
      // The loads of vertex k + 1 are issued before vertex k is consumed (the loop is bound by load latency).
      double e_nx = 0;
      D3 v_nx = n > 1 ? vert(1, &e_nx) : v0;
      for (int k = 1; k < n; ++k) {
        const double ek = e_nx;
        const D3 vk = v_nx;
        if (k + 1 < n) v_nx = vert(k + 1, &e_nx);
        store3(a.out.vertices, vb + k, xform_point(X_WM, vk));
        a.out.pressures[vb + k] = ek;
        fd[1 + k] = vj + (swapped ? (n - 1 - k) : k);
        vsum = vsum + vk;
        if (k >= 2) {  // triangle (0, k-1, k)
          const D3 cr = cross(prev - r_MA, vk - r_MA);
          normal_M = normal_M + cr;
          const double mag = norm(cr);
          double_area += mag;
          p_scaled = p_scaled + mag * ((r_MA + prev) + vk);
        }
        prev = vk;
      }
      const double area = double_area / 2;
      const D3 nM = normalized(normal_M);
      const D3 p_as = p_scaled / 6;
      const D3 cM = area != 0.0 ? p_as / area : vsum / (double)n;
      // field Transform (mesh_field_linear.h:377-383,502-509).
      const D3 n_W = normalized(rot(X_WM, nM));
      const D3 c_W = xform_point(X_WM, cM);
      const double at_Mo = e0 - dot(gradA_M, v0);
      const double at_Wo = at_Mo - dot(gA_W, make_d3(X_WM.m[9], X_WM.m[10], X_WM.m[11]));
      const double e_c = dot(gA_W, c_W) + at_Wo;
      store_face(a, f, pair, h.y - pr.elem_base_a, h.z - pr.elem_base_b, area, n_W, c_W, e_c, gA_W, gB_W, soft_soft, swapped);
    } else {
      // Triangle representation: n fan triangles about the polygon centroid
      // (contact_surface_utility.cc:137-164,218-316). Faces are counted by
      // polygon vertices, vertices by n + 1 per polygon.
      const long long fb = a.vert_off[pair] + vj;
      const int vloc = vj + pj;  // local to the surface
      const long long vb = a.vert_off[pair] + a.poly_off[pair] + vloc;
      if (fb + n > a.out.face_capacity || vb + n + 1 > a.out.vert_capacity ||
          3 * (fb + n) > a.out.face_data_capacity) {
        raise_error(a.counters, 5, 6);
        continue;
      }
      const D3 nhat = make_d3(a.soup.nhat[0 * a.soup.capacity + i], a.soup.nhat[1 * a.soup.capacity + i],
                              a.soup.nhat[2 * a.soup.capacity + i]);
      // Pass 1: vertices + pressures out, CalcPolygonCentroid.
      D3 accum = make_d3(0, 0, 0), vsum = v0;
      double total_weight = 0;
      store3(a.out.vertices, vb, xform_point(X_WM, v0));
      a.out.pressures[vb] = e0;
      D3 prev = v0;
      double e_prev = e0;
      for (int k = 1; k < n; ++k) {
        double ek;
        const D3 vk = vert(k, &ek);
        store3(a.out.vertices, vb + k, xform_point(X_WM, vk));
        a.out.pressures[vb + k] = ek;
        vsum = vsum + vk;
        if (k >= 2) {
          const double weight = dot(cross(prev - v0, vk - v0), nhat);
          accum = accum + weight * (((v0 + prev) + vk) / 3.0);
          total_weight += weight;
        }
        prev = vk;
        e_prev = ek;
      }
      D3 p_c;
      if (n == 3) {
        p_c = vsum / 3.0;  // ((v0 + v1) + v2) / 3
      } else {
        if (total_weight == 0) {
          accum = vsum;  // ((0 + v0) + v1) + ... : adding to zero first is exact
          total_weight = n;
        }
        p_c = accum / total_weight;
      }
      const double e_cv = dot(gradA_M, p_c - v0) + e0;
      const D3 cW = xform_point(X_WM, p_c);
      store3(a.out.vertices, vb + n, cW);
      a.out.pressures[vb + n] = e_cv;
      const double third = 1.0 / 3.0;
      // Pass 2: triangle (v_{k-1}, v_k, centroid), k = 0 pairs the last vertex (still in prev) with vertex 0.
      int k1 = n - 1;
      for (int k = 0; k < n; ++k) {
        double ek = e0;
        const D3 vk = k == 0 ? v0 : vert(k, &ek);
        // TriangleSurfaceMesh derived quantities in M (triangle_surface_mesh.h:452-487), then rotated.
        const D3 cr = cross(vk - prev, p_c - prev);
        const double nm = norm(cr);
        const double area = 0.5 * nm;
        const D3 nM = (nm != 0.0) ? cr / nm : cr;
        const D3 n_W = rot(X_WM, nM);
        const D3 c_W = ((xform_point(X_WM, prev) + xform_point(X_WM, vk)) + cW) / 3.0;
        double e_c = third * e_prev;
        e_c += third * ek;
        e_c += third * e_cv;
        int i0 = vloc + k1, i1 = vloc + k;
        if (swapped) {  // SurfaceTriangle::ReverseWinding swaps vertices 0 and 1
          const int tmp = i0;
          i0 = i1;
          i1 = tmp;
        }
        int* fd = a.out.face_data + 3 * (fb + k);
        fd[0] = i0;
        fd[1] = i1;
        fd[2] = vloc + n;
        store_face(a, fb + k, pair, h.y - pr.elem_base_a, h.z - pr.elem_base_b, area, n_W, c_W, e_c, gA_W, gB_W, soft_soft, swapped);
        k1 = k;
        prev = vk;
        e_prev = ek;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_surface: one warp per non-null surface: total area, area-weighted centroid
// (polygon_surface_mesh.cc:208-217, triangle_surface_mesh.h:476-486), offsets,
// ids, and the per-scene statistics record.
struct SurfaceArgs {
  const int* surf_index;     // [num_pairs + 1]: surfaces before each pair
  int pair_begin;            // this chunk's pair range
  int pair_end;
  const int* surf_pair;
  const PairRec* pairs;
  const Xf* poses;           // world poses, one per geometry (the all-zero-area centroid of a triangle mesh)
  const long long* poly_off;
  const long long* vert_off;
  int representation;
  const double* areas;
  const double* centroids;
  long long* surf_face_off;  // [num_surfaces + 1]
  long long* surf_vert_off;
  double* surf_total_area;   // [num_surfaces]
  double* surf_centroid;     // [num_surfaces*3]
  int* surf_scene;
  long long* surf_id_m;
  long long* surf_id_n;
  unsigned char* surf_flags;
  long long* scene_stats;    // optional [num_scenes * 6]
  int* big_list;             // surfaces with more than kBigSurfaceFaces faces (this chunk)
  Counters* counters;
};

// Surfaces with more faces than this are summed by the whole grid (k_surface_big_*), not by one warp.
constexpr long long kBigSurfaceFaces = 4096;
constexpr long long kBigSegment = 2048;

__device__ __forceinline__ void surface_face_range(const SurfaceArgs& a, int p, long long* f0, long long* f1, long long* v0) {
  if (a.representation == 1) {
    *f0 = a.poly_off[p];
    *f1 = a.poly_off[p + 1];
    *v0 = a.vert_off[p];
  } else {
    *f0 = a.vert_off[p];
    *f1 = a.vert_off[p + 1];
    *v0 = a.vert_off[p] + a.poly_off[p];
  }
}

// Centroid of a surface whose faces all have zero area. Polygons: the running centroid stays the LAST polygon's
// (polygon_surface_mesh.cc:210-217). Triangles: the area-weighted sum stays the zero vector of the frame the mesh was
// built in (triangle_surface_mesh.h:456-486) -- frame M of a mesh pair, carried to the world by TransformVertices
// (p_WMo), the world frame itself for the half-space pairs.
__device__ __forceinline__ void zero_area_centroid(const SurfaceArgs& a, int p, long long f1, double* c) {
  if (a.representation == 1) {
    for (int d = 0; d < 3; ++d) c[d] = a.centroids[3 * (f1 - 1) + d];
    return;
  }
  const PairRec& pr = a.pairs[p];
  const bool mesh_pair = pr.kind == kSoftRigid || pr.kind == kSoftSoft;
  for (int d = 0; d < 3; ++d) c[d] = mesh_pair ? a.poses[pr.geom_a].m[9 + d] : 0.0;
}

__global__ void __launch_bounds__(128) k_surface(SurfaceArgs a) {
  if (a.counters->error != 0) return;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long first_surface = a.pair_begin == 0 ? 0 : a.surf_index[a.pair_begin];
  const long long num_surfaces = a.surf_index[a.pair_end];
  double warp_area = 0;  // lane 0: total area of this warp's surfaces (one atomic per warp at the end)
  for (long long s = first_surface + (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < num_surfaces;
       s += (long long)gridDim.x * warps_per_block) {
    const int p = a.surf_pair[s];
    long long f0, f1, v0;
    surface_face_range(a, p, &f0, &f1, &v0);
    const bool big = f1 - f0 > kBigSurfaceFaces;  // summed by k_surface_big_sum / _finish
    double area = 0, cx = 0, cy = 0, cz = 0;
    for (long long f = f0 + lane; f < f1 && !big; f += 32) {
      const double w = a.areas[f];
      area += w;
      cx += w * a.centroids[3 * f];
      cy += w * a.centroids[3 * f + 1];
      cz += w * a.centroids[3 * f + 2];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      area += __shfl_xor_sync(0xffffffffu, area, d);
      cx += __shfl_xor_sync(0xffffffffu, cx, d);
      cy += __shfl_xor_sync(0xffffffffu, cy, d);
      cz += __shfl_xor_sync(0xffffffffu, cz, d);
    }
    if (lane == 0) {
      if (big) {
        a.big_list[atomicAdd(&a.counters->big_surfaces, 1ull)] = (int)s;  // accumulators start at 0 below
      } else if (area != 0.0) {
        cx /= area;
        cy /= area;
        cz /= area;
      } else {
        double c[3];
        zero_area_centroid(a, p, f1, c);
        cx = c[0];
        cy = c[1];
        cz = c[2];
      }
      const PairRec& pr = a.pairs[p];
      a.surf_face_off[s] = f0;
      a.surf_vert_off[s] = v0;
      a.surf_total_area[s] = area;
      a.surf_centroid[3 * s] = cx;
      a.surf_centroid[3 * s + 1] = cy;
      a.surf_centroid[3 * s + 2] = cz;
      a.surf_scene[s] = pr.scene;
      a.surf_id_m[s] = pr.id_m;
      a.surf_id_n[s] = pr.id_n;
      // bit0 HasGradE_M, bit1 HasGradE_N, bit2 swapped.
      a.surf_flags[s] = pr.kind == kSoftSoft ? 3 : (pr.swapped ? (2 | 4) : 1);
      if (!big) warp_area += area;
      if (a.scene_stats) {
        unsigned long long* st = (unsigned long long*)(a.scene_stats + 6ll * pr.scene);
        atomicAdd(st + 3, (unsigned long long)(f1 - f0));
        atomicAdd(st + 4, 1ull);
      }
    }
  }
  if (lane == 0 && warp_area != 0.0) atomicAdd(&a.counters->total_area, warp_area);
}

// Large surfaces (tens of thousands of faces: one warp would take milliseconds): every warp of the
// grid sums segments of 2048 faces and adds its partial sums to the surface's accumulators.
__global__ void __launch_bounds__(128) k_surface_big_sum(SurfaceArgs a) {
  if (a.counters->error != 0) return;
  const int lane = threadIdx.x & 31;
  const long long num_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nbig = (long long)a.counters->big_surfaces;
  long long unit0 = 0;  // units (segments) before the current surface, to spread the warps evenly
  for (long long k = 0; k < nbig; ++k) {
    const int s = a.big_list[k];
    long long f0, f1, v0;
    surface_face_range(a, a.surf_pair[s], &f0, &f1, &v0);
    const long long nseg = (f1 - f0 + kBigSegment - 1) / kBigSegment;
    // First segment of this surface that falls to this warp.
    long long j = (gwarp - unit0 % num_warps + num_warps) % num_warps;
    for (; j < nseg; j += num_warps) {
      const long long b = f0 + j * kBigSegment, e = min(b + kBigSegment, f1);
      double area = 0, cx = 0, cy = 0, cz = 0;
      for (long long f = b + lane; f < e; f += 32) {
        const double w = a.areas[f];
        area += w;
        cx += w * a.centroids[3 * f];
        cy += w * a.centroids[3 * f + 1];
        cz += w * a.centroids[3 * f + 2];
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        area += __shfl_xor_sync(0xffffffffu, area, d);
        cx += __shfl_xor_sync(0xffffffffu, cx, d);
        cy += __shfl_xor_sync(0xffffffffu, cy, d);
        cz += __shfl_xor_sync(0xffffffffu, cz, d);
      }
      if (lane == 0) {
        atomicAdd(&a.surf_total_area[s], area);
        atomicAdd(&a.surf_centroid[3 * s], cx);
        atomicAdd(&a.surf_centroid[3 * s + 1], cy);
        atomicAdd(&a.surf_centroid[3 * s + 2], cz);
      }
    }
    unit0 += nseg;
  }
}

__global__ void k_surface_big_finish(SurfaceArgs a) {
  if (a.counters->error != 0) return;
  const long long nbig = (long long)a.counters->big_surfaces;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nbig; k += (long long)gridDim.x * blockDim.x) {
    const int s = a.big_list[k];
    const double area = a.surf_total_area[s];
    if (area != 0.0) {
      a.surf_centroid[3 * s] /= area;
      a.surf_centroid[3 * s + 1] /= area;
      a.surf_centroid[3 * s + 2] /= area;
    } else {
      long long f0, f1, v0;
      surface_face_range(a, a.surf_pair[s], &f0, &f1, &v0);
      zero_area_centroid(a, a.surf_pair[s], f1, &a.surf_centroid[3 * s]);
    }
    atomicAdd(&a.counters->total_area, area);
  }
}

// Between chunks: folds the chunk-local list cursors into totals / maxima and zeroes them.
__global__ void k_next_chunk(Counters* c) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long fr = c->frontier_max;
    for (int l = 0; l < 64; ++l) fr = max(fr, c->frontier[l]);
    c->frontier_max = max(fr, max(c->surv_tri, c->surv_tet));  // the survivor lists live in the frontier buffers
    c->leaf_tri_total += c->leaf_tri;
    c->leaf_tet_total += c->leaf_tet;
    c->polys_total += c->polys;
    c->slow_total += c->slow;
    c->leaf_tri_max = max(c->leaf_tri_max, c->leaf_tri);
    c->leaf_tet_max = max(c->leaf_tet_max, c->leaf_tet);
    c->polys_max = max(c->polys_max, c->polys);
    c->slow_max = max(c->slow_max, c->slow);
    c->undecided_max = max(c->undecided_max, c->undecided);
    c->small_items_max = max(c->small_items_max, max(c->small_items[0], c->small_items[1]));
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x < 64) c->frontier[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
      c->leaf_tri = c->leaf_tet = c->polys = c->slow = c->surv_tri = c->surv_tet = c->big_surfaces = c->work_tri = c->work_tet =
          c->plane_emit = c->small_pairs[0] = c->small_pairs[1] = c->small_work[0] = c->small_work[1] = c->undecided = 0;
      c->small_items[0] = c->small_items[1] = c->small_items_work[0] = c->small_items_work[1] = 0;
    }
  }
}

// Writes the sentinel offsets and the face / vertex totals.
__global__ void k_finalize(SurfaceArgs a, int num_pairs) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long polys = a.poly_off[num_pairs], pverts = a.vert_off[num_pairs];
  const long long faces = a.representation == 1 ? polys : pverts;
  const long long verts = a.representation == 1 ? pverts : pverts + polys;
  const long long ns = (long long)a.counters->surfaces;
  a.surf_face_off[ns] = faces;
  a.surf_vert_off[ns] = verts;
  a.counters->faces = (unsigned long long)faces;
  a.counters->vertices = (unsigned long long)verts;
  a.counters->poly_verts = (unsigned long long)pverts;
  a.counters->polygons = (unsigned long long)polys;
}

// Per-scene statistics that depend only on the static pair list.
__global__ void k_scene_stats(const PairRec* pairs, int num_pairs, long long* scene_stats) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= num_pairs) return;
  const PairRec& pr = pairs[p];
  if (pr.kind > kSoftSoft && pr.kind != kSoftMeshRigidPlane && pr.kind != kSoftPlaneRigidMesh) return;
  unsigned long long* st = (unsigned long long*)(scene_stats + 6ll * pr.scene);
  atomicAdd(st + 0, 1ull);
}

// ---------------------------------------------------------------------------
// Half-space pairs (CalcRigidCompliant's special cases, hydroelastic_calculator.cc:25-41). There is
// one mesh per pair, so no dual-tree descent: k_plane_setup poses the half space in the mesh's frame
// and tests the mesh's root box against it; k_plane<false> runs one thread per element of every pair
// whose root box straddles the boundary, counts its polygon and lists the elements that emit one;
// after the scan k_plane<true> runs the same arithmetic again for the listed elements and writes
// the polygon straight into its final slot (the polygons are at most quadrilaterals out of one record
// each: recomputing is cheaper than staging them). Vertices are per
// polygon; the reference welds the ones cut from the same mesh edge (cut_edges,
// mesh_plane_intersection.cc:205-231): face quantities do not depend on that.
struct PlaneArgs {
  const PairRec* pairs;
  const Xf* poses;
  Xf* pair_X;                       // per pair: m[0..2] unit normal, m[3] displacement (mesh frame), m[4] active,
                                    // m[5] pressure scale, m[6] margin
  Pools pools;
  const double* tet_vp;             // [4 * pool tet]: pressure at the tetrahedron's vertices
  const int* plane_pairs;           // static list of the half-space pairs (pair indices, ascending)
  const long long* plane_elem_off;  // [num_plane_pairs + 1]: work items before each of them (its mesh's elements, padded
                                    // to whole blocks of kPlaneBlock leaves)
  const int* plane_nelem;           // [num_plane_pairs]: elements of the pair's mesh
  const int* plane_block_base;      // [num_plane_pairs]: pool index of the mesh's first leaf-block box
  const int* order_tet;             // per pool position: the asset-local element at that place of the BVH's leaf order
  const int* order_tri;
  const ElemBox* blocks;            // box of every run of kPlaneBlock consecutive leaves
  const double* geom_hs;            // [2 * flat geometry]: pressure scale, margin of a compliant half space
  int q_begin, q_end;               // this chunk's range of the list
  unsigned int* emit_list;          // [elements of the chunk's half-space pairs]: the ones that emit a polygon
                                    // (count pass -> pack pass; Counters::plane_emit entries)
  int* pair_polys;
  int* pair_verts;
  unsigned long long* pair_cursor;
  const long long* poly_off;
  const long long* vert_off;
  int representation;
  ResultArrays out;
  Counters* counters;
};

__global__ void k_plane_setup(PlaneArgs a) {
  const int q = a.q_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.q_end) return;
  const int p = a.plane_pairs[q];
  const PairRec pr = a.pairs[p];
  const Xf X_WS = a.poses[pr.geom_a], X_WR = a.poses[pr.geom_b];
  Xf rec;
#pragma unroll
  for (int k = 0; k < 12; ++k) rec.m[k] = 0.0;
  D3 nhat = make_d3(0, 0, 1);
  double disp = 0;
  bool active = true;
  const bool mesh_is_soft = pr.kind == kSoftMeshRigidPlane;
  if (mesh_is_soft) {
    rigid_half_space_in_mesh_frame(X_WS, X_WR, &nhat, &disp);
  } else {
    rec.m[5] = a.geom_hs[2 * pr.geom_a];
    rec.m[6] = a.geom_hs[2 * pr.geom_a + 1];
    const int st = compliant_half_space_in_mesh_frame(X_WR, X_WS, rec.m[6], &nhat, &disp);
    if (st) {
      raise_error(a.counters, st, p);
      active = false;
    }
  }
  // The mesh's root box against the boundary plane: every vertex strictly outside (both cases), or
  // every vertex strictly inside (slicing only): no polygon. Exact fp64 box, padded against rounding.
  const BvhNode& root = a.pools.nodes[mesh_is_soft ? pr.root_a : pr.root_b];
  const double s = dot(nhat, load3(root.c)) - disp;
  const double r = (fabs(nhat.x) * root.h[0] + fabs(nhat.y) * root.h[1]) + fabs(nhat.z) * root.h[2];
  const double pad = 1e-12 * (fabs(s) + r + fabs(disp)) + 1e-300;
  if (s - r > pad) active = false;
  if (mesh_is_soft && s + r < -pad) active = false;
  rec.m[0] = nhat.x;
  rec.m[1] = nhat.y;
  rec.m[2] = nhat.z;
  rec.m[3] = disp;
  rec.m[4] = active ? 1.0 : 0.0;
  a.pair_X[p] = rec;
  atomicAdd(&a.counters->bv_tests, 1ull);
  if (active) atomicAdd(&a.counters->root_overlaps, 1ull);
}

template <bool kPack>
__device__ __forceinline__ void plane_pass(const PlaneArgs& a) {
  if (kPack && a.counters->error != 0) return;
  const long long w0 = a.plane_elem_off[a.q_begin];
  const long long num_elems = a.plane_elem_off[a.q_end] - w0;
  // Count pass: every element of the chunk's half-space pairs. Pack pass: the emitting elements only,
  // from the list the count pass wrote.
  const long long count = kPack ? (long long)a.counters->plane_emit : num_elems;
  const long long rounded = (count + 31ll) & ~31ll;  // whole warps reach the collectives
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride) {
    WorldPoly P;
    P.n = 0;
    P.status = 0;
    int pair = -1, elem = 0;
    long long item = 0;
    bool mesh_is_soft = false, swapped = false, worked = false;
    if (i < count) {
      item = kPack ? (long long)a.emit_list[i] : i;
      const long long w = w0 + item;
      const int q = find_plane_pair(a.plane_elem_off, a.q_begin, a.q_end, w, w0, num_elems);
      pair = a.plane_pairs[q];
      const Xf rec = a.pair_X[pair];
      const int k = (int)(w - a.plane_elem_off[q]);   // position in the mesh's leaf order (k >= elements: padding)
      bool live = rec.m[4] != 0.0 && k < a.plane_nelem[q];
      const D3 nhat = make_d3(rec.m[0], rec.m[1], rec.m[2]);
      mesh_is_soft = a.pairs[pair].kind == kSoftMeshRigidPlane;
      if (!kPack && live) {
        // Bvh::Collide(plane) (bvh.h:285-309, mesh_plane_intersection.cc:310-330), flattened to one cut of the tree:
        // the box of this block of kPlaneBlock consecutive leaves against the boundary plane. All 128 threads of the
        // block look at the same box. Slicing (compliant mesh x rigid half space): a box entirely on either side has
        // no polygon; clipping (compliant half space x rigid mesh): only a box entirely outside.
        const ElemBox& bx = a.blocks[a.plane_block_base[q] + k / kPlaneBlock];
        const double s = dot(nhat, load3(bx.c)) - rec.m[3];
        const double r = (fabs(nhat.x) * bx.h[0] + fabs(nhat.y) * bx.h[1]) + fabs(nhat.z) * bx.h[2];
        const double pad = 1e-12 * (fabs(s) + r + fabs(rec.m[3])) + 1e-300;
        if (s - r > pad || (mesh_is_soft && s + r < -pad)) live = false;
      }
      if (live) {
        const PairRec pr = a.pairs[pair];
        elem = (mesh_is_soft ? a.order_tet[pr.elem_base_a + k] : a.order_tri[pr.elem_base_b + k]);
        swapped = pr.swapped != 0;
        if (mesh_is_soft) {
          const int t = pr.elem_base_a + elem;
          if (kPack) slice_tet_with_plane(a.pools.tets[t], a.tet_vp + 4ll * t, nhat, rec.m[3], a.poses[pr.geom_a], &P);
          else P.n = count_tet_slice(a.pools.tets[t], nhat, rec.m[3]);  // same heights, same sign code
        } else {
          clip_tri_by_half_space(a.pools.tris[pr.elem_base_b + elem], nhat, rec.m[3], rec.m[5], rec.m[6],
                                 a.poses[pr.geom_b], a.poses[pr.geom_a], &P);
        }
        worked = true;
        if (!kPack && P.status) raise_error(a.counters, P.status, pair);
      }
    }
    if (!kPack) {
      const unsigned tets = __ballot_sync(0xffffffffu, worked && mesh_is_soft);
      const unsigned tris = __ballot_sync(0xffffffffu, worked && !mesh_is_soft);
      if (lane == 0 && tets) atomicAdd(&a.counters->plane_tets, (unsigned long long)__popc(tets));
      if (lane == 0 && tris) atomicAdd(&a.counters->plane_tris, (unsigned long long)__popc(tris));
    }
    const int n = P.status ? 0 : P.n;
    const bool valid = n > 0;
    if (!kPack) {  // the emitting elements, for the pack pass (the list holds every element: it cannot overflow)
      const unsigned long long slot = warp_append_flag(&a.counters->plane_emit, valid);
      if (valid) a.emit_list[slot] = (unsigned int)item;
    }
    // The polygons of a warp mostly belong to one geometry pair: one lane updates the pair's counters
    // (or claims the warp's slots) for all of them.
    const unsigned validmask = __ballot_sync(0xffffffffu, valid);
    if (!validmask) continue;
    const int lead = __ffs(validmask) - 1;
    const int pair0 = __shfl_sync(0xffffffffu, pair, lead);
    const bool uniform = __all_sync(0xffffffffu, !valid || pair == pair0);
    if (!kPack) {
      if (uniform) {
        int total = valid ? n : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
        if (lane == lead) {
          atomicAdd(&a.pair_polys[pair0], __popc(validmask));
          atomicAdd(&a.pair_verts[pair0], total);
        }
      } else if (valid) {
        atomicAdd(&a.pair_polys[pair], 1);
        atomicAdd(&a.pair_verts[pair], n);
      }
      continue;
    }
    // hi 32 bits: polygons placed, lo 32: polygon vertices placed (as in k_pack).
    unsigned long long cur = 0;
    {
      const unsigned long long v = valid ? ((1ull << 32) | (unsigned long long)n) : 0ull;
      if (uniform) {
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long base = 0;
        if (lane == lead) base = atomicAdd(&a.pair_cursor[pair0], total);
        base = __shfl_sync(0xffffffffu, base, lead);
        cur = base + incl - v;
      } else if (valid) {
        cur = atomicAdd(&a.pair_cursor[pair], v);
      }
    }
    if (!valid) continue;
    const int pj = (int)(cur >> 32);
    const int vj = (int)(cur & 0xffffffffull);
    const bool polygons = a.representation == 1 /*DBH_POLYGON*/;
    long long f0, vb, fd0;
    int vloc;
    if (polygons) {
      f0 = a.poly_off[pair] + pj;
      vb = a.vert_off[pair] + vj;
      fd0 = f0 + vb;
      vloc = vj;
      if (f0 >= a.out.face_capacity || vb + n > a.out.vert_capacity || fd0 + n + 1 > a.out.face_data_capacity) {
        raise_error(a.counters, 5, 6);
        continue;
      }
    } else {
      f0 = a.vert_off[pair] + vj;
      vloc = vj + pj;
      vb = a.vert_off[pair] + a.poly_off[pair] + vloc;
      fd0 = 3 * f0;
      if (f0 + n > a.out.face_capacity || vb + n + 1 > a.out.vert_capacity || 3 * (f0 + n) > a.out.face_data_capacity) {
        raise_error(a.counters, 5, 6);
        continue;
      }
    }
    PackedPoly K;
    pack_world_polygon(P, a.representation, swapped, vloc, &K);
    for (int k = 0; k < K.nv; ++k) {
      store3(a.out.vertices, vb + k, K.v[k]);
      a.out.pressures[vb + k] = K.e[k];
      a.out.vertex_origin[vb + k] = k < P.n ? P.origin[k] : -1;
    }
    for (int k = 0; k < K.nfd; ++k) a.out.face_data[fd0 + k] = K.fd[k];
    const D3 zero = make_d3(0, 0, 0);
    for (int k = 0; k < K.nf; ++k) {
      const long long f = f0 + k;
      a.out.areas[f] = K.area[k];
      store3(a.out.normals, f, swapped ? -K.normal[k] : K.normal[k]);  // ReverseFaceWinding negates normals
      store3(a.out.centroids, f, K.centroid[k]);
      a.out.e_centroid[f] = K.e_c[k];
      // The compliant body is the first-constructed M; swapped: grad_eM <-> grad_eN (contact_surface.cc:58).
      store3(a.out.grad_eM, f, swapped ? zero : P.grad_W);
      store3(a.out.grad_eN, f, swapped ? P.grad_W : zero);
      a.out.elem_a[f] = mesh_is_soft ? elem : -1;
      a.out.elem_b[f] = mesh_is_soft ? -1 : elem;
      a.out.face_pair[f] = pair;
    }
  }
}

// The pack pass runs under a 128-register cap (4 resident blocks per SM): measured 1.5 % faster on the ball / plate /
// half-space workload than the uncapped build (140 registers); a count pass with register-accumulated statistics and
// block-aggregated list appends measured no different from this one and was dropped (profiles/README.md, r2a).
template <bool kPack>
__global__ void __launch_bounds__(128, kPack ? 4 : 1) k_plane(PlaneArgs a) {
  plane_pass<kPack>(a);
}
// ---------------------------------------------------------------------------
// k_wrench: HydroelasticTractionCalculator::ComputeSpatialForcesAtCentroidFromHydroelasticModel +
// ShiftSpatialForcesAtCentroidToBodyOrigins (hydroelastic_traction_calculator.cc:24-83) over the packed,
// device-resident results of the last query: one warp per contact surface, faces strided over the lanes,
// a shuffle reduction of the six components. A is the body of geometry M (id_M), B of geometry N.
// Parity-green on the B200 (tests/test_zz_gpu_wrench.py); timed as bench.py's e2e.spatial_forces.
// The per-face arithmetic is also checked on the host (tests/test_plane_math_cpu.py).
struct WrenchArgs {
  const int* surf_pair;
  const PairRec* pairs;
  const Xf* poses;
  const double* body_origin;   // [G*3] p_WBo per geometry's body, or null: the geometry frame's origin
  const double* body_vel;      // [G*6] V_WB = (w, v) at the body origin, or null: at rest
  const double* props;         // [G*3] hydroelastic modulus (+inf: rigid), Hunt-Crossley dissipation, dynamic friction
  const long long* surf_face_off;
  const long long* surf_vert_off;
  const double* surf_centroid;
  int representation;
  const double* areas;
  const double* normals;
  const double* centroids;
  const double* e_centroid;
  const double* vertices;
  const double* pressures;
  const int* face_data;
  double vslip_regularizer;
  double* forces;              // [num_surfaces * 18]: F_Ac_W, F_Ao_W, F_Bo_W, each (tau, f)
  Counters* counters;
};

__global__ void __launch_bounds__(128) k_wrench(WrenchArgs a) {
  if (a.counters->error != 0) return;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long num_surfaces = (long long)a.counters->surfaces;
  for (long long s = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); s < num_surfaces;
       s += (long long)gridDim.x * warps_per_block) {
    const PairRec pr = a.pairs[a.surf_pair[s]];
    const int gM = pr.swapped ? pr.geom_b : pr.geom_a, gN = pr.swapped ? pr.geom_a : pr.geom_b;
    TractionBodies d;
    d.p_WAo = a.body_origin ? load3(a.body_origin + 3ll * gM) : load3(a.poses[gM].m + 9);
    d.p_WBo = a.body_origin ? load3(a.body_origin + 3ll * gN) : load3(a.poses[gN].m + 9);
    const D3 zero = make_d3(0, 0, 0);
    d.w_WA = a.body_vel ? load3(a.body_vel + 6ll * gM) : zero;
    d.v_WA = a.body_vel ? load3(a.body_vel + 6ll * gM + 3) : zero;
    d.w_WB = a.body_vel ? load3(a.body_vel + 6ll * gN) : zero;
    d.v_WB = a.body_vel ? load3(a.body_vel + 6ll * gN + 3) : zero;
    d.p_WC = load3(a.surf_centroid + 3 * s);
    d.dissipation = combined_hunt_crossley_dissipation(a.props[3ll * gM], a.props[3ll * gN], a.props[3ll * gM + 1],
                                                       a.props[3ll * gN + 1]);
    d.mu_coulomb = combined_dynamic_friction(a.props[3ll * gM + 2], a.props[3ll * gN + 2]);
    d.vslip_regularizer = a.vslip_regularizer;
    const long long f0 = a.surf_face_off[s], f1 = a.surf_face_off[s + 1], v0 = a.surf_vert_off[s];
    Spatial6 acc{zero, zero};
    for (long long f = f0 + lane; f < f1; f += 32) {
      const D3 nhat = load3(a.normals + 3 * f);
      if (a.representation == 1 /*DBH_POLYGON*/) {
        acc = acc + polygon_face_force(d, a.areas[f], nhat, load3(a.centroids + 3 * f), a.e_centroid[f]);
      } else {
        const long long i0 = v0 + a.face_data[3 * f], i1 = v0 + a.face_data[3 * f + 1], i2 = v0 + a.face_data[3 * f + 2];
        acc = acc + triangle_face_force(d, a.areas[f], nhat, load3(a.vertices + 3 * i0), load3(a.vertices + 3 * i1),
                                        load3(a.vertices + 3 * i2), a.pressures[i0], a.pressures[i1], a.pressures[i2]);
      }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      acc.rot.x += __shfl_xor_sync(0xffffffffu, acc.rot.x, k);
      acc.rot.y += __shfl_xor_sync(0xffffffffu, acc.rot.y, k);
      acc.rot.z += __shfl_xor_sync(0xffffffffu, acc.rot.z, k);
      acc.trans.x += __shfl_xor_sync(0xffffffffu, acc.trans.x, k);
      acc.trans.y += __shfl_xor_sync(0xffffffffu, acc.trans.y, k);
      acc.trans.z += __shfl_xor_sync(0xffffffffu, acc.trans.z, k);
    }
    if (lane == 0) {
      Spatial6 F_Ao, F_Bo;
      shift_to_body_origins(d, acc, &F_Ao, &F_Bo);
      double* out = a.forces + 18 * s;
      store3(out, 0, acc.rot);
      store3(out, 1, acc.trans);
      store3(out, 2, F_Ao.rot);
      store3(out, 3, F_Ao.trans);
      store3(out, 4, F_Bo.rot);
      store3(out, 5, F_Bo.trans);
    }
  }
}

// ---------------------------------------------------------------------------
// k_discrete: the geometry-only half of DiscreteUpdateManager::AppendDiscreteContactPairsForHydroelasticContact
// (discrete_update_manager.cc:896-1036) over the packed, device-resident results of the last query: one thread per
// face, 9 doubles out per face (p_WC, nhat_BA_W, phi0, fn0, stiffness; stiffness == 0 marks a face the reference
// skips), in face order, so face f of surface s is record surf_face_offset[s] + f. 72 B per face leave the GPU
// instead of the ~250 B of the complete ContactSurface.
struct DiscreteArgs {
  const PairRec* pairs;
  const int* face_pair;
  const double* areas;
  const double* normals;
  const double* centroids;
  const double* e_centroid;
  const double* grad_eM;
  const double* grad_eN;
  double* out;             // [num_faces * 9]
  Counters* counters;
};

__global__ void __launch_bounds__(256) k_discrete(DiscreteArgs a) {
  if (a.counters->error != 0) return;
  const long long num_faces = (long long)a.counters->faces;
  unsigned int kept = 0;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < num_faces; f += (long long)gridDim.x * blockDim.x) {
    const PairRec& pr = a.pairs[a.face_pair[f]];
    // bit0 HasGradE_M, bit1 HasGradE_N as k_surface reports them: both for two compliant bodies; else the compliant
    // body is M unless the ids were swapped.
    const bool both = pr.kind == kSoftSoft;
    const bool has_M = both || !pr.swapped, has_N = both || pr.swapped;
    DiscreteFacePair d;
    kept += discrete_face_pair(a.areas[f], load3(a.normals + 3 * f), load3(a.centroids + 3 * f), a.e_centroid[f],
                               load3(a.grad_eM + 3 * f), load3(a.grad_eN + 3 * f), has_M, has_N, &d);
    double* o = a.out + 9 * f;
    store3(o, 0, d.p_WC);
    store3(o, 1, d.nhat_BA_W);
    o[6] = d.phi0;
    o[7] = d.fn0;
    o[8] = d.stiffness;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, k);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&a.counters->discrete_pairs, (unsigned long long)kept);
}

// ---- k_deformable: the deformable contact data of every face of the last query (hydro_wrench.cuh) -------------------
// One thread per face, kDeformableRecord = 24 doubles out.
//   tet-triangle surface (deformable surface x rigid pressure field):
//     p_WC[3], nhat_W[3], R_WC[9] (column-major), pressure, grad_W[3], barycentric[3] in the deformable triangle, 0, 0.
//     The deformable body is the pair's triangle side, whatever the order of the two ids (a swapped surface's normals
//     and gradients are swapped back here).
//   tet-tet surface (two deformable volumes, signed-distance fields):
//     p_WC[3], nhat_W[3], R_WC[9], signed distance, barycentric[4] in M's tetrahedron, barycentric[4] in N's.
//   Faces of half-space surfaces get NaN records.
constexpr int kDeformableRecord = 24;
struct DeformableArgs {
  const PairRec* pairs;
  const Xf* poses;
  const TriRec* tris;
  const TetRec* tets;
  const int* face_pair;
  const int* elem_a;
  const int* elem_b;
  const double* normals;
  const double* centroids;
  const double* e_centroid;
  const double* grad_eM;
  const double* grad_eN;
  double* out;             // [num_faces * kDeformableRecord]
  Counters* counters;
};

__global__ void __launch_bounds__(256) k_deformable(DeformableArgs a) {
  if (a.counters->error != 0) return;
  const long long num_faces = (long long)a.counters->faces;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < num_faces; f += (long long)gridDim.x * blockDim.x) {
    const PairRec& pr = a.pairs[a.face_pair[f]];
    double* o = a.out + kDeformableRecord * f;
    if (pr.kind == kSoftRigid) {
      const bool swapped = pr.swapped != 0;
      D3 n = load3(a.normals + 3 * f);
      if (swapped) n = -n;
      const D3 grad = load3((swapped ? a.grad_eN : a.grad_eM) + 3 * f);
      const TriRec& t = a.tris[pr.elem_base_b + a.elem_b[f]];
      const DeformableFaceContact c =
          deformable_face_contact(n, load3(a.centroids + 3 * f), a.e_centroid[f], grad, a.poses[pr.geom_a], a.poses[pr.geom_b],
                                  load3(t.v[0]), load3(t.v[1]), load3(t.v[2]));
      store3(o, 0, c.p_WC);
      store3(o, 1, c.nhat_W);
      for (int k = 0; k < 9; ++k) o[6 + k] = c.R_WC[k];
      o[15] = c.pressure;
      store3(o + 16, 0, c.grad_W);
      store3(o + 19, 0, c.barycentric);
      o[22] = o[23] = 0.0;
    } else if (pr.kind == kSoftSoft) {
      const TetRec& t0 = a.tets[pr.elem_base_a + a.elem_a[f]];
      const TetRec& t1 = a.tets[pr.elem_base_b + a.elem_b[f]];
      const DeformablePairContact c = deformable_pair_contact(load3(a.normals + 3 * f), load3(a.centroids + 3 * f), a.e_centroid[f],
                                                              a.poses[pr.geom_a], a.poses[pr.geom_b], t0.v, t1.v);
      store3(o, 0, c.p_WC);
      store3(o, 1, c.nhat_W);
      for (int k = 0; k < 9; ++k) o[6 + k] = c.R_WC[k];
      o[15] = c.signed_distance;
      for (int k = 0; k < 4; ++k) {
        o[16 + k] = c.bary0[k];
        o[20 + k] = c.bary1[k];
      }
    } else {
      const double kNaN = __longlong_as_double(0x7ff8000000000000ll);
      for (int k = 0; k < kDeformableRecord; ++k) o[k] = kNaN;
    }
  }
}

}  // namespace dbh
