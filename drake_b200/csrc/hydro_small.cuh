// k_small_pairs: the mid phase of SMALL geometry pairs, fused and kept on chip.
//
// The generic pipeline (hydro_kernels.cuh) walks the dual BVH of every geometry pair of a chunk together,
// one launch per tree level, through frontier lists in global memory: the right shape for a pair of
// 10^6-tetrahedron meshes (millions of node pairs per level), the wrong one for the batched configs, whose
// pairs are tiny (ball_plate: 512 x 24, 512 x 44, 44 x 24 elements; ~1500 box tests and ~320 leaf pairs per
// pair): ten launches per query whose frontier and leaf lists make a round trip through L2 / HBM.
//
// Here ONE WARP owns one geometry pair from the root box pair to the survivor list:
//   * Bvh::Collide (bvh.h:220-262) as a depth-first dual descent over a stack of node pairs in the warp's
//     SHARED MEMORY: every iteration the 32 lanes pop the 32 topmost node pairs and run ONE conservative fp32
//     box test each (the six inequalities of k_traverse's overlap6_f32, branch-free); a lane whose pair
//     overlaps pushes its 2 or 4 child pairs back untested (ballot + popc: no atomics, no barriers) -- leaf
//     pairs go to a leaf buffer in shared memory. One test per lane keeps the lanes full whatever the mix of
//     inner nodes and leaves (the first version expanded 2 x 2 children per lane: half its test slots idled
//     once the smaller tree had bottomed out; profiles/README.md);
//   * whenever 32 leaf pairs are waiting, the lanes run the leaf predicate (BoxesOverlap on the elements'
//     own boxes, boxes_overlap.cc:197-255: fp32 on the leaves' node boxes with a two-sided margin, exact
//     fp64 on the exact boxes when undecided -- those few go to a list that k_leaf_undecided settles: the fp64
//     predicate's registers and stack frame would cost every warp of this kernel a fifth of its occupancy) and the
//     conservative stage-A culls of the narrow phase (the same tri_stage_a / tet_stage_a as k_filter) on
//     them, and stage the survivors in shared memory;
//   * full groups of 32 survivors go to the global survivor list (one atomic + one coalesced 512 B store),
//     which k_clip consumes exactly as it consumes k_filter's.
// Warps claim pairs dynamically (one atomic per pair). Nothing but the survivors (16 B each) and, in
// recording mode, the candidate list leaves the SM. The pair's pose (PairF32) is warp-uniform and lives in
// registers.
//
// A pair is SMALL when both trees have at most 32767 nodes (node / element indices local to the asset fit
// 15 bits, so a stack entry is 4 bytes). The node-pair stack is LIFO, so it stays near
// (levels x 3 x 32) entries; should a pose ever need more than the configured capacity the query fails with
// DBH_ERR_CAPACITY (detail 9) and dbh_batch_grow_capacity() doubles it (or hands the batch's pairs to the
// generic pipeline), the library's usual protocol.
#pragma once
#include "hydro_kernels.cuh"

namespace dbh {

constexpr int kSmallWarps = 4;        // per block
constexpr int kSmallLeafCap = 64;     // 31 waiting + 32 from one round of box tests, rounded up
constexpr int kSmallSurvCap = 64;     // 31 waiting + 32 from one filter batch, rounded up
constexpr int kSmallMaxNodes = 32767; // per asset: stack codes are 15 bits + a leaf flag

struct SmallArgs {
  ClipArgs c;                 // pools, pair_X, pair_f32, counters, verify (its list pointers are unused here)
  const PairRec* pairs;
  const int* small_list;      // overlapping small pairs of this chunk and kind (written by k_pairs):
                              // Counters::small_pairs[kind] entries
  int4* surv;                 // survivor list of this kind (shared with k_filter: same cursor)
  unsigned long long surv_capacity;
  int4* cand;                 // recording mode: every leaf pair {pair, pool elem a, pool elem b, candidate flag}
  unsigned long long cand_capacity;
  int record;
  int stack_cap;              // node-pair stack entries per warp
  int2* items;                // kPass 1 writes, kPass 2 reads: {pair, node pair} work items of the large small pairs
  unsigned long long items_capacity;
  int split_at;               // pass 1: a large pair's descent stops when this many node pairs wait
};

__host__ __device__ constexpr int small_smem_bytes_per_warp(int stack_cap) {
  return stack_cap * 4 + kSmallLeafCap * 4 + kSmallSurvCap * 16;
}

// The conservative fp32 box test of the descent, branch-free: the same six inequalities as overlap6_f32 (the face
// axes of both boxes; outward-rounded boxes plus a 4e-6 relative pad), all evaluated and OR-ed -- in a warp whose
// 32 lanes test 32 different node pairs some lane survives every axis anyway, so early exits only add divergence.
__device__ __forceinline__ bool small_overlap6(const Node32& A, const Node32& B, const PairF32& f) {
  const float tx = fmaf(f.R[0], B.c[0], fmaf(f.R[3], B.c[1], fmaf(f.R[6], B.c[2], f.p[0]))) - A.c[0];
  const float ty = fmaf(f.R[1], B.c[0], fmaf(f.R[4], B.c[1], fmaf(f.R[7], B.c[2], f.p[1]))) - A.c[1];
  const float tz = fmaf(f.R[2], B.c[0], fmaf(f.R[5], B.c[1], fmaf(f.R[8], B.c[2], f.p[2]))) - A.c[2];
  const float delta = 4e-6f * (A.l1 + (B.l1 + f.S));
  const float ax = A.h[0] + delta, ay = A.h[1] + delta, az = A.h[2] + delta;
  const float bx = B.h[0] + delta, by = B.h[1] + delta, bz = B.h[2] + delta;
  bool sep = fabsf(tx) > ax + fmaf(bx, f.AR[0], fmaf(by, f.AR[1], bz * f.AR[2]));
  sep |= fabsf(ty) > ay + fmaf(bx, f.AR[3], fmaf(by, f.AR[4], bz * f.AR[5]));
  sep |= fabsf(tz) > az + fmaf(bx, f.AR[6], fmaf(by, f.AR[7], bz * f.AR[8]));
  sep |= fabsf(fmaf(tx, f.R[0], fmaf(ty, f.R[1], tz * f.R[2]))) > bx + fmaf(ax, f.AR[0], fmaf(ay, f.AR[3], az * f.AR[6]));
  sep |= fabsf(fmaf(tx, f.R[3], fmaf(ty, f.R[4], tz * f.R[5]))) > by + fmaf(ax, f.AR[1], fmaf(ay, f.AR[4], az * f.AR[7]));
  sep |= fabsf(fmaf(tx, f.R[6], fmaf(ty, f.R[7], tz * f.R[8]))) > bz + fmaf(ax, f.AR[2], fmaf(ay, f.AR[5], az * f.AR[8]));
  return !sep;
}

// A leaf's fp32 node box as the leaf predicate's operand. The node box contains the element's exact box and
// exceeds it by at most |c - c32| + 2 ulp per half width (see the upload): far inside the predicate's 1e-5
// relative decision margin, which is all boxes_overlap_f32 asks of its operands.
__device__ __forceinline__ BoxF small_leaf_box(const Node32& n) {
  BoxF b;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b.c[d] = n.c[d];
    b.h[d] = n.h[d];
  }
  b.l1 = n.l1;
  return b;
}

// kPass: 0 the whole descent of every pair of the work list (ball_plate);
//        1 the same, except that a LARGE small pair (PairRec::small == 2: thousands of nodes, up to 10^5 box tests -- one
//          warp would work on it long after the others have run out of pairs: dense clutter ran at half its occupancy)
//          is only descended until SmallArgs::split_at node pairs are waiting, which then go to the item list
//          {pair, node pair};
//        2 the items: the same descent started from the item's node pair instead of the root pair.
template <bool kTet, int kMinBlocks, int kPass>
__global__ void __launch_bounds__(kSmallWarps * 32, kMinBlocks) k_small_pairs(SmallArgs a) {
  extern __shared__ __align__(16) unsigned char small_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned char* mine = small_smem + (size_t)warp * small_smem_bytes_per_warp(a.stack_cap);
  int4* const surv = reinterpret_cast<int4*>(mine);                       // 16 B aligned first
  unsigned* const stack = reinterpret_cast<unsigned*>(mine + kSmallSurvCap * 16);
  unsigned* const leaf = stack + a.stack_cap;
  Counters* const ctr = a.c.counters;
  // The query already failed: every warp takes the same decision for itself (no block-wide collectives here).
  if (__shfl_sync(0xffffffffu, lane == 0 ? ctr->error : 0, 0) != 0) return;
  const int count = (int)min(kPass == 2 ? min(ctr->small_items[kTet], a.items_capacity) : ctr->small_pairs[kTet],
                             (unsigned long long)0x7fffffff);
  unsigned n_tests = 0, n_cand = 0, n_exact = 0, n_leaf = 0;   // warp-uniform
  const Node32* const nodes = a.c.pools.nodes32;
  int4* const surv_list = a.surv;
  unsigned long long* const surv_cursor = kTet ? &ctr->surv_tet : &ctr->surv_tri;
  int fill = 0;       // staged survivors: carried from pair to pair, so the list is written 32 entries at a time
  auto flush = [&](int n) {   // the first n staged survivors (n <= 32) go to the global list
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(surv_cursor, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < n) {
      if (base + lane < a.surv_capacity) surv_list[base + lane] = surv[lane];
      else raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 1);
    }
  };
  bool failed = false;
  for (;;) {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(kPass == 2 ? &ctr->small_items_work[kTet] : &ctr->small_work[kTet], 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= count) break;
    int p;
    unsigned start = 1u | (1u << 16);   // the root pair
    if constexpr (kPass == 2) {
      const int2 item = a.items[t];
      p = item.x;
      start = (unsigned)item.y;
    } else {
      p = a.small_list[t];
    }
    const PairRec pr = a.pairs[p];
    const PairF32 f = a.c.pair_f32[p];
    const Node32* const nodes_a = nodes + (pr.root_a - 1);   // node 0 of each asset (1 is the root)
    const Node32* const nodes_b = nodes + (pr.root_b - 1);
    // A stack entry is a node pair still to be TESTED: asset-local node indices (15 bits each). The root pair
    // passed the exact test in k_pairs; it is simply tested again in fp32 (which cannot reject it).
    if (lane == 0) stack[0] = start;
    int top = 1, nleaf = 0;
    bool split = kPass == 1 && pr.small == 2;   // warp-uniform
    __syncwarp();
    while (top > 0 || nleaf > 0) {
      if (kPass == 1 && split && top >= a.split_at) {
        // Hand the waiting node pairs to the item list (the leaf pairs found so far are finished here).
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctr->small_items[kTet], (unsigned long long)top);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int k = lane; k < top; k += 32) {
          if (base + k < a.items_capacity) a.items[base + k] = make_int2(p, (int)stack[k]);
          else raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 11);
        }
        __syncwarp();
        top = 0;
        split = false;
        continue;
      }
      if (nleaf >= 32 || top == 0) {
        // ---- leaf predicate + stage A on up to 32 waiting leaf pairs -------------------------------
        const int cnt = min(32, nleaf);
        nleaf -= cnt;
        bool keep = false, is_cand = false, undecided = false;
        int4 c = make_int4(p, 0, 0, 0);
        if (lane < cnt) {
          const unsigned e = leaf[nleaf + lane];
          const Node32 la = nodes_a[e & 0xffffu], lb = nodes_b[e >> 16];
          c.y = ~la.child;   // pool elements
          c.z = ~lb.child;
          const int decided = boxes_overlap_f32(small_leaf_box(la), small_leaf_box(lb), f);
          undecided = decided == 0;   // 2.6 % of ball_plate's leaf pairs: handed to k_leaf_undecided (fp64)
          is_cand = decided > 0;
          if (is_cand) {
            const TetCullA ta = a.c.pools.cull_a[c.y];
            if (kTet) keep = tet_stage_a(a.c, c, f, ta, a.c.pools.cull_a[c.z]);
            else keep = tri_stage_a(a.c, c, f, ta, a.c.pools.cull_tri[c.z]);
          }
          c.w = is_cand ? 1 : 0;
          if (is_cand && !keep && a.c.verify) {
            keep = true;
            c.w |= 2;
          }
        }
        n_cand += __popc(__ballot_sync(0xffffffffu, is_cand));
        const unsigned um = __ballot_sync(0xffffffffu, undecided);
        n_exact += __popc(um);
        n_leaf += cnt;
        int cand_index = 0;   // + 1; 0: not recording
        if (a.record) {   // the candidate list of the parity tests (dbh_batch_copy_candidates)
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(kTet ? &ctr->leaf_tet : &ctr->leaf_tri, (unsigned long long)cnt);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (lane < cnt) {
            if (base + lane < a.cand_capacity) {
              a.cand[base + lane] = make_int4(c.x, c.y, c.z, c.w & 1);
              cand_index = (int)(base + lane) + 1;
            } else {
              raise_error(ctr, 5, 2);
            }
          }
        }
        if (um) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(&ctr->undecided, (unsigned long long)__popc(um));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (undecided) {
            const unsigned long long slot = base + __popc(um & lt);
            if (slot < a.c.undecided_capacity) a.c.undecided[slot] = make_int4(c.x, c.y, c.z, cand_index);
            else raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 10);
          }
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) surv[fill + __popc(km & lt)] = c;
        fill += __popc(km);
        __syncwarp();
        if (fill >= 32) {
          flush(32);
          __syncwarp();
          int4 moved = make_int4(0, 0, 0, 0);
          if (lane < fill - 32) moved = surv[32 + lane];
          __syncwarp();
          if (lane < fill - 32) surv[lane] = moved;
          fill -= 32;
          __syncwarp();
        }
      } else {
        // ---- one box test per lane on the up to 32 topmost node pairs ---------------------------------
        const int cnt = min(32, top);
        top -= cnt;
        n_tests += cnt;
        unsigned e = 0, first = 0;   // first: the first of this lane's new entries
        int kids = 0;                // 0, 2 or 4 new entries
        bool leaf_hit = false;
        if (lane < cnt) {
          e = stack[top + lane];
          const Node32 A = nodes_a[e & 0xffffu], B = nodes_b[e >> 16];
          if (small_overlap6(A, B, f)) {
            const bool la = A.child < 0, lb = B.child < 0;
            leaf_hit = la && lb;
            // Children are adjacent: (ca, cb), (ca + 1, cb), (ca, cb + 1), (ca + 1, cb + 1); a leaf stays as it is
            // while the other side descends. (Descending only the larger of two inner boxes -- 2 child pairs instead
            // of 4 -- runs as many box tests on ball_plate in more iterations: 4 % slower, profiles/README.md.)
            const unsigned ca = la ? (e & 0xffffu) : (unsigned)(A.child - (pr.root_a - 1));
            const unsigned cb = lb ? (e >> 16) : (unsigned)(B.child - (pr.root_b - 1));
            first = ca | (cb << 16);
            kids = leaf_hit ? 0 : ((la || lb) ? 2 : 4);
            if (!leaf_hit && lb) kids |= 8;   // two entries that differ in a (bit 3: step of the second entry)
          }
        }
        const unsigned ml = __ballot_sync(0xffffffffu, leaf_hit);
        const unsigned m2 = __ballot_sync(0xffffffffu, (kids & 7) == 2);
        const unsigned m4 = __ballot_sync(0xffffffffu, (kids & 7) == 4);
        if (leaf_hit) leaf[nleaf + __popc(ml & lt)] = e;
        nleaf += __popc(ml);
        const int pushed = 2 * __popc(m2) + 4 * __popc(m4);
        __syncwarp();   // every lane has read its entry before the slots are written again (the ballots do not order memory)
        if (top + pushed > a.stack_cap) {
          failed = true;   // warp-uniform
        } else {
          unsigned* dst = stack + top + 2 * __popc(m2 & lt) + 4 * __popc(m4 & lt);
          if ((kids & 7) == 4) {
            dst[0] = first;
            dst[1] = first + 1u;
            dst[2] = first + 0x10000u;
            dst[3] = first + 0x10001u;
          } else if ((kids & 7) == 2) {
            dst[0] = first;
            dst[1] = first + ((kids & 8) ? 1u : 0x10000u);
          }
          top += pushed;
        }
        __syncwarp();
        if (failed) {
          if (lane == 0) raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 9);
          top = 0;
          nleaf = 0;
        }
      }
    }
    if (failed) break;
  }
  if (fill > 0 && !failed) flush(fill);
  if (lane == 0) {
    if (n_tests) atomicAdd(&ctr->bv_tests, (unsigned long long)n_tests);
    if (kTet && n_tests) atomicAdd(&ctr->bv_tests_tet, (unsigned long long)n_tests);
    if (n_cand) atomicAdd(kTet ? &ctr->cand_tet : &ctr->cand_tri, (unsigned long long)n_cand);
    if (n_exact) atomicAdd(&ctr->leaf_exact, (unsigned long long)n_exact);
    // (Recording mode counts the leaf pairs through the leaf-list cursors.)
    if (!a.record && n_leaf) atomicAdd(kTet ? &ctr->small_leaf_tet : &ctr->small_leaf_tri, (unsigned long long)n_leaf);
  }
}

// ---- k_leaf_undecided: the leaf pairs k_small_pairs / k_filter could not decide in fp32 ----------------------
// One thread per entry: the exact fp64 leaf predicate (BoxesOverlap on the elements' exact boxes), then -- for a
// candidate -- the same stage-A culls, and the survivor joins the same survivor list (order within the list is free).
struct UndecidedArgs {
  ClipArgs c;                 // pools, pair_X, pair_f32, counters, verify
  const PairRec* pairs;
  int4* surv[2];              // [0] tet-tri, [1] tet-tet
  unsigned long long surv_capacity;
  int4* cand[2];              // recording mode: the entries' flags are set here
};

__global__ void __launch_bounds__(128) k_leaf_undecided(UndecidedArgs a) {
  Counters* const ctr = a.c.counters;
  if (ctr->error != 0) return;
  const unsigned long long count = min(ctr->undecided, a.c.undecided_capacity);
  const unsigned long long rounded = (count + 31ull) & ~31ull;
  unsigned n_cand[2] = {0, 0};
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < rounded;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    bool keep = false, is_cand = false, tet = false;
    int4 c = make_int4(0, 0, 0, 0);
    if (i < count) {
      const int4 e = a.c.undecided[i];
      c = make_int4(e.x, e.y, e.z, 0);
      tet = a.pairs[e.x].kind == kSoftSoft;
      is_cand = tet ? leaf_predicate_exact<true>(a.c, c) : leaf_predicate_exact<false>(a.c, c);
      if (is_cand) {
        const PairF32 f = a.c.pair_f32[c.x];
        const TetCullA ta = a.c.pools.cull_a[c.y];
        if (tet) keep = tet_stage_a(a.c, c, f, ta, a.c.pools.cull_a[c.z]);
        else keep = tri_stage_a(a.c, c, f, ta, a.c.pools.cull_tri[c.z]);
        c.w = 1;
        if (!keep && a.c.verify) {
          keep = true;
          c.w |= 2;
        }
        if (e.w) a.cand[tet][e.w - 1].w = 1;
      }
    }
    n_cand[0] += __popc(__ballot_sync(0xffffffffu, is_cand && !tet));
    n_cand[1] += __popc(__ballot_sync(0xffffffffu, is_cand && tet));
    const unsigned long long s0 = warp_append_flag(&ctr->surv_tri, keep && !tet);
    const unsigned long long s1 = warp_append_flag(&ctr->surv_tet, keep && tet);
    if (keep) {
      const unsigned long long slot = tet ? s1 : s0;
      if (slot < a.surv_capacity) a.surv[tet][slot] = c;
      else raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 1);
    }
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_cand[0]) atomicAdd(&ctr->cand_tri, (unsigned long long)n_cand[0]);
    if (n_cand[1]) atomicAdd(&ctr->cand_tet, (unsigned long long)n_cand[1]);
  }
}

}  // namespace dbh
