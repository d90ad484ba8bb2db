// Kernels of the surface-seeded soft-soft intersector (design and reference mapping: hydro_bfs.cuh).
//   k_bfs_seed     the tet-tet survivors the mid phase found over the two boundary trees enter the visited set;
//                  a pair reached through several boundary triangles is kept once (duplicates are switched off in place);
//   k_bfs_advance  the next wave = what the last one appended (two cursors, one thread);
//   k_bfs_wave     one thread per queued pair: CalcContactPolygon with face codes (tet_pair_polygon_with_faces), the
//                  polygon into the staging soup exactly as k_clip<tet> emits it (same emit_polygon: k_scan / k_pack /
//                  k_surface downstream are unchanged), then for each intersected face the neighbour across it
//                  (field_intersection.cc:327-369) -- one hash insert (atomicCAS, linear probing) per face, the new
//                  pairs appended to the queue with one atomic per warp and face.
// The queue is the tet-tet survivor list itself (the frontier buffer the descent has released): seeds at the front,
// every wave appends behind; its cursor is Counters::surv_tet, so the capacity protocol of the lists (overflow ->
// DBH_ERR_CAPACITY -> dbh_batch_grow_capacity) covers it. The visited set has at least four times the queue's capacity.
#pragma once
#include "hydro_bfs.cuh"
#include "hydro_kernels.cuh"

namespace dbh {

struct BfsArgs {
  ClipArgs c;                  // pools, pair_X, soup, pair_polys / pair_verts, counters (list pointers unused)
  const PairRec* pairs;
  int4* queue;                 // {pair, pool tet of M, pool tet of N, flags (bit0: to be processed, bit1: verify mode's mark)}
  unsigned long long queue_capacity;
  unsigned long long* table;   // visited set: kBfsEmpty or bfs_key(pair, local tet of M, local tet of N)
  unsigned long long table_mask;
  const int4* neighbors;       // per pool tetrahedron: asset-local neighbour across face f (-1: boundary)
  int first;                   // k_bfs_advance: 1 = the wave of the seeds
};

// The 64-bit finaliser of MurmurHash3: the keys of one batch differ in a few bit fields only (the scenes instance the
// same meshes), every output bit must depend on all of them.
__device__ __forceinline__ unsigned long long bfs_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// true: the key was not in the set (and now is).
__device__ __forceinline__ bool bfs_visit(const BfsArgs& a, unsigned long long key) {
  unsigned long long slot = bfs_hash(key) & a.table_mask;
  for (int probe = 0; probe < 100000; ++probe) {
    const unsigned long long old = atomicCAS(a.table + slot, kBfsEmpty, key);
    if (old == kBfsEmpty) return true;
    if (old == key) return false;
    slot = (slot + 1) & a.table_mask;
  }
  raise_error(a.c.counters, 5 /*DBH_ERR_CAPACITY*/, 12);   // cannot happen below 50 % load
  return false;
}

__global__ void __launch_bounds__(128) k_bfs_seed(BfsArgs a) {
  Counters* const ctr = a.c.counters;
  if (ctr->error != 0) return;
  const unsigned long long count = min(ctr->surv_tet, a.queue_capacity);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const int4 q = a.queue[i];
    const PairRec& pr = a.pairs[q.x];
    if (!bfs_visit(a, bfs_key(q.x, q.y - pr.elem_base_a, q.z - pr.elem_base_b))) a.queue[i].w = 0;
  }
}

__global__ void k_bfs_advance(BfsArgs a) {
  Counters* const ctr = a.c.counters;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long tail = min(ctr->surv_tet, a.queue_capacity);
  ctr->bfs_begin = a.first ? 0ull : ctr->bfs_end;
  ctr->bfs_end = tail;
  if (ctr->bfs_end > ctr->bfs_begin) {
    ++ctr->bfs_waves;
    ctr->bfs_pairs += ctr->bfs_end - ctr->bfs_begin;
  }
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_bfs_wave(BfsArgs a) {
  extern __shared__ double smem[];
  Counters* const ctr = a.c.counters;
  // One decision per warp (the emit and the appends below are warp-wide collectives).
  if (__shfl_sync(0xffffffffu, (threadIdx.x & 31) == 0 ? ctr->error : 0, 0) != 0) return;
  PolyScratchT<kThreads, 8> P{smem + threadIdx.x};
  const unsigned long long begin = ctr->bfs_begin, end = min(ctr->bfs_end, a.queue_capacity);
  for (unsigned long long base = begin + (unsigned long long)blockIdx.x * kThreads; base < end;
       base += (unsigned long long)gridDim.x * kThreads) {
    const unsigned long long i = base + threadIdx.x;
    int4 q = make_int4(0, 0, 0, 0);
    if (i < end) q = a.queue[i];
    const bool active = (q.w & 1) != 0;
    TetPairPolygon r;
    r.status = 0;
    r.n = 0;
    r.buf = 0;
    r.faces = 0;
    r.grad0 = make_d3(0, 0, 0);
    r.nhat_M = make_d3(0, 0, 0);
    r.f0_Mo = 0;
    if (active) {
      const Xf X = a.c.pair_X[q.x];
      const TetRec& tet0 = a.c.pools.tets[q.y];
      const TetRec& tet1 = a.c.pools.tets[q.z];
      const TetAux& aux0 = a.c.pools.aux[q.y];
      const TetAux& aux1 = a.c.pools.aux[q.z];
      tet_pair_polygon_with_faces(X, tet0, tet1, aux0, aux1, P, &r);
      if (r.status == 1) {
        // More than 8 intermediate vertices (a clip plane coplanar with the polygon): again on the exceptional path's
        // storage, emitted with its per-lane emit.
        LocalPoly L;
        tet_pair_polygon_with_faces(X, tet0, tet1, aux0, aux1, L, &r);
        if (r.status == 1) {
          raise_error(ctr, 10 /*DBH_ERR_POLYGON_SIZE*/, q.x);
        } else if (r.status == 0) {
          atomicAdd(&ctr->slow_total, 1ull);
          emit_polygon_slow(a.c, L, r.buf, r.n, q, r.grad0, r.f0_Mo, r.nhat_M);
        }
        r.n = 0;   // nothing left for the warp-wide emit; r.faces stays
      }
      if (r.status == 9) {
        raise_error(ctr, 9 /*DBH_ERR_PLANE_NORMAL*/, q.x);
        r.n = 0;
        r.faces = 0;
      }
    }
    emit_polygon(a.c, P, r.buf, r.n, q, r.grad0, r.f0_Mo, r.nhat_M);   // all 32 lanes
    // The neighbours across the intersected faces (field_intersection.cc:327-369). The four hash inserts of one
    // tetrahedron's faces are issued together (first probe; a collision continues alone), and the warp's new pairs
    // are appended with ONE atomic on the queue cursor (a single address for the whole grid: one atomic per face and
    // warp ran the first version of this kernel at the L2's same-address rate).
    int4 nb_a = make_int4(-1, -1, -1, -1), nb_b = nb_a;
    int base_a = 0, base_b = 0;
    if (active && r.faces) {
      const PairRec& pr = a.pairs[q.x];
      base_a = pr.elem_base_a;
      base_b = pr.elem_base_b;
      if (r.faces & 0x0Fu) nb_a = a.neighbors[q.y];
      if (r.faces & 0xF0u) nb_b = a.neighbors[q.z];
    }
    const int local_a = q.y - base_a, local_b = q.z - base_b;
    unsigned new_mask = 0;   // bit f: the neighbour across face f joins the queue
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int4 nb = half ? nb_b : nb_a;
      const int nbv[4] = {nb.x, nb.y, nb.z, nb.w};
      unsigned long long key[4], old[4], slot[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool valid = active && ((r.faces >> (4 * half + k)) & 1u) && nbv[k] >= 0;
        key[k] = kBfsEmpty;
        old[k] = 0;
        slot[k] = 0;
        if (valid) {
          key[k] = half ? bfs_key(q.x, local_a, nbv[k]) : bfs_key(q.x, nbv[k], local_b);
          slot[k] = bfs_hash(key[k]) & a.table_mask;
          old[k] = atomicCAS(a.table + slot[k], kBfsEmpty, key[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (key[k] == kBfsEmpty) continue;
        bool is_new = old[k] == kBfsEmpty;
        if (!is_new && old[k] != key[k]) {   // the slot holds another key: linear probing from the next one
          unsigned long long s = slot[k];
          bool settled = false;
          for (int probe = 0; probe < 100000 && !settled; ++probe) {
            s = (s + 1) & a.table_mask;
            const unsigned long long o = atomicCAS(a.table + s, kBfsEmpty, key[k]);
            if (o == kBfsEmpty) {
              is_new = true;
              settled = true;
            } else if (o == key[k]) {
              settled = true;
            }
          }
          if (!settled) raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 12);   // cannot happen below 50 % load
        }
        if (is_new) new_mask |= 1u << (4 * half + k);
      }
    }
    {
      const int lane = threadIdx.x & 31;
      const int mine = __popc(new_mask);
      int incl = mine;   // inclusive prefix sum over the warp
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long slot0 = 0;
      if (lane == 0 && total) slot0 = atomicAdd(&ctr->surv_tet, (unsigned long long)total);
      slot0 = __shfl_sync(0xffffffffu, slot0, 0) + (unsigned long long)(incl - mine);
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        if (!((new_mask >> f) & 1u)) continue;
        int4 next = q;
        next.w = 1;
        const int4 nb = f < 4 ? nb_a : nb_b;
        const int k = f & 3;
        const int neighbor = k == 0 ? nb.x : (k == 1 ? nb.y : (k == 2 ? nb.z : nb.w));
        if (f < 4) next.y = base_a + neighbor; else next.z = base_b + neighbor;
        if (slot0 < a.queue_capacity) a.queue[slot0] = next;
        else raise_error(ctr, 5 /*DBH_ERR_CAPACITY*/, 1);
        ++slot0;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && end > begin) atomicAdd(&ctr->full_clips_tet, end - begin);
}

}  // namespace dbh
