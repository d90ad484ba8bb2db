// The pure polygon primitives of the exact stage (stage B of the narrow phase), shared by the clip kernels
// (hydro_kernels.cuh) and, compiled for the host, by the CPU test suite (tests/cpp/plane_host_check.cc):
//   RemoveNearlyDuplicateVertices   mesh_intersection.cc:121-170
//   ClipPolygonByHalfSpace          mesh_intersection.cc:76-119 (+ CalcIntersection :49-74)
//   the clip loop of IntersectTetrahedra   field_intersection.cc:146-182
//   IsFaceNormalInNormalDirection   contact_surface_utility.cc:336-354
// templated on the polygon storage (shared-memory scratch in the fast path, local arrays in the exceptional
// one). Moving them here did not change a single instruction of any kernel (profiles/*sass_fingerprints*).
#pragma once
#include "hydro_math.cuh"
#include "hydro_types.h"

namespace dbh {

constexpr double kCosAlphaDevice = -0.38268343236508973;  // cos(5*pi/8) as std::cos returns it

// The exceptional path's scratch: twice the capacity, in local memory. Only
// candidates whose intermediate polygon outgrows 8 vertices get here (a clip
// plane coplanar with the polygon makes the inside/outside pattern of its
// vertices arbitrary, so a clip can add more than one vertex).
struct LocalPoly {
  static constexpr int kCap = 2 * kMaxPolyVerts;
  D3 v[2][kCap];
  __host__ __device__ __forceinline__ D3 get(int buf, int slot) const { return v[buf][slot]; }
  __host__ __device__ __forceinline__ void set(int buf, int slot, D3 p) { v[buf][slot] = p; }
};

// RemoveNearlyDuplicateVertices (mesh_intersection.cc:121-170), in place.
template <class Poly>
__host__ __device__ __forceinline__ int remove_nearly_duplicate(Poly& P, int buf, int n) {
  if (n <= 1) return n;
  const double kEpsSquared = 1e-14 * 1e-14;
  // std::unique: keep the first of each run, comparing against the last kept vertex.
  int kept_count = 1;
  D3 kept = P.get(buf, 0);
  const D3 first_v = kept;
  for (int i = 1; i < n; ++i) {
    const D3 q = P.get(buf, i);
    const D3 d = kept - q;
    if (!(dot(d, d) < kEpsSquared)) {
      P.set(buf, kept_count, q);
      ++kept_count;
      kept = q;
    }
  }
  n = kept_count;
  if (n >= 3) {
    const D3 d = first_v - kept;  // polygon[0] - polygon.back()
    if (dot(d, d) < kEpsSquared) --n;
  }
  return n;
}

// ClipPolygonByHalfSpace (mesh_intersection.cc:76-119) from `buf` into `1-buf`.
// Half space: unit normal nh, displacement d (height = nh.p - d, inside <= 0).
// n >= 1.
template <class Poly>
__host__ __device__ __forceinline__ int clip_by_half_space(Poly& P, int buf, int n, D3 nh, double d) {
  int m = 0;
  D3 prev = P.get(buf, n - 1);
  double dp = dot(nh, prev) - d;
  for (int i = 0; i < n; ++i) {
    const D3 cur = P.get(buf, i);
    const double dc = dot(nh, cur) - d;
    const bool cur_in = dc <= 0, prev_in = dp <= 0;
    if (cur_in != prev_in) {
      // CalcIntersection(current, previous, H): a = height(current), b = height(previous).
      const double wa = dp / (dp - dc);
      const double wb = 1.0 - wa;
      if (m < Poly::kCap) P.set(1 - buf, m, wa * cur + wb * prev);
      ++m;
    }
    if (cur_in) {
      if (m < Poly::kCap) P.set(1 - buf, m, cur);
      ++m;
    }
    prev = cur;
    dp = dc;
  }
  return m;
}

// The same polygon clipped the way IntersectTetrahedra walks it
// (field_intersection.cc:146-182): edges (i, i+1), wa = d_j / (d_j - d_i),
// p = wa P_i + wb P_j. n >= 1.
template <class Poly>
__host__ __device__ __forceinline__ int clip_by_half_space_edges(Poly& P, int buf, int n, D3 nh, double d) {
  int m = 0;
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
      ++m;
      if (dj > 0) {
        const double wa = dj / (dj - di);
        const double wb = 1.0 - wa;
        if (m < Poly::kCap) P.set(1 - buf, m, wa * cur + wb * nxt);
        ++m;
      }
    } else if (dj <= 0) {
      const double wa = dj / (dj - di);
      const double wb = 1.0 - wa;
      if (m < Poly::kCap) P.set(1 - buf, m, wa * cur + wb * nxt);
      ++m;
    }
    cur = nxt;
    di = dj;
  }
  return m;
}

// cos(theta) > cos(5 pi / 8) for cos(theta) = ngrad . normalized(v), with ngrad
// already normalised (contact_surface_utility.cc:336-354). When |v| is 1 to
// within 1e-12 the normalisation cannot move the dot product by more than
// 1e-12, so outside an 1e-11 band around the threshold the comparison is
// decided without the sqrt and the three divisions; inside the band (and for
// any other |v|) the exact expression is evaluated.
__host__ __device__ __forceinline__ bool face_normal_along_gradient(D3 ngrad, D3 v) {
  const double s = dot(v, v);
  const double c0 = dot(ngrad, v);
  if (fabs(s - 1.0) < 1e-12) {
    if (c0 > kCosAlphaDevice + 1e-11) return true;
    if (c0 < kCosAlphaDevice - 1e-11) return false;
  }
  return dot(ngrad, normalized(v)) > kCosAlphaDevice;
}

}  // namespace dbh
