// fp64 3-vector helpers for the device kernels. The evaluation order of every
// expression is part of the parity contract with the reference's scalar CPU
// path (no FMA contraction: this translation unit is compiled with
// -fmad=false; dot products associate left to right; division and sqrt are
// IEEE-rounded). See DESIGN.md "Arithmetic contract".
#pragma once
#include "hydro_types.h"

namespace dbh {

struct D3 {
  double x, y, z;
};

__host__ __device__ __forceinline__ D3 make_d3(double x, double y, double z) { return D3{x, y, z}; }
__host__ __device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a) { return D3{-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ D3 operator*(double s, D3 a) { return D3{s * a.x, s * a.y, s * a.z}; }
__host__ __device__ __forceinline__ D3 operator/(D3 a, double s) { return D3{a.x / s, a.y / s, a.z / s}; }
__host__ __device__ __forceinline__ double dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ D3 cross(D3 a, D3 b) {
  return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ __forceinline__ double norm(D3 a) { return sqrt(dot(a, a)); }
// Eigen normalized(): the zero vector maps to itself.
__host__ __device__ __forceinline__ D3 normalized(D3 a) {
  const double z = dot(a, a);
  if (z > 0) return a / sqrt(z);
  return a;
}
__host__ __device__ __forceinline__ D3 load3(const double* p) { return D3{p[0], p[1], p[2]}; }

// R (column-major in X.m[0..8]) times v.
__host__ __device__ __forceinline__ D3 rot(const Xf& X, D3 v) {
  return D3{(X.m[0] * v.x + X.m[3] * v.y) + X.m[6] * v.z, (X.m[1] * v.x + X.m[4] * v.y) + X.m[7] * v.z,
            (X.m[2] * v.x + X.m[5] * v.y) + X.m[8] * v.z};
}
// R^T times v.
__host__ __device__ __forceinline__ D3 rot_t(const Xf& X, D3 v) {
  return D3{(X.m[0] * v.x + X.m[1] * v.y) + X.m[2] * v.z, (X.m[3] * v.x + X.m[4] * v.y) + X.m[5] * v.z,
            (X.m[6] * v.x + X.m[7] * v.y) + X.m[8] * v.z};
}
// p_AQ = p_AB + R_AB p_BQ (math/rigid_transform.h:848-850).
__host__ __device__ __forceinline__ D3 xform_point(const Xf& X, D3 v) {
  const D3 r = rot(X, v);
  return D3{X.m[9] + r.x, X.m[10] + r.y, X.m[11] + r.z};
}
// X_AC = X_BA^-1 X_BC (math/rigid_transform.h:760-770; scalar variant of
// fast_pose_composition_functions.cc:637-668).
__host__ __device__ __forceinline__ Xf invert_and_compose(const Xf& X_BA, const Xf& X_BC) {
  Xf X_AC;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      X_AC.m[c * 3 + r] = (X_BA.m[r * 3 + 0] * X_BC.m[c * 3 + 0] + X_BA.m[r * 3 + 1] * X_BC.m[c * 3 + 1]) +
                          X_BA.m[r * 3 + 2] * X_BC.m[c * 3 + 2];
    }
  }
  const double dx = X_BC.m[9] - X_BA.m[9], dy = X_BC.m[10] - X_BA.m[10], dz = X_BC.m[11] - X_BA.m[11];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    X_AC.m[9 + r] = (X_BA.m[r * 3 + 0] * dx + X_BA.m[r * 3 + 1] * dy) + X_BA.m[r * 3 + 2] * dz;
  }
  return X_AC;
}

// World box of a geometry the way FCL's CollisionObject::computeAABB builds the boxes of its broad phase: the local
// box translated when the rotation is the identity, else the cube around the transformed box centre that holds the
// box's bounding sphere. A half space (z <= 0) is bounded -- from above, in z -- only under the identity rotation.
__host__ __device__ __forceinline__ void world_box(const GeomBound& g, const Xf& X, double mn[3], double mx[3]) {
  const double kInf = 1.0 / 0.0;
  bool identity = true;
#pragma unroll
  for (int k = 0; k < 9; ++k) identity = identity && fabs(X.m[k] - ((k % 4 == 0) ? 1.0 : 0.0)) <= 1e-12;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mn[k] = -kInf;
    mx[k] = kInf;
  }
  if (g.kind == kBoundHalfSpace) {
    if (identity) mx[2] = X.m[11];
  } else if (g.kind == kBoundBox) {
    if (identity) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        mn[k] = X.m[9 + k] + (g.c[k] - g.h[k]);
        mx[k] = X.m[9 + k] + (g.c[k] + g.h[k]);
      }
    } else {
      const D3 c = xform_point(X, make_d3(g.c[0], g.c[1], g.c[2]));
      const double r = norm(make_d3(g.h[0], g.h[1], g.h[2]));
      mn[0] = c.x - r; mx[0] = c.x + r;
      mn[1] = c.y - r; mx[1] = c.y + r;
      mn[2] = c.z - r; mx[2] = c.z + r;
    }
  }
}
__host__ __device__ __forceinline__ bool world_boxes_overlap(const GeomBound& ga, const Xf& Xa, const GeomBound& gb,
                                                             const Xf& Xb) {
  double amn[3], amx[3], bmn[3], bmx[3];
  world_box(ga, Xa, amn, amx);
  world_box(gb, Xb, bmn, bmx);
  bool overlap = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) overlap = overlap && amn[k] <= bmx[k] && bmn[k] <= amx[k];
  return overlap;
}

// Gottschalk 15-axis SAT between boxes with half widths a, b under (R, t) =
// pose of b's frame in a's frame: the scalar BoxesOverlap
// (boxes_overlap.cc:197-255). X.m[0..8] = R column-major, t given separately.
__host__ __device__ __forceinline__ bool boxes_overlap(const double a[3], const double b[3], const Xf& X, D3 tv) {
  const double kEpsilon = 0.000001;
  const double t[3] = {tv.x, tv.y, tv.z};
  double ar[3][3];  // abs_r(i, j)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) ar[i][j] = fabs(X.m[j * 3 + i]) + kEpsilon;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (fabs(t[i]) > a[i] + ((b[0] * ar[i][0] + b[1] * ar[i][1]) + b[2] * ar[i][2])) return false;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (fabs((t[0] * X.m[i * 3 + 0] + t[1] * X.m[i * 3 + 1]) + t[2] * X.m[i * 3 + 2]) >
        b[i] + ((a[0] * ar[0][i] + a[1] * ar[1][i]) + a[2] * ar[2][i]))
      return false;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      if (fabs(t[i2] * X.m[j * 3 + i1] - t[i1] * X.m[j * 3 + i2]) >
          ((a[i1] * ar[i2][j] + a[i2] * ar[i1][j]) + b[j1] * ar[i][j2]) + b[j2] * ar[i][j1])
        return false;
    }
  }
  return true;
}

}  // namespace dbh
