// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference's hydroelastic contact-surface path. Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, load or call anything under oracle/. The product
// (drake_b200/) never includes or links these files.
//
// Small fixed-size linear algebra with an explicitly pinned evaluation order.
// The reference uses Eigen (Eigen 5.0.1, absent from /root/reference); its
// 3-vector reductions are restated here as plain left-to-right sums with no
// FMA contraction (build with -ffp-contract=off), cf. SURVEY.md §7 "Eigen
// reduction order is not visible here".
#pragma once
#include <cmath>
#include <cstdint>

namespace ho {

struct Vec3 {
  double x{0}, y{0}, z{0};
  Vec3() = default;
  Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(const Vec3& a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(const Vec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(const Vec3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3& operator+=(Vec3& a, const Vec3& b) { a = a + b; return a; }

inline double dot(const Vec3& a, const Vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double squared_norm(const Vec3& a) { return dot(a, a); }
inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
// Eigen's normalized(): the zero vector maps to itself (relied on at
// polygon_surface_mesh.cc:189-193).
inline Vec3 normalized(const Vec3& a) {
  const double z = squared_norm(a);
  if (z > 0) return a / std::sqrt(z);
  return a;
}

// 3x3 rotation, column-major like Drake's RotationMatrix (math/rotation_matrix.h).
struct Mat3 {
  double m[9]{1, 0, 0, 0, 1, 0, 0, 0, 1};  // m[c*3 + r]
  double operator()(int r, int c) const { return m[c * 3 + r]; }
  double& operator()(int r, int c) { return m[c * 3 + r]; }
};

inline Vec3 operator*(const Mat3& R, const Vec3& v) {
  return {(R(0, 0) * v.x + R(0, 1) * v.y) + R(0, 2) * v.z,
          (R(1, 0) * v.x + R(1, 1) * v.y) + R(1, 2) * v.z,
          (R(2, 0) * v.x + R(2, 1) * v.y) + R(2, 2) * v.z};
}
inline Mat3 transpose(const Mat3& R) {
  Mat3 T;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T(r, c) = R(c, r);
  return T;
}
inline Mat3 operator*(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C(r, c) = (A(r, 0) * B(0, c) + A(r, 1) * B(1, c)) + A(r, 2) * B(2, c);
  return C;
}

// Rigid transform X_AB: 12 doubles = R column-major, then p (math/rigid_transform.h:1004-1007).
struct Xform {
  Mat3 R;
  Vec3 p;
  static Xform FromArray(const double* a) {
    Xform X;
    for (int i = 0; i < 9; ++i) X.R.m[i] = a[i];
    X.p = {a[9], a[10], a[11]};
    return X;
  }
  void ToArray(double* a) const {
    for (int i = 0; i < 9; ++i) a[i] = R.m[i];
    a[9] = p.x; a[10] = p.y; a[11] = p.z;
  }
};
// p_AQ = p_AB + R_AB * p_BQ (math/rigid_transform.h:848-850).
inline Vec3 operator*(const Xform& X, const Vec3& v) { return X.p + X.R * v; }
// X_AC = X_AB * X_BC (fast_pose_composition_functions.cc ComposeXXNoAlias).
inline Xform operator*(const Xform& A, const Xform& B) {
  Xform C;
  C.R = A.R * B.R;
  C.p = A.p + A.R * B.p;
  return C;
}
// this = X_BA, other = X_BC: returns X_AC = X_BA^-1 * X_BC
// (math/rigid_transform.h:760-770; scalar variant
// fast_pose_composition_functions.cc:637-668: R_AC = R_BA^T R_BC,
// p_AC = R_BA^T (p_BC - p_BA), dot products left to right).
inline Xform InvertAndCompose(const Xform& X_BA, const Xform& X_BC) {
  Xform X_AC;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      X_AC.R(r, c) = (X_BA.R(0, r) * X_BC.R(0, c) + X_BA.R(1, r) * X_BC.R(1, c)) +
                     X_BA.R(2, r) * X_BC.R(2, c);
  const Vec3 d = X_BC.p - X_BA.p;
  X_AC.p = {(X_BA.R(0, 0) * d.x + X_BA.R(1, 0) * d.y) + X_BA.R(2, 0) * d.z,
            (X_BA.R(0, 1) * d.x + X_BA.R(1, 1) * d.y) + X_BA.R(2, 1) * d.z,
            (X_BA.R(0, 2) * d.x + X_BA.R(1, 2) * d.y) + X_BA.R(2, 2) * d.z};
  return X_AC;
}
inline Xform Inverse(const Xform& X) {
  Xform I;
  I.R = transpose(X.R);
  I.p = I.R * (-X.p);
  return I;
}

}  // namespace ho
