// Oracle (test infrastructure; never linked into the product): the reference's Bvh<Obb, Mesh>, the hierarchy its
// hydroelastic geometries actually carry (hydroelastic_internal.h:71-100,268: SoftMesh / RigidMesh hold a
// Bvh<Obb, ...>), restated from
//   obb.cc:25-39      Obb ctor + Obb::HasOverlap(a, b, X_GH) = BoxesOverlap(a.hw, b.hw, X_GA^-1 (X_GH X_HB))
//   obb.cc:124-131    PadBoundary (kTolerance = 2e-14, obb.h:162)
//   obb.cc:133-209    ObbMaker::CalcOrientationByPca
//   obb.cc:211-263    ObbMaker::CalcOrientedBox
//   obb.cc:265-297    ObbMaker::CalcVolumeGradient
//   obb.cc:303-355    ObbMaker::OptimizeObbVolume (one gradient direction, 20-step line search) and Compute
//   bvh.cc:18-136     the top-down build: median split of the element centroids along the box's longest axis
//   bvh.h:220-262     Collide (shared with ho_bvh.h)
// so that the CPU baseline's box-test count is the reference's. One deliberate difference: the reference
// takes the principal axes from Eigen's closed-form SelfAdjointEigenSolver::computeDirect; Eigen is absent here, so
// the symmetric 3x3 eigenproblem is solved by cyclic Jacobi rotations. Both return an orthonormal eigenbasis of the
// same covariance matrix; where eigenvalues repeat either basis is arbitrary (the reference's own issue #14067), so
// individual boxes -- and with them the box-test count -- can differ from Drake's by a small fraction. The candidate
// element pairs the traversal yields are a superset of the contact-producing pairs either way; the surfaces do not
// depend on the hierarchy (tests/test_oracle_kat.py).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <set>
#include <utility>
#include <vector>

#include "ho_bvh.h"
#include "ho_math.h"
#include "ho_mesh.h"

namespace ho {

struct Obb {
  Xform pose;  // X_HB: the box frame B in the hierarchy frame H
  Vec3 half_width;
  double Volume() const { return 8.0 * half_width.x * half_width.y * half_width.z; }  // Obb::CalcVolume (obb.h:99-103)
};

// Obb ctor + PadBoundary (obb.cc:25-33,124-131).
inline Obb MakeObb(const Xform& X_HB, Vec3 half_width) {
  const Vec3& c = X_HB.p;
  const double max_position = std::max({std::fabs(c.x), std::fabs(c.y), std::fabs(c.z)});
  const double max_half_width = std::max({half_width.x, half_width.y, half_width.z});
  const double scale = std::max(max_position, max_half_width);
  const double incr = std::max(scale * std::numeric_limits<double>::epsilon(), 2e-14);
  return Obb{X_HB, {half_width.x + incr, half_width.y + incr, half_width.z + incr}};
}

// Obb::HasOverlap(a_G, b_H, X_GH) (obb.cc:35-41).
inline bool ObbHasOverlap(const Obb& a, const Obb& b, const Xform& X_GH) {
  const Xform X_AB = InvertAndCompose(a.pose, X_GH * b.pose);
  return BoxesOverlap(a.half_width, b.half_width, X_AB.R, X_AB.p);
}

// RotationMatrix(RollPitchYaw(r, p, y)) = Rz(y) Ry(p) Rx(r) (math/roll_pitch_yaw.h ToRotationMatrix).
inline Mat3 RpyMatrix(double r, double p, double y) {
  const double c0 = std::cos(r), c1 = std::cos(p), c2 = std::cos(y);
  const double s0 = std::sin(r), s1 = std::sin(p), s2 = std::sin(y);
  Mat3 R;
  R(0, 0) = c2 * c1;
  R(0, 1) = c2 * s1 * s0 - s2 * c0;
  R(0, 2) = c2 * s1 * c0 + s2 * s0;
  R(1, 0) = s2 * c1;
  R(1, 1) = s2 * s1 * s0 + c2 * c0;
  R(1, 2) = s2 * s1 * c0 - c2 * s0;
  R(2, 0) = -s1;
  R(2, 1) = c1 * s0;
  R(2, 2) = c1 * c0;
  return R;
}

// Eigenvectors of a symmetric 3x3 matrix by cyclic Jacobi, sorted by ASCENDING eigenvalue like Eigen's solver.
inline void SymmetricEigen3(double A[3][3], double values[3], double vectors[3][3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * (diag + 1e-300)) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
  for (int k = 0; k < 3; ++k) {
    values[k] = A[order[k]][order[k]];
    for (int r = 0; r < 3; ++r) vectors[r][k] = V[r][order[k]];
  }
}

class ObbMaker {
 public:
  ObbMaker(const std::vector<Vec3>& verts, const std::set<int>& vertices) : verts_(verts), vertices_(vertices) {}

  // obb.cc:133-209.
  Mat3 CalcOrientationByPca() const {
    const double one_over_n = 1.0 / static_cast<double>(vertices_.size());
    Vec3 p_MC{0, 0, 0};
    for (int v : vertices_) p_MC += verts_[v];
    p_MC = p_MC * one_over_n;
    double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int v : vertices_) {
      const Vec3 d = verts_[v] - p_MC;
      const double a[3] = {d.x, d.y, d.z};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cov[r][c] += a[r] * a[c];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r][c] *= one_over_n;
    double values[3], vec[3][3];
    SymmetricEigen3(cov, values, vec);
    const Vec3 Bx{vec[0][2], vec[1][2], vec[2][2]};  // largest eigenvalue first
    const Vec3 By{vec[0][1], vec[1][1], vec[2][1]};
    const Vec3 Bz = cross(Bx, By);
    Mat3 R;
    R(0, 0) = Bx.x; R(1, 0) = Bx.y; R(2, 0) = Bx.z;
    R(0, 1) = By.x; R(1, 1) = By.y; R(2, 1) = By.z;
    R(0, 2) = Bz.x; R(1, 2) = Bz.y; R(2, 2) = Bz.z;
    return R;
  }

  // obb.cc:211-263.
  Obb CalcOrientedBox(const Mat3& R_MB) const {
    const double inf = std::numeric_limits<double>::infinity();
    Vec3 lo{inf, inf, inf}, hi{-inf, -inf, -inf};
    const Mat3 R_FM = transpose(R_MB);
    for (int v : vertices_) {
      const Vec3 p = R_FM * verts_[v];
      lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
      hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
    }
    const Vec3 half_width = (hi - lo) / 2.;
    const Vec3 p_FBo = (hi + lo) / 2.;
    Xform X;
    X.R = R_MB;
    X.p = R_MB * p_FBo;
    return MakeObb(X, half_width);
  }

  // obb.cc:265-297.
  Vec3 CalcVolumeGradient(const Obb& box) const {
    const double volume_0 = box.Volume();
    const Mat3& R_MB = box.pose.R;
    constexpr double h = 5. * M_PI / 180.;
    constexpr double one_over_h = 1.0 / h;
    const double volume_roll = CalcOrientedBox(R_MB * RpyMatrix(h, 0., 0.)).Volume();
    const double volume_pitch = CalcOrientedBox(R_MB * RpyMatrix(0., h, 0.)).Volume();
    const double volume_yaw = CalcOrientedBox(R_MB * RpyMatrix(0., 0., h)).Volume();
    return {(volume_roll - volume_0) * one_over_h, (volume_pitch - volume_0) * one_over_h,
            (volume_yaw - volume_0) * one_over_h};
  }

  // obb.cc:303-347.
  Obb OptimizeObbVolume(const Obb& box0) const {
    const Vec3 dV = CalcVolumeGradient(box0);
    const double len = norm(dV);
    if (len <= 1e-9) return box0;
    const Mat3& R_MB0 = box0.pose.R;
    const double volume0 = box0.Volume();
    double volume = volume0;
    Obb box = box0;
    const double kMinVolumeImprovement = 0.001;
    double increment = 0.1 * volume0 / len;
    const double min_increment = increment / 1000000.;
    double step = 0.;
    for (int i = 0; i < 20; ++i) {
      step -= increment;
      const Obb try_box = CalcOrientedBox(R_MB0 * RpyMatrix(step * dV.x, step * dV.y, step * dV.z));
      const double try_volume = try_box.Volume();
      if (try_volume < volume) {
        const double improvement = (volume - try_volume) / volume;
        volume = try_volume;
        box = try_box;
        if (improvement < kMinVolumeImprovement) break;
        increment *= 1.5;
        continue;
      }
      step += increment;
      if (increment <= min_increment) break;
      increment /= 10.;
    }
    return box.Volume() < volume0 ? box : box0;
  }

  // obb.cc:349-355.
  Obb Compute() const { return OptimizeObbVolume(CalcOrientedBox(CalcOrientationByPca())); }

 private:
  const std::vector<Vec3>& verts_;
  const std::set<int>& vertices_;
};

struct ObbNode {
  Obb bv;
  int left{-1}, right{-1};
  int num_index{0};
  int indices[3]{0, 0, 0};
  bool is_leaf() const { return left < 0; }
};

// Bvh<Obb, Mesh> (bvh.cc:18-136).
struct ObbBvh {
  std::vector<ObbNode> nodes;  // nodes[0] is the root
  int max_per_leaf{1};

  template <int kVerts>
  void Build(const std::vector<Vec3>& verts, const std::vector<std::array<int, kVerts>>& elements, int max_elements_per_leaf) {
    max_per_leaf = max_elements_per_leaf;
    std::vector<std::pair<int, Vec3>> centroids;
    centroids.reserve(elements.size());
    for (int i = 0; i < static_cast<int>(elements.size()); ++i) {
      Vec3 c{0, 0, 0};
      for (int v = 0; v < kVerts; ++v) c += verts[elements[i][v]];
      centroids.emplace_back(i, c / kVerts);
    }
    nodes.clear();
    nodes.reserve(2 * elements.size());
    BuildTree<kVerts>(verts, elements, centroids, 0, static_cast<int>(centroids.size()));
  }

  template <int kVerts>
  int BuildTree(const std::vector<Vec3>& verts, const std::vector<std::array<int, kVerts>>& elements,
                std::vector<std::pair<int, Vec3>>& c, int start, int end) {
    std::set<int> vertices;
    for (int i = start; i < end; ++i)
      for (int v = 0; v < kVerts; ++v) vertices.insert(elements[c[i].first][v]);
    const int me = static_cast<int>(nodes.size());
    nodes.emplace_back();
    nodes[me].bv = ObbMaker(verts, vertices).Compute();
    const int n = end - start;
    if (n <= max_per_leaf) {
      nodes[me].num_index = n;
      for (int i = 0; i < n; ++i) nodes[me].indices[i] = c[start + i].first;
      return me;
    }
    // The axis of greatest spread is the box axis with the largest half width, expressed in the mesh frame
    // (bvh.cc:83-91).
    int axis = 0;
    const Vec3 hw = nodes[me].bv.half_width;
    if (hw.y > hw[axis]) axis = 1;
    if (hw.z > hw[axis]) axis = 2;
    const Mat3& R = nodes[me].bv.pose.R;
    const Vec3 Baxis{R(0, axis), R(1, axis), R(2, axis)};
    std::sort(c.begin() + start, c.begin() + end, [&Baxis](const std::pair<int, Vec3>& a, const std::pair<int, Vec3>& b) {
      return dot(Baxis, a.second) < dot(Baxis, b.second);
    });
    const int mid = start + n / 2;
    const int l = BuildTree<kVerts>(verts, elements, c, start, mid);
    const int r = BuildTree<kVerts>(verts, elements, c, mid, end);
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
};

inline ObbBvh MakeObbBvh(const VolumeMesh& m) {
  ObbBvh b;
  b.Build<4>(m.verts, m.tets, 1);
  return b;
}
inline ObbBvh MakeObbBvh(const TriMesh& m) {
  ObbBvh b;
  b.Build<3>(m.verts, m.tris, 3);
  return b;
}

// Bvh::Collide (bvh.h:220-262) over two OBB hierarchies.
inline void CollideObb(const ObbBvh& A, const ObbBvh& B, const Xform& X_AB, std::vector<std::pair<int, int>>* out,
                       int64_t* bv_tests = nullptr) {
  std::vector<std::pair<int, int>> stack;
  stack.emplace_back(0, 0);
  int64_t tests = 0;
  while (!stack.empty()) {
    const auto [ia, ib] = stack.back();
    stack.pop_back();
    const ObbNode& a = A.nodes[ia];
    const ObbNode& b = B.nodes[ib];
    ++tests;
    if (!ObbHasOverlap(a.bv, b.bv, X_AB)) continue;
    if (a.is_leaf() && b.is_leaf()) {
      for (int i = 0; i < a.num_index; ++i)
        for (int j = 0; j < b.num_index; ++j) out->emplace_back(a.indices[i], b.indices[j]);
    } else if (b.is_leaf()) {
      stack.emplace_back(a.left, ib);
      stack.emplace_back(a.right, ib);
    } else if (a.is_leaf()) {
      stack.emplace_back(ia, b.left);
      stack.emplace_back(ia, b.right);
    } else {
      stack.emplace_back(a.left, b.left);
      stack.emplace_back(a.right, b.left);
      stack.emplace_back(a.left, b.right);
      stack.emplace_back(a.right, b.right);
    }
  }
  if (bv_tests) *bv_tests += tests;
}

}  // namespace ho
