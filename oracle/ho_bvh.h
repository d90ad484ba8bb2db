// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// Bounding volumes, box-box SAT, the reference-shaped Bvh<Aabb> and its
// dual-tree traversal, and the topology-independent leaf-level candidate
// predicate (SURVEY.md §7 "hard parts"):
//   geometry/proximity/aabb.{h,cc}, boxes_overlap.cc:197-255 (scalar variant),
//   bvh.h:220-262, bvh.cc:18-136.
#pragma once
#include <limits>
#include <algorithm>
#include <memory>
#include <set>
#include <utility>
#include <vector>

#include "ho_mesh.h"

namespace ho {

struct Aabb {
  Vec3 center, half_width;
};

// AabbMaker::Compute (aabb.cc:102-117) over an explicit vertex list.
template <typename It>
inline Aabb MakeAabb(const std::vector<Vec3>& verts, It begin, It end) {
  It it = begin;
  Vec3 mx = verts[*it], mn = mx;
  for (++it; it != end; ++it) {
    const Vec3& v = verts[*it];
    mn = {std::min(mn.x, v.x), std::min(mn.y, v.y), std::min(mn.z, v.z)};
    mx = {std::max(mx.x, v.x), std::max(mx.y, v.y), std::max(mx.z, v.z)};
  }
  Aabb b;
  b.center = (mn + mx) / 2;
  b.half_width = mx - b.center;
  return b;
}

// Scalar BoxesOverlap (boxes_overlap.cc:197-255): Gottschalk 15-axis SAT.
inline bool BoxesOverlap(const Vec3& a, const Vec3& b, const Mat3& r, const Vec3& t) {
  const double kEpsilon = 0.000001;
  double abs_r[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) abs_r[i][j] = std::abs(r(i, j)) + kEpsilon;
  for (int i = 0; i < 3; ++i) {
    if (std::abs(t[i]) > a[i] + ((b.x * abs_r[i][0] + b.y * abs_r[i][1]) + b.z * abs_r[i][2])) return false;
  }
  for (int i = 0; i < 3; ++i) {
    if (std::abs((t.x * r(0, i) + t.y * r(1, i)) + t.z * r(2, i)) >
        b[i] + ((a.x * abs_r[0][i] + a.y * abs_r[1][i]) + a.z * abs_r[2][i]))
      return false;
  }
  int i1 = 1;
  for (int i = 0; i < 3; ++i) {
    const int i2 = (i1 + 1) % 3;
    int j1 = 1;
    for (int j = 0; j < 3; ++j) {
      const int j2 = (j1 + 1) % 3;
      if (std::abs(t[i2] * r(i1, j) - t[i1] * r(i2, j)) >
          ((a[i1] * abs_r[i2][j] + a[i2] * abs_r[i1][j]) + b[j1] * abs_r[i][j2]) + b[j2] * abs_r[i][j1])
        return false;
      j1 = j2;
    }
    i1 = i2;
  }
  return true;
}

// Aabb::HasOverlap(a_G, b_H, X_GH) (aabb.cc:16-33).
inline bool AabbHasOverlap(const Aabb& a_G, const Aabb& b_H, const Xform& X_GH) {
  const Vec3 t = (X_GH * b_H.center) - a_G.center;
  return BoxesOverlap(a_G.half_width, b_H.half_width, X_GH.R, t);
}

// Per-element boxes in the element's own body frame.
inline std::vector<Aabb> ElementBoxes(const VolumeMesh& m) {
  std::vector<Aabb> out(m.tets.size());
  for (size_t e = 0; e < m.tets.size(); ++e) out[e] = MakeAabb(m.verts, m.tets[e].begin(), m.tets[e].end());
  return out;
}
inline std::vector<Aabb> ElementBoxes(const TriMesh& m) {
  std::vector<Aabb> out(m.tris.size());
  for (size_t e = 0; e < m.tris.size(); ++e) out[e] = MakeAabb(m.verts, m.tris[e].begin(), m.tris[e].end());
  return out;
}

// Reference-shaped Bvh<Aabb, Mesh> (bvh.cc:18-136): top-down, median split
// along the longest box axis, BV from the unique vertex set of the range.
struct BvhNode {
  Aabb bv;
  int left{-1}, right{-1};  // children (internal) or -1
  int num_index{0};
  int indices[3]{0, 0, 0};
  bool is_leaf() const { return left < 0; }
};

struct Bvh {
  std::vector<BvhNode> nodes;  // nodes[0] is the root
  int max_per_leaf{1};

  template <class Mesh, int kVerts>
  void Build(const Mesh& mesh, const std::vector<std::array<int, kVerts>>& elements, int max_elements_per_leaf) {
    max_per_leaf = max_elements_per_leaf;
    std::vector<std::pair<int, Vec3>> centroids;
    centroids.reserve(elements.size());
    for (int i = 0; i < static_cast<int>(elements.size()); ++i) {
      Vec3 c{0, 0, 0};
      for (int v = 0; v < kVerts; ++v) c += mesh.verts[elements[i][v]];
      c = c / kVerts;
      centroids.emplace_back(i, c);
    }
    nodes.clear();
    nodes.reserve(2 * elements.size());
    BuildTree<Mesh, kVerts>(mesh, elements, centroids, 0, static_cast<int>(centroids.size()));
  }

  template <class Mesh, int kVerts>
  int BuildTree(const Mesh& mesh, const std::vector<std::array<int, kVerts>>& elements,
                std::vector<std::pair<int, Vec3>>& c, int start, int end) {
    std::set<int> vertices;
    for (int i = start; i < end; ++i)
      for (int v = 0; v < kVerts; ++v) vertices.insert(elements[c[i].first][v]);
    const int me = static_cast<int>(nodes.size());
    nodes.emplace_back();
    nodes[me].bv = MakeAabb(mesh.verts, vertices.begin(), vertices.end());
    const int n = end - start;
    if (n <= max_per_leaf) {
      nodes[me].num_index = n;
      for (int i = 0; i < n; ++i) nodes[me].indices[i] = c[start + i].first;
      return me;
    }
    int axis = 0;
    const Vec3 hw = nodes[me].bv.half_width;
    if (hw.y > hw[axis]) axis = 1;
    if (hw.z > hw[axis]) axis = 2;
    std::sort(c.begin() + start, c.begin() + end,
              [axis](const std::pair<int, Vec3>& a, const std::pair<int, Vec3>& b) { return a.second[axis] < b.second[axis]; });
    const int mid = start + n / 2;
    const int l = BuildTree<Mesh, kVerts>(mesh, elements, c, start, mid);
    const int r = BuildTree<Mesh, kVerts>(mesh, elements, c, mid, end);
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
};

// BvhUpdater::Update (bvh_updater.h:45-103): the topology stays, every box is refitted bottom-up to the mesh's
// current vertex positions; Aabb::set_bounds (aabb.h:188-191) puts the centre at (lower + upper) / 2 and the half width
// at (upper - lower) / 2. Returns (lower, upper) of the node.
template <class Mesh, int kVerts>
inline std::pair<Vec3, Vec3> RefitBvh(Bvh* bvh, const Mesh& mesh, const std::vector<std::array<int, kVerts>>& elements,
                                      int node = 0) {
  const double inf = std::numeric_limits<double>::infinity();
  BvhNode& n = bvh->nodes[node];
  Vec3 lower{inf, inf, inf}, upper{-inf, -inf, -inf};
  if (n.is_leaf()) {
    for (int e = 0; e < n.num_index; ++e)
      for (int v = 0; v < kVerts; ++v) {
        const Vec3& p = mesh.verts[elements[n.indices[e]][v]];
        lower = {std::min(lower.x, p.x), std::min(lower.y, p.y), std::min(lower.z, p.z)};
        upper = {std::max(upper.x, p.x), std::max(upper.y, p.y), std::max(upper.z, p.z)};
      }
  } else {
    const auto l = RefitBvh<Mesh, kVerts>(bvh, mesh, elements, n.left);
    const auto r = RefitBvh<Mesh, kVerts>(bvh, mesh, elements, n.right);
    lower = {std::min(l.first.x, r.first.x), std::min(l.first.y, r.first.y), std::min(l.first.z, r.first.z)};
    upper = {std::max(l.second.x, r.second.x), std::max(l.second.y, r.second.y), std::max(l.second.z, r.second.z)};
  }
  bvh->nodes[node].bv.center = (lower + upper) / 2;
  bvh->nodes[node].bv.half_width = (upper - lower) / 2;
  // Aabb::lower() / upper() of the refitted box: what the parent reads.
  const Aabb& b = bvh->nodes[node].bv;
  return {b.center - b.half_width, b.center + b.half_width};
}

inline Bvh MakeBvh(const VolumeMesh& m) {
  Bvh b;
  b.Build<VolumeMesh, 4>(m, m.tets, 1);  // MeshTraits<VolumeMesh>::kMaxElementPerBvhLeaf (bvh.h:37-40)
  return b;
}
inline Bvh MakeBvh(const TriMesh& m, int max_per_leaf = 3) {
  Bvh b;
  b.Build<TriMesh, 3>(m, m.tris, max_per_leaf);  // bvh.h:32-35
  return b;
}

// Bvh::Collide (bvh.h:220-262): explicit stack, 4-way descent. `bv_tests`
// (optional) counts node-pair tests.
inline void Collide(const Bvh& A, const Bvh& B, const Xform& X_AB, std::vector<std::pair<int, int>>* out,
                    int64_t* bv_tests = nullptr) {
  std::vector<std::pair<int, int>> stack;
  stack.emplace_back(0, 0);
  int64_t tests = 0;
  while (!stack.empty()) {
    const auto [ia, ib] = stack.back();
    stack.pop_back();
    const BvhNode& a = A.nodes[ia];
    const BvhNode& b = B.nodes[ib];
    ++tests;
    if (!AabbHasOverlap(a.bv, b.bv, X_AB)) continue;
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

// Leaf-level candidate predicate, evaluated by brute force: pair (i, j) is a
// candidate iff the per-element AABBs pass Aabb::HasOverlap. Independent of
// any tree topology; result sorted lexicographically.
inline std::vector<std::pair<int, int>> LeafCandidatesBruteForce(const std::vector<Aabb>& boxes_A,
                                                                 const std::vector<Aabb>& boxes_B, const Xform& X_AB) {
  std::vector<std::pair<int, int>> out;
  for (int i = 0; i < static_cast<int>(boxes_A.size()); ++i)
    for (int j = 0; j < static_cast<int>(boxes_B.size()); ++j)
      if (AabbHasOverlap(boxes_A[i], boxes_B[j], X_AB)) out.emplace_back(i, j);
  return out;
}

// Same set through a tree with one element per leaf. Internal-node tests use
// half widths inflated by `kSlack` so an ancestor can never reject a pair
// whose leaves pass (the leaf test itself is the exact reference arithmetic).
inline std::vector<std::pair<int, int>> LeafCandidatesTree(const Bvh& A, const Bvh& B, const Xform& X_AB) {
  std::vector<std::pair<int, int>> out;
  std::vector<std::pair<int, int>> stack;
  stack.emplace_back(0, 0);
  while (!stack.empty()) {
    const auto [ia, ib] = stack.back();
    stack.pop_back();
    const BvhNode& a = A.nodes[ia];
    const BvhNode& b = B.nodes[ib];
    const bool both_leaves = a.is_leaf() && b.is_leaf();
    if (both_leaves) {
      if (AabbHasOverlap(a.bv, b.bv, X_AB)) out.emplace_back(a.indices[0], b.indices[0]);
      continue;
    }
    const Vec3 t = (X_AB * b.bv.center) - a.bv.center;
    const double s = 1e-12 * (std::abs(t.x) + std::abs(t.y) + std::abs(t.z) + a.bv.half_width.x + a.bv.half_width.y +
                              a.bv.half_width.z + b.bv.half_width.x + b.bv.half_width.y + b.bv.half_width.z) +
                     1e-300;
    const Vec3 ha = a.bv.half_width + Vec3{s, s, s};
    const Vec3 hb = b.bv.half_width + Vec3{s, s, s};
    if (!BoxesOverlap(ha, hb, X_AB.R, t)) continue;
    if (b.is_leaf()) {
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
  std::sort(out.begin(), out.end());
  return out;
}

}  // namespace ho
