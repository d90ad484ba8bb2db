// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// Mesh / field data types and the synthetic generators, restated from:
//   geometry/proximity/volume_mesh.{h,cc}, triangle_surface_mesh.h,
//   mesh_field_linear.h, make_sphere_mesh.h, make_ellipsoid_mesh.h,
//   make_box_mesh.cc, make_sphere_field.h, make_ellipsoid_field.h,
//   make_box_field.cc, distance_to_point_callback.cc,
//   volume_to_surface_mesh.cc.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

#include "ho_math.h"

namespace ho {

using Tet = std::array<int, 4>;
using Tri = std::array<int, 3>;

// VolumeMesh<double> (volume_mesh.h:94-359).
struct VolumeMesh {
  std::vector<Vec3> verts;
  std::vector<Tet> tets;
  // inward_normals[e][f]: unit normal of the face opposite vertex f, pointing
  // into the tet (volume_mesh.cc:91-138).
  std::vector<std::array<Vec3, 4>> inward_normals;

  VolumeMesh() = default;
  VolumeMesh(std::vector<Tet> t, std::vector<Vec3> v) : verts(std::move(v)), tets(std::move(t)) {
    ComputePositionDependentQuantities();
  }
  int num_elements() const { return static_cast<int>(tets.size()); }
  int num_vertices() const { return static_cast<int>(verts.size()); }

  void ComputePositionDependentQuantities() {
    inward_normals.clear();
    inward_normals.reserve(tets.size());
    for (const Tet& t : tets) {
      const Vec3& a = verts[t[0]];
      const Vec3& b = verts[t[1]];
      const Vec3& c = verts[t[2]];
      const Vec3& d = verts[t[3]];
      const Vec3 ab = b - a, ac = c - a, ad = d - a, bc = c - b, bd = d - b;
      inward_normals.push_back({normalized(cross(bd, bc)), normalized(cross(ac, ad)),
                                normalized(cross(ad, ab)), normalized(cross(ab, ac))});
    }
  }

  // volume_mesh.h:361-448. Returns false when degenerate.
  bool MaybeCalcGradBarycentric(int e, int i, Vec3* out) const {
    const Vec3& V = verts[tets[e][i]];
    const Vec3& A = verts[tets[e][(i + 1) % 4]];
    const Vec3& B = verts[tets[e][(i + 2) % 4]];
    const Vec3& C = verts[tets[e][(i + 3) % 4]];
    const Vec3 AV = V - A, AB = B - A, AC = C - A;
    const Vec3 area_vector = cross(AB, AC);
    const double signed_volume = dot(area_vector, AV);
    constexpr double kEps = std::numeric_limits<double>::epsilon();
    if (std::abs(signed_volume) <= kEps) return false;
    *out = area_vector / signed_volume;
    return true;
  }

  double CalcTetVolume(int e) const {
    const Vec3& a = verts[tets[e][0]];
    const Vec3 ab = verts[tets[e][1]] - a, ac = verts[tets[e][2]] - a, ad = verts[tets[e][3]] - a;
    return dot(cross(ab, ac), ad) / 6.0;
  }
};

// VolumeMeshFieldLinear<double,double> = MeshFieldLinear (mesh_field_linear.h).
struct VolumeField {
  std::vector<double> values;      // per vertex
  std::vector<Vec3> gradients;     // per tet
  std::vector<double> values_at_Mo;  // per tet
  bool degenerate{false};

  VolumeField() = default;
  // mesh_field_linear.h:185-203 with MeshGradientMode::kOkOrMarkDegenerate.
  VolumeField(std::vector<double> v, const VolumeMesh& mesh) : values(std::move(v)) {
    gradients.reserve(mesh.tets.size());
    for (int e = 0; e < mesh.num_elements(); ++e) {
      Vec3 g{0, 0, 0};
      for (int i = 0; i < 4; ++i) {
        Vec3 gi;
        if (!mesh.MaybeCalcGradBarycentric(e, i, &gi)) {
          degenerate = true;
          gradients.clear();
          return;
        }
        g += values[mesh.tets[e][i]] * gi;  // volume_mesh.h:275-289
      }
      gradients.push_back(g);
    }
    values_at_Mo.reserve(mesh.tets.size());
    for (int e = 0; e < mesh.num_elements(); ++e) {
      const int v0 = mesh.tets[e][0];
      values_at_Mo.push_back(values[v0] - dot(gradients[e], mesh.verts[v0]));  // :502-509
    }
  }
  // mesh_field_linear.h:334-351.
  double EvaluateCartesian(int e, const Vec3& p) const { return dot(gradients[e], p) + values_at_Mo[e]; }
};

// TriangleSurfaceMesh<double> (triangle_surface_mesh.h:452-487).
struct TriMesh {
  std::vector<Vec3> verts;
  std::vector<Tri> tris;
  std::vector<double> areas;
  std::vector<Vec3> normals;
  double total_area{0};
  Vec3 centroid;

  TriMesh() = default;
  TriMesh(std::vector<Tri> t, std::vector<Vec3> v) : verts(std::move(v)), tris(std::move(t)) {
    ComputePositionDependentQuantities();
  }
  int num_elements() const { return static_cast<int>(tris.size()); }
  void ComputePositionDependentQuantities() {
    areas.assign(tris.size(), 0.0);
    normals.assign(tris.size(), Vec3{});
    total_area = 0;
    centroid = Vec3{};
    for (size_t f = 0; f < tris.size(); ++f) {
      const Vec3& A = verts[tris[f][0]];
      const Vec3& B = verts[tris[f][1]];
      const Vec3& C = verts[tris[f][2]];
      const Vec3 c = cross(B - A, C - A);
      const double n = norm(c);
      const double face_area = 0.5 * n;
      areas[f] = face_area;
      total_area += face_area;
      normals[f] = (n != 0.0) ? c / n : c;
      centroid += face_area * (A + B + C);
    }
    if (total_area != 0.0) centroid = centroid / (3. * total_area);
  }
};

enum class Tessellation { kSingleInteriorVertex = 0, kDenseInteriorVertices = 1 };

namespace detail {

inline uint64_t EdgeKey(int a, int b) {
  if (a > b) std::swap(a, b);
  return (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b);
}

// make_sphere_mesh.h:101-207.
inline void SplitOctohedron(const std::array<int, 6>& vi, const std::vector<Vec3>& p, std::vector<Tet>* tets) {
  auto sq_len = [&](int a, int b) { return squared_norm(p[a] - p[b]); };
  auto volume = [&](int a, int b, int c, int d) {
    const Vec3 AB = p[b] - p[a], AC = p[c] - p[a], AD = p[d] - p[a];
    return dot(cross(AB, AC), AD) / 6.0;
  };
  auto tet_quality = [&](const std::array<int, 4>& l) {
    const int a = vi[l[0]], b = vi[l[1]], c = vi[l[2]], d = vi[l[3]];
    const double v = volume(a, b, c, d);
    double denom = 0.0;
    denom += sq_len(a, b);
    denom += sq_len(a, c);
    denom += sq_len(a, d);
    denom += sq_len(b, c);
    denom += sq_len(b, d);
    denom += sq_len(c, d);
    return v * v / (denom * denom * denom);
  };
  using Split = std::array<std::array<int, 4>, 4>;
  auto split_quality = [&](const Split& s) {
    double cost = 0;
    for (const auto& t : s) cost += tet_quality(t);
    return cost;
  };
  const int e = 0, f = 1, g = 2, h = 3, i = 4, j = 5;
  const Split kGh{{{g, h, i, e}, {g, f, h, e}, {g, i, h, j}, {g, h, f, j}}};
  const Split kFi{{{f, i, e, h}, {f, i, h, j}, {f, i, j, g}, {f, e, i, g}}};
  const Split kEj{{{e, h, j, i}, {e, i, j, g}, {e, g, j, f}, {e, f, j, h}}};
  double max_quality = -1.0;
  const Split* best = nullptr;
  for (const Split* cand : {&kGh, &kFi, &kEj}) {
    const double q = split_quality(*cand);
    if (q > max_quality) {
      max_quality = q;
      best = cand;
    }
  }
  if (!(max_quality > 0.0)) throw std::logic_error("SplitOctohedron: non-positive quality");
  for (int t = 0; t < 4; ++t)
    tets->push_back({vi[(*best)[t][0]], vi[(*best)[t][1]], vi[(*best)[t][2]], vi[(*best)[t][3]]});
}

struct SphereLevel {
  std::vector<Vec3> verts;
  std::vector<Tet> tets;
  std::vector<bool> is_boundary;
};

// make_sphere_mesh.h:30-81.
inline SphereLevel MakeSphereMeshLevel0() {
  SphereLevel m;
  m.verts = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  m.tets = {{1, 5, 2, 0}, {2, 5, 3, 0}, {3, 5, 4, 0}, {4, 5, 1, 0},
            {2, 6, 1, 0}, {3, 6, 2, 0}, {4, 6, 3, 0}, {1, 6, 4, 0}};
  m.is_boundary.assign(7, true);
  m.is_boundary[0] = false;
  return m;
}

// make_sphere_mesh.h:294-341 (dense refinement, x8).
inline SphereLevel RefineUnitSphereMesh(const SphereLevel& in) {
  SphereLevel out;
  out.verts = in.verts;
  out.is_boundary = in.is_boundary;
  std::unordered_map<uint64_t, int> edge_vertex;
  static const int kEdges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  for (const Tet& t : in.tets) {
    for (const auto& e : kEdges) {
      int a = t[e[0]], b = t[e[1]];
      if (a > b) std::swap(a, b);
      const uint64_t key = EdgeKey(a, b);
      if (edge_vertex.find(key) == edge_vertex.end()) {
        Vec3 v = (in.verts[a] + in.verts[b]) / 2.0;
        const bool bnd = in.is_boundary[a] && in.is_boundary[b];
        if (bnd) v = v / norm(v);  // Eigen normalize(): v /= norm()
        edge_vertex.emplace(key, static_cast<int>(out.verts.size()));
        out.is_boundary.push_back(bnd);
        out.verts.push_back(v);
      }
    }
  }
  out.tets.reserve(in.tets.size() * 8);
  for (const Tet& t : in.tets) {
    const int a = t[0], b = t[1], c = t[2], d = t[3];
    const int e = edge_vertex.at(EdgeKey(a, b));
    const int f = edge_vertex.at(EdgeKey(a, c));
    const int g = edge_vertex.at(EdgeKey(a, d));
    const int h = edge_vertex.at(EdgeKey(b, c));
    const int i = edge_vertex.at(EdgeKey(b, d));
    const int j = edge_vertex.at(EdgeKey(c, d));
    out.tets.push_back({a, e, f, g});
    out.tets.push_back({b, h, e, i});
    out.tets.push_back({f, h, c, j});
    out.tets.push_back({j, g, i, d});
    SplitOctohedron({e, f, g, h, i, j}, out.verts, &out.tets);
  }
  return out;
}

// make_sphere_mesh.h:370-445 (sparse refinement, x4; centre vertex is index 0).
inline SphereLevel RefineUnitSphereMeshOnSurface(const SphereLevel& in, int center_index) {
  SphereLevel out;
  out.verts = in.verts;
  std::unordered_map<uint64_t, int> edge_vertex;
  static const int kEdges[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  for (const Tet& t : in.tets) {
    if (t[3] != center_index) throw std::logic_error("sparse sphere: centre must be 4th vertex");
    for (const auto& e : kEdges) {
      int a = t[e[0]], b = t[e[1]];
      if (a > b) std::swap(a, b);
      const uint64_t key = EdgeKey(a, b);
      if (edge_vertex.find(key) == edge_vertex.end()) {
        Vec3 v = (in.verts[a] + in.verts[b]) / 2.0;
        v = v / norm(v);
        edge_vertex.emplace(key, static_cast<int>(out.verts.size()));
        out.verts.push_back(v);
      }
    }
  }
  out.tets.reserve(in.tets.size() * 4);
  const int v0 = center_index;
  for (const Tet& t : in.tets) {
    const int a = t[0], b = t[1], c = t[2];
    const int d = edge_vertex.at(EdgeKey(b, c));
    const int e = edge_vertex.at(EdgeKey(a, b));
    const int f = edge_vertex.at(EdgeKey(a, c));
    out.tets.push_back({a, e, f, v0});
    out.tets.push_back({b, d, e, v0});
    out.tets.push_back({c, f, d, v0});
    out.tets.push_back({e, d, f, v0});
  }
  return out;
}

}  // namespace detail

// make_sphere_mesh.h:470-507.
inline VolumeMesh MakeUnitSphereMesh(int refinement_level, Tessellation strategy) {
  if (refinement_level < 0) throw std::logic_error("negative refinement level");
  detail::SphereLevel m = detail::MakeSphereMeshLevel0();
  if (strategy == Tessellation::kSingleInteriorVertex) {
    for (int l = 1; l <= refinement_level; ++l) m = detail::RefineUnitSphereMeshOnSurface(m, 0);
  } else {
    for (int l = 1; l <= refinement_level; ++l) m = detail::RefineUnitSphereMesh(m);
  }
  return VolumeMesh(std::move(m.tets), std::move(m.verts));
}

// make_sphere_mesh.h:576-590.
inline int SphereRefinementLevel(double radius, double resolution_hint) {
  const double edge_length = std::min(resolution_hint, 2.0 * radius);
  const int L = std::max(
      0, static_cast<int>(std::ceil(std::log2(M_PI / std::asin(edge_length / (2.0 * radius)))) - 2));
  return std::min(L, 8);
}

// make_sphere_mesh.h:543-598.
inline VolumeMesh MakeSphereVolumeMesh(double radius, double resolution_hint, Tessellation strategy) {
  if (!(resolution_hint > 0.0)) throw std::logic_error("resolution_hint must be positive");
  VolumeMesh unit = MakeUnitSphereMesh(SphereRefinementLevel(radius, resolution_hint), strategy);
  std::vector<Vec3> verts;
  verts.reserve(unit.verts.size());
  for (const Vec3& v : unit.verts) verts.push_back(v * radius);
  return VolumeMesh(std::move(unit.tets), std::move(verts));
}

// make_ellipsoid_mesh.h:50-73.
inline VolumeMesh MakeEllipsoidVolumeMesh(double a, double b, double c, double resolution_hint,
                                          Tessellation strategy) {
  if (!(resolution_hint > 0.0)) throw std::logic_error("resolution_hint must be positive");
  const double r = std::max({a, b, c});
  VolumeMesh unit = MakeSphereVolumeMesh(1.0, resolution_hint / r, strategy);
  std::vector<Vec3> verts;
  verts.reserve(unit.verts.size());
  for (const Vec3& v : unit.verts) verts.push_back({a * v.x, b * v.y, c * v.z});
  return VolumeMesh(std::move(unit.tets), std::move(verts));
}

// volume_to_surface_mesh.cc:31-187. tri_to_tet (optional) receives the tet
// index each boundary triangle came from.
inline TriMesh ConvertVolumeToSurfaceMesh(const VolumeMesh& volume, std::vector<int>* tri_to_tet = nullptr) {
  struct BoundaryFace {
    Tri vertices;
    int tet;
  };
  std::map<std::array<int, 3>, BoundaryFace> face_map;  // key = sorted triplet
  static const int kFaces[4][3] = {{1, 2, 3}, {3, 2, 0}, {1, 3, 0}, {2, 1, 0}};
  for (int ti = 0; ti < volume.num_elements(); ++ti) {
    const Tet& t = volume.tets[ti];
    for (const auto& f : kFaces) {
      const int v0 = t[f[0]], v1 = t[f[1]], v2 = t[f[2]];
      std::array<int, 3> key{v0, v1, v2};
      std::sort(key.begin(), key.end());
      auto it = face_map.find(key);
      if (it != face_map.end()) {
        face_map.erase(it);
      } else {
        face_map.emplace(key, BoundaryFace{{v0, v1, v2}, ti});
      }
    }
  }
  std::set<int> vertex_set;
  for (const auto& kv : face_map)
    for (int v : kv.second.vertices) vertex_set.insert(v);
  std::vector<int> boundary_vertices(vertex_set.begin(), vertex_set.end());
  std::unordered_map<int, int> volume_to_surface;
  std::vector<Vec3> surface_vertices;
  surface_vertices.reserve(boundary_vertices.size());
  for (size_t i = 0; i < boundary_vertices.size(); ++i) {
    surface_vertices.push_back(volume.verts[boundary_vertices[i]]);
    volume_to_surface.emplace(boundary_vertices[i], static_cast<int>(i));
  }
  std::vector<Tri> faces;
  faces.reserve(face_map.size());
  if (tri_to_tet) tri_to_tet->clear();
  for (const auto& kv : face_map) {
    const Tri& f = kv.second.vertices;
    faces.push_back({volume_to_surface.at(f[0]), volume_to_surface.at(f[1]), volume_to_surface.at(f[2])});
    if (tri_to_tet) tri_to_tet->push_back(kv.second.tet);
  }
  return TriMesh(std::move(faces), std::move(surface_vertices));
}

// make_sphere_mesh.h:614-620.
inline TriMesh MakeSphereSurfaceMesh(double radius, double resolution_hint) {
  return ConvertVolumeToSurfaceMesh(
      MakeSphereVolumeMesh(radius, resolution_hint, Tessellation::kSingleInteriorVertex));
}
// make_ellipsoid_mesh.h:90-95.
inline TriMesh MakeEllipsoidSurfaceMesh(double a, double b, double c, double resolution_hint) {
  return ConvertVolumeToSurfaceMesh(
      MakeEllipsoidVolumeMesh(a, b, c, resolution_hint, Tessellation::kSingleInteriorVertex));
}

// proximity_utilities.h:29-31.
inline double DistanceToPointRelativeTolerance(double size) { return 1e-14 * std::max(1.0, size); }

namespace detail {
// make_box_mesh.cc:69-134.
inline VolumeMesh MakeBoxWithPointMa(double length) {
  std::vector<Tet> e = {{8, 0, 4, 11}, {3, 8, 1, 9},  {0, 8, 4, 13}, {8, 7, 3, 14}, {8, 5, 4, 10}, {8, 0, 1, 9},
                        {5, 8, 1, 14}, {0, 8, 2, 9},  {2, 8, 3, 9},  {8, 4, 6, 10}, {7, 8, 6, 10}, {8, 5, 1, 11},
                        {4, 8, 6, 13}, {8, 5, 7, 14}, {8, 7, 6, 12}, {8, 6, 2, 12}, {5, 8, 7, 10}, {5, 8, 4, 11},
                        {8, 3, 1, 14}, {8, 2, 3, 12}, {7, 8, 3, 12}, {8, 0, 2, 13}, {6, 8, 2, 13}, {0, 8, 1, 11}};
  const double L = length / 2.0;
  std::vector<Vec3> v = {{-L, -L, -L}, {-L, -L, L}, {-L, L, -L}, {-L, L, L},  {L, -L, -L},
                         {L, -L, L},   {L, L, -L},  {L, L, L},   {0., 0., 0.}, {-L, 0., 0.},
                         {L, 0., 0.},  {0., -L, 0.}, {0., L, 0.}, {0., 0., -L}, {0., 0., L}};
  return VolumeMesh(std::move(e), std::move(v));
}

// make_box_mesh.cc:141-296.
inline VolumeMesh MakeBoxWithLineMa(const Vec3& size) {
  std::vector<Tet> e = {
      {5, 10, 4, 13}, {3, 1, 8, 16},  {7, 8, 5, 16},  {6, 9, 4, 12},  {8, 1, 5, 16},  {2, 10, 3, 11}, {6, 7, 10, 12},
      {10, 5, 4, 12}, {10, 0, 1, 11}, {7, 6, 10, 14}, {8, 7, 5, 12},  {3, 8, 7, 16},  {6, 9, 2, 15},  {9, 0, 2, 15},
      {1, 8, 10, 11}, {1, 3, 8, 11},  {8, 3, 10, 11}, {2, 9, 10, 11}, {0, 9, 2, 11},  {10, 9, 0, 11}, {8, 3, 7, 14},
      {8, 7, 10, 14}, {9, 6, 2, 14},  {10, 6, 9, 14}, {10, 2, 3, 14}, {3, 8, 10, 14}, {9, 2, 10, 14}, {9, 4, 0, 15},
      {9, 6, 4, 15},  {10, 9, 4, 13}, {4, 9, 0, 13},  {8, 10, 5, 13}, {1, 8, 5, 13},  {0, 10, 1, 13}, {9, 10, 0, 13},
      {8, 1, 10, 13}, {10, 8, 5, 12}, {7, 8, 10, 12}, {9, 10, 4, 12}, {6, 10, 9, 12}};
  const Vec3 h = size / 2.0;
  int z_dir = 0;
  double L2 = h[0];
  for (int i = 1; i < 3; ++i)
    if (h[i] > L2) {  // Eigen maxCoeff: first maximal index
      L2 = h[i];
      z_dir = i;
    }
  const int x_dir = (z_dir + 1) % 3;
  const double L1 = h[x_dir];
  const double Lm = L2 - L1;
  std::vector<Vec3> v = {{-L1, -L1, -L2}, {-L1, -L1, L2}, {-L1, L1, -L2}, {-L1, L1, L2}, {L1, -L1, -L2}, {L1, -L1, L2},
                         {L1, L1, -L2},   {L1, L1, L2},   {0, 0, Lm},     {0, 0, -Lm},   {0, 0, 0},      {-L1, 0, 0},
                         {L1, 0, 0},      {0, -L1, 0},    {0, L1, 0},     {0, 0, -L2},   {0, 0, L2}};
  if (z_dir == 0)
    for (Vec3& p : v) p = Vec3(-p.z, p.y, p.x);
  if (z_dir == 1)
    for (Vec3& p : v) p = Vec3(p.x, -p.z, p.y);
  return VolumeMesh(std::move(e), std::move(v));
}

// make_box_mesh.cc:303-471.
inline VolumeMesh MakeBoxWithRectangleMa(const Vec3& size) {
  std::vector<Tet> e = {
      {4, 11, 6, 16},  {8, 5, 1, 14},   {5, 9, 8, 17},   {5, 4, 8, 13},   {5, 8, 1, 17},   {0, 11, 2, 12},
      {7, 9, 6, 13},   {3, 7, 9, 15},   {3, 2, 9, 12},   {9, 7, 6, 15},   {1, 9, 3, 17},   {6, 11, 2, 16},
      {2, 11, 9, 12},  {9, 8, 1, 12},   {9, 5, 8, 13},   {5, 9, 7, 13},   {8, 9, 11, 12},  {8, 0, 1, 12},
      {11, 4, 6, 13},  {9, 1, 3, 12},   {4, 5, 8, 14},   {7, 3, 9, 17},   {9, 5, 7, 17},   {8, 11, 10, 12},
      {0, 8, 10, 12},  {11, 0, 10, 12}, {8, 9, 1, 17},   {11, 6, 2, 15},  {11, 9, 6, 15},  {2, 3, 9, 15},
      {11, 2, 9, 15},  {4, 0, 10, 16},  {11, 4, 10, 16}, {11, 0, 2, 16},  {0, 11, 10, 16}, {4, 8, 10, 14},
      {0, 4, 10, 14},  {0, 8, 1, 14},   {8, 0, 10, 14},  {9, 11, 6, 13},  {9, 8, 11, 13},  {8, 4, 10, 13},
      {11, 8, 10, 13}, {4, 11, 10, 13}};
  const Vec3 h = size / 2.0;
  int x_dir = 0;
  double L1 = h[0];
  for (int i = 1; i < 3; ++i)
    if (h[i] < L1) {  // Eigen minCoeff: first minimal index
      L1 = h[i];
      x_dir = i;
    }
  double L2 = h[1], L3 = h[2];
  if (x_dir == 2) {
    L2 = h[1];
    L3 = h[0];
  }
  if (x_dir == 1) {
    L3 = h[2];
    L2 = h[0];
  }
  const double Lm2 = L2 - L1, Lm3 = L3 - L1;
  std::vector<Vec3> v = {{-L1, -L2, -L3}, {-L1, -L2, L3}, {-L1, L2, -L3}, {-L1, L2, L3}, {L1, -L2, -L3}, {L1, -L2, L3},
                         {L1, L2, -L3},   {L1, L2, L3},   {0, -Lm2, Lm3}, {0, Lm2, Lm3}, {0, -Lm2, -Lm3}, {0, Lm2, -Lm3},
                         {-L1, 0, 0},     {L1, 0, 0},     {0, -L2, 0},    {0, L2, 0},    {0, 0, -L3},    {0, 0, L3}};
  if (x_dir == 2)
    for (Vec3& p : v) p = Vec3(-p.z, p.y, p.x);
  if (x_dir == 1)
    for (Vec3& p : v) p = Vec3(-p.y, p.x, p.z);
  return VolumeMesh(std::move(e), std::move(v));
}
}  // namespace detail

// make_box_mesh.cc:475-504.
inline VolumeMesh MakeBoxVolumeMeshWithMaAndSymmetricTriangles(const Vec3& size) {
  const Vec3 half_box = size / 2.;
  const double min_half = std::min({half_box.x, half_box.y, half_box.z});
  const double tol = DistanceToPointRelativeTolerance(min_half);
  Vec3 ma;
  for (int i = 0; i < 3; ++i) {
    const double before = half_box[i] - min_half;
    ma[i] = (before > tol) ? before : 0.;
  }
  const int zeros = (ma.x == 0) + (ma.y == 0) + (ma.z == 0);
  if (zeros == 3) return detail::MakeBoxWithPointMa(size.x);
  if (zeros == 2) return detail::MakeBoxWithLineMa(size);
  if (zeros != 1) throw std::logic_error("box MA classification failed");
  return detail::MakeBoxWithRectangleMa(size);
}

inline TriMesh MakeBoxSurfaceMeshWithSymmetricTriangles(const Vec3& size) {
  return ConvertVolumeToSurfaceMesh(MakeBoxVolumeMeshWithMaAndSymmetricTriangles(size));
}

// make_box_mesh.cc:768-890 (regular grid, 6 tets per cell).
inline VolumeMesh MakeBoxVolumeMesh(const Vec3& size, double resolution_hint) {
  if (!(resolution_hint > 0.)) throw std::logic_error("resolution_hint must be positive");
  const int nv[3] = {1 + static_cast<int>(std::ceil(size.x / resolution_hint)),
                     1 + static_cast<int>(std::ceil(size.y / resolution_hint)),
                     1 + static_cast<int>(std::ceil(size.z / resolution_hint))};
  // Eigen::LinSpaced(n, low, high) for doubles: low + i*step, with the last
  // entry pinned to high (step = (high-low)/(n-1)); n==1 yields high.
  auto linspaced = [](int n, double low, double high) {
    std::vector<double> c(n);
    if (n == 1) {
      c[0] = high;
      return c;
    }
    const double step = (high - low) / (n - 1);
    for (int i = 0; i < n; ++i) c[i] = low + i * step;
    c[n - 1] = high;
    return c;
  };
  const auto xs = linspaced(nv[0], -size.x / 2, size.x / 2);
  const auto ys = linspaced(nv[1], -size.y / 2, size.y / 2);
  const auto zs = linspaced(nv[2], -size.z / 2, size.z / 2);
  std::vector<Vec3> verts;
  verts.reserve(static_cast<size_t>(nv[0]) * nv[1] * nv[2]);
  for (int i = 0; i < nv[0]; ++i)
    for (int j = 0; j < nv[1]; ++j)
      for (int k = 0; k < nv[2]; ++k) verts.push_back({xs[i], ys[j], zs[k]});
  auto seq = [&](int i, int j, int k) { return i * nv[1] * nv[2] + j * nv[2] + k; };
  std::vector<Tet> tets;
  tets.reserve(static_cast<size_t>(6) * (nv[0] - 1) * (nv[1] - 1) * (nv[2] - 1));
  for (int i = 0; i < nv[0] - 1; ++i)
    for (int j = 0; j < nv[1] - 1; ++j)
      for (int k = 0; k < nv[2] - 1; ++k) {
        int v[8];
        int s = 0;
        for (int l = 0; l < 2; ++l)
          for (int m = 0; m < 2; ++m)
            for (int n = 0; n < 2; ++n) v[s++] = seq(i + l, j + m, k + n);
        tets.push_back({v[0], v[7], v[4], v[6]});
        tets.push_back({v[0], v[7], v[6], v[2]});
        tets.push_back({v[0], v[7], v[2], v[3]});
        tets.push_back({v[0], v[7], v[3], v[1]});
        tets.push_back({v[0], v[7], v[1], v[5]});
        tets.push_back({v[0], v[7], v[5], v[4]});
      }
  return VolumeMesh(std::move(tets), std::move(verts));
}

// make_sphere_field.h:38-73.
inline VolumeField MakeSpherePressureField(double radius, const VolumeMesh& mesh, double modulus, double margin = 0.0) {
  std::vector<double> p;
  p.reserve(mesh.verts.size());
  const double kExtentEpsilon = 1e-14;
  const double min_extent = -margin / (radius - margin);
  for (const Vec3& r : mesh.verts) {
    const double distance = radius - norm(r);
    double extent = (distance - margin) / (radius - margin);
    if (std::abs(distance / radius) < kExtentEpsilon) extent = min_extent;
    p.push_back(modulus * extent);
  }
  return VolumeField(std::move(p), mesh);
}

// make_ellipsoid_field.h:41-92.
inline VolumeField MakeEllipsoidPressureField(double a, double b, double c, const VolumeMesh& mesh, double modulus,
                                              double margin = 0.0) {
  const Vec3 scale{1.0 / (a - margin), 1.0 / (b - margin), 1.0 / (c - margin)};
  std::vector<double> p;
  p.reserve(mesh.verts.size());
  constexpr double kExtentEpsilon = 1e-14;
  for (const Vec3& r : mesh.verts) {
    const Vec3 u{scale.x * r.x, scale.y * r.y, scale.z * r.z};
    double extent = 1.0 - norm(u);
    if (margin == 0 && std::abs(extent) < kExtentEpsilon) extent = 0.0;
    p.push_back(modulus * extent);
  }
  return VolumeField(std::move(p), mesh);
}

namespace detail {
inline double Sign(double x) { return (x < 0.0) ? -1.0 : 1.0; }
// distance_to_point_callback.cc:405-494 for dim = 3: returns (nearest point N, gradient).
inline std::pair<Vec3, Vec3> ComputeDistanceToBox(const Vec3& h, const Vec3& q) {
  enum class Loc { kInside, kBoundary, kOutside };
  Vec3 c;
  Loc loc[3];
  for (int i = 0; i < 3; ++i) {
    const double tol = DistanceToPointRelativeTolerance(h[i]);
    if (std::abs(q[i]) > h[i] + tol) {
      loc[i] = Loc::kOutside;
      c[i] = Sign(q[i]) * h[i];
    } else if (std::abs(q[i]) >= h[i] - tol) {
      loc[i] = Loc::kBoundary;
      c[i] = Sign(q[i]) * h[i];
    } else {
      loc[i] = Loc::kInside;
      c[i] = q[i];
    }
  }
  Vec3 n = c;
  Vec3 grad{0, 0, 0};
  const bool any_out = loc[0] == Loc::kOutside || loc[1] == Loc::kOutside || loc[2] == Loc::kOutside;
  const bool any_bnd = loc[0] == Loc::kBoundary || loc[1] == Loc::kBoundary || loc[2] == Loc::kBoundary;
  if (any_out) {
    const Vec3 nq = q - n;
    grad = nq / norm(nq);
  } else if (any_bnd) {
    for (int i = 0; i < 3; ++i)
      if (loc[i] == Loc::kBoundary) grad[i] = Sign(c[i]);
    grad = grad / norm(grad);
  } else {
    double min_dist = std::numeric_limits<double>::infinity();
    int axis = -1;
    for (int i = 0; i < 3; ++i)
      for (const double bound : {h[i], -h[i]}) {
        const double dist = std::abs(bound - q[i]);
        if (dist < min_dist) {
          min_dist = dist;
          axis = i;
        }
      }
    const double sign = Sign(q[axis]);
    n[axis] = sign * h[axis];
    grad[axis] = sign;
  }
  return {n, grad};
}
}  // namespace detail

// make_box_field.cc:15-56.
inline VolumeField MakeBoxPressureField(const Vec3& size, const VolumeMesh& mesh, double modulus, double margin = 0.0) {
  const Vec3 half = size / 2.0;
  const double min_half = std::min({half.x, half.y, half.z});
  std::vector<double> p;
  p.reserve(mesh.verts.size());
  for (const Vec3& r : mesh.verts) {
    const auto [n, grad] = detail::ComputeDistanceToBox(half, r);
    const double signed_distance = dot(grad, r - n);
    const double extent = (-signed_distance - margin) / (min_half - margin);
    p.push_back(modulus * extent);
  }
  return VolumeField(std::move(p), mesh);
}

}  // namespace ho
