// Oracle (test infrastructure): the reference's hydroelastic representations of Cylinder and Capsule restated --
//   meshing_utilities.cc:21-46            SplitTriangularPrismToTetrahedra, SplitPyramidToTetrahedra
//   make_cylinder_mesh.cc:47-425          MakeCylinderVolumeMeshWithMa (long / medium / short medial axis)
//   make_capsule_mesh.cc:13-180           MakeCapsuleVolumeMesh
//   make_cylinder_field.cc:18-75          MakeCylinderPressureField (over DistanceToPoint(Cylinder),
//                                         distance_to_point_callback.cc:194-288,406-482)
//   make_capsule_field.h:55-90            MakeCapsulePressureField
// Pinned by the reference's own tests (make_cylinder_mesh_test.cc, make_capsule_mesh_test.cc, make_cylinder_field_test.cc,
// make_capsule_field_test.cc: vertex / tetrahedron counts per class, positive volumes, boundary and medial pressures),
// restated in tests/test_oracle_cylinder_capsule.py. Never linked into the product.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "ho_mesh.h"

namespace ho {
namespace detail {

// meshing_utilities.cc:21-33.
inline std::vector<Tet> SplitTriangularPrismToTetrahedra(int v0, int v1, int v2, int v3, int v4, int v5) {
  std::vector<Tet> elements;
  int previous = v3;
  for (int next : {v4, v1, v2}) {
    elements.push_back({previous, next, v0, v5});
    previous = next;
  }
  return elements;
}
// meshing_utilities.cc:35-46.
inline std::vector<Tet> SplitPyramidToTetrahedra(int v0, int v1, int v2, int v3, int v4) {
  std::vector<Tet> elements;
  int previous = v3;
  for (int next : {v4, v1}) {
    elements.push_back({previous, next, v0, v2});
    previous = next;
  }
  return elements;
}
inline void Append(const std::vector<Tet>& more, std::vector<Tet>* all) { all->insert(all->end(), more.begin(), more.end()); }

}  // namespace detail

// make_cylinder_mesh.cc:325-425 with the three Calc*CylinderVolumeMeshWithMa helpers (:47-260).
inline VolumeMesh MakeCylinderVolumeMeshWithMa(double radius, double length, double resolution_hint) {
  if (!(resolution_hint > 0.0) || !(radius > 0.0) || !(length > 0.0)) throw std::invalid_argument("cylinder mesh");
  using detail::Append;
  const double top_z = length / 2.;
  const double bottom_z = -top_z;
  const double tolerance = DistanceToPointRelativeTolerance(std::min(top_z, radius));
  enum class Class { kLong, kMedium, kShort } cls = Class::kMedium;
  if (top_z - radius > tolerance) {
    cls = Class::kLong;
  } else if (radius - top_z > tolerance) {
    cls = Class::kShort;
  }
  const int n_rim = std::max(3, static_cast<int>(std::ceil(2. * M_PI * radius / resolution_hint)));
  std::vector<Vec3> V;
  const int c_lo = 0;
  V.emplace_back(0, 0, bottom_z);
  const int c_hi = 1;
  V.emplace_back(0, 0, top_z);
  std::vector<int> bottom(n_rim), top(n_rim);
  const double dphi = 2. * M_PI / n_rim;
  for (int i = 0; i < n_rim; ++i) {
    const double x = radius * std::cos(dphi * i);
    const double y = radius * std::sin(dphi * i);
    bottom[i] = static_cast<int>(V.size());
    V.emplace_back(x, y, bottom_z);
    top[i] = static_cast<int>(V.size());
    V.emplace_back(x, y, top_z);
  }
  std::vector<Tet> E;
  if (cls == Class::kLong) {
    const double z_off = top_z - radius;
    int medial[2];
    medial[0] = static_cast<int>(V.size());
    V.emplace_back(0, 0, -z_off);
    medial[1] = static_cast<int>(V.size());
    V.emplace_back(0, 0, z_off);
    int i = n_rim - 1;
    for (int j = 0; j < n_rim; ++j) {
      E.push_back({c_lo, bottom[i], bottom[j], medial[0]});
      E.push_back({c_hi, top[j], top[i], medial[1]});
      Append(detail::SplitTriangularPrismToTetrahedra(medial[0], bottom[i], bottom[j], medial[1], top[i], top[j]), &E);
      i = j;
    }
  } else if (cls == Class::kMedium) {
    const int medial = static_cast<int>(V.size());
    V.emplace_back(0, 0, 0);
    int i = n_rim - 1;
    for (int j = 0; j < n_rim; ++j) {
      E.push_back({c_lo, bottom[i], bottom[j], medial});
      E.push_back({c_hi, top[j], top[i], medial});
      Append(detail::SplitPyramidToTetrahedra(top[i], top[j], bottom[j], bottom[i], medial), &E);
      i = j;
    }
  } else {
    const int center = static_cast<int>(V.size());
    V.emplace_back(0, 0, 0);
    std::vector<int> medial(n_rim);
    const double r_disc = radius - length / 2.;
    const double shrink = r_disc / radius;
    for (int i = 0; i < n_rim; ++i) {
      const double x = V.at(bottom[i]).x * shrink;
      const double y = V.at(bottom[i]).y * shrink;
      medial[i] = static_cast<int>(V.size());
      V.emplace_back(x, y, 0);
    }
    int i = n_rim - 1;
    for (int j = 0; j < n_rim; ++j) {
      Append(detail::SplitTriangularPrismToTetrahedra(c_lo, bottom[i], bottom[j], center, medial[i], medial[j]),
             &E);
      Append(detail::SplitTriangularPrismToTetrahedra(center, medial[i], medial[j], c_hi, top[i], top[j]), &E);
      Append(detail::SplitTriangularPrismToTetrahedra(bottom[i], medial[i], top[i], bottom[j], medial[j], top[j]), &E);
      i = j;
    }
  }
  return VolumeMesh(std::move(E), std::move(V));
}

// make_capsule_mesh.cc:13-180.
inline VolumeMesh MakeCapsuleVolumeMesh(double radius, double length, double resolution_hint) {
  if (!(resolution_hint > 0.0) || !(radius > 0.0) || !(length > 0.0)) throw std::invalid_argument("capsule mesh");
  using detail::Append;
  const double zm = length / 2.;
  const double top_z = zm + radius;
  const double bottom_z = -top_z;
  const int n_rim = static_cast<int>(std::clamp(2. * M_PI * radius / resolution_hint, 3.0, 706.0));
  const int n_lat = n_rim / 2;
  std::vector<Vec3> V;
  const int m_hi = static_cast<int>(V.size());
  V.emplace_back(0, 0, zm);
  const int m_lo = static_cast<int>(V.size());
  V.emplace_back(0, 0, (-zm));
  const int top = static_cast<int>(V.size());
  V.emplace_back(0, 0, top_z);
  const int bottom = static_cast<int>(V.size());
  V.emplace_back(0, 0, bottom_z);
  std::vector<int> cap_hi(n_lat * n_rim);
  std::vector<int> cap_lo(n_lat * n_rim);
  const double theta_step = M_PI_2 / n_lat;
  const double phi_step = 2 * M_PI / n_rim;
  for (int i = 0; i < n_lat; ++i) {
    const double theta = M_PI_2 - i * theta_step;
    const double s = std::sin(theta);
    const double z_hi = radius * std::cos(theta) + zm;
    const double z_lo = -z_hi;
    for (int j = 0; j < n_rim; ++j) {
      const double phi = j * phi_step;
      const double x = radius * s * std::cos(phi);
      const double y = radius * s * std::sin(phi);
      cap_hi[n_rim * i + j] = static_cast<int>(V.size());
      V.emplace_back(x, y, z_hi);
      cap_lo[n_rim * i + j] = static_cast<int>(V.size());
      V.emplace_back(x, y, z_lo);
    }
  }
  std::vector<Tet> E;
  for (int i = 0; i < n_lat - 1; ++i) {
    for (int j = 0; j < n_rim; ++j) {
      const int j1 = (j + 1) % n_rim;
      Append(detail::SplitPyramidToTetrahedra(cap_hi[(i + 1) * n_rim + j],
                                              cap_hi[(i + 1) * n_rim + j1],
                                              cap_hi[i * n_rim + j1], cap_hi[i * n_rim + j],
                                              m_hi),
             &E);
      Append(detail::SplitPyramidToTetrahedra(cap_lo[i * n_rim + j],
                                              cap_lo[i * n_rim + j1],
                                              cap_lo[(i + 1) * n_rim + j1],
                                              cap_lo[(i + 1) * n_rim + j], m_lo),
             &E);
    }
  }
  const int last = (n_lat - 1) * n_rim;
  for (int j = 0; j < n_rim; ++j) {
    const int j1 = (j + 1) % n_rim;
    E.push_back({top, cap_hi[last + j1], cap_hi[last + j], m_hi});
    E.push_back({bottom, cap_lo[last + j], cap_lo[last + j1], m_lo});
    Append(detail::SplitTriangularPrismToTetrahedra(m_lo, cap_lo[j], cap_lo[j1], m_hi, cap_hi[j],
                                                    cap_hi[j1]),
           &E);
  }
  return VolumeMesh(std::move(E), std::move(V));
}

namespace detail {
// DistanceToPoint::operator()(Cylinder) at the identity pose (distance_to_point_callback.cc:194-288) over the 2-D
// ComputeDistanceToBox (:406-482): the signed distance only.
inline double CylinderSignedDistance(double radius, double length, const Vec3& p_GQ) {
  const double r_Q = std::sqrt(p_GQ.x * p_GQ.x + p_GQ.y * p_GQ.y);
  const bool near_center_line = r_Q < DistanceToPointRelativeTolerance(radius);
  const double v0 = near_center_line ? 1.0 : p_GQ.x / r_Q;
  const double v1 = near_center_line ? 0.0 : p_GQ.y / r_Q;
  const double h[2] = {radius, length / 2.0};
  const double p[2] = {r_Q, p_GQ.z};
  enum class Loc { kInside, kBoundary, kOutside } loc[2];
  double c[2];
  for (int i = 0; i < 2; ++i) {
    const double tol = DistanceToPointRelativeTolerance(h[i]);
    if (std::abs(p[i]) > h[i] + tol) {
      loc[i] = Loc::kOutside;
      c[i] = Sign(p[i]) * h[i];
    } else if (std::abs(p[i]) >= h[i] - tol) {
      loc[i] = Loc::kBoundary;
      c[i] = Sign(p[i]) * h[i];
    } else {
      loc[i] = Loc::kInside;
      c[i] = p[i];
    }
  }
  double n[2] = {c[0], c[1]}, grad[2] = {0, 0};
  if (loc[0] == Loc::kOutside || loc[1] == Loc::kOutside) {
    const double d0 = p[0] - c[0], d1 = p[1] - c[1];
    const double len = std::sqrt(d0 * d0 + d1 * d1);
    grad[0] = d0 / len;
    grad[1] = d1 / len;
  } else if (loc[0] == Loc::kBoundary || loc[1] == Loc::kBoundary) {
    for (int i = 0; i < 2; ++i)
      if (loc[i] == Loc::kBoundary) grad[i] = Sign(c[i]);
    const double len = std::sqrt(grad[0] * grad[0] + grad[1] * grad[1]);
    grad[0] = grad[0] / len;
    grad[1] = grad[1] / len;
  } else {
    double min_dist = std::numeric_limits<double>::infinity();
    int axis = -1;
    for (int i = 0; i < 2; ++i)
      for (const double bound : {h[i], -h[i]}) {
        const double dist = std::abs(bound - p[i]);
        if (dist < min_dist) {
          min_dist = dist;
          axis = i;
        }
      }
    const double sign = Sign(p[axis]);
    n[axis] = sign * h[axis];
    grad[axis] = sign;
  }
  const Vec3 p_GN{n[0] * v0, n[0] * v1, n[1]};
  const Vec3 grad_G{grad[0] * v0, grad[0] * v1, grad[1]};
  return dot(grad_G, p_GQ - p_GN);
}
}  // namespace detail

// make_cylinder_field.cc:18-75.
inline VolumeField MakeCylinderPressureField(double radius, double length, const VolumeMesh& mesh, double modulus,
                                             double margin = 0.0) {
  if (!(modulus > 0) || !(radius > margin) || !(length > 2 * margin)) throw std::invalid_argument("cylinder field");
  const double min_half_size = std::min(radius, length / 2.0);
  std::vector<double> p_v;
  p_v.reserve(mesh.verts.size());
  for (const Vec3& r_CV : mesh.verts) {
    const double signed_distance = detail::CylinderSignedDistance(radius, length, r_CV);
    const double extent = (-signed_distance - margin) / (min_half_size - margin);
    p_v.push_back(std::min(modulus * extent, modulus));
  }
  const double min_pressure = -modulus * margin / (min_half_size - margin);
  // CollectUniqueVertices(IdentifyBoundaryFaces(tetrahedra)): the vertices of the converted surface mesh are exactly
  // those; match them back by position (ConvertVolumeToSurfaceMesh keeps their order and coordinates).
  std::map<std::array<int, 3>, int> face_count;
  static const int kFaces[4][3] = {{1, 2, 3}, {3, 2, 0}, {1, 3, 0}, {2, 1, 0}};
  for (const Tet& t : mesh.tets)
    for (const auto& f : kFaces) {
      std::array<int, 3> key{t[f[0]], t[f[1]], t[f[2]]};
      std::sort(key.begin(), key.end());
      ++face_count[key];
    }
  for (const auto& [key, count] : face_count)
    if (count == 1)
      for (int bv : key) p_v[bv] = min_pressure;
  return VolumeField(std::move(p_v), mesh);
}

// make_capsule_field.h:55-90.
inline VolumeField MakeCapsulePressureField(double radius, double length, const VolumeMesh& mesh, double modulus,
                                            double margin = 0.0) {
  if (!(modulus > 0) || !(radius > margin)) throw std::invalid_argument("capsule field");
  const Vec3& a = mesh.verts.at(0);
  const Vec3& b = mesh.verts.at(1);
  if (!(a.x == 0 && a.y == 0 && a.z == length / 2 && b.x == 0 && b.y == 0 && b.z == -length / 2))
    throw std::invalid_argument("capsule field: vertices 0 and 1 must be the ends of the medial axis");
  const double pressure_gradient = modulus / (radius - margin);
  const double surface_pressure = -margin * pressure_gradient;
  std::vector<double> p_v(mesh.verts.size(), surface_pressure);
  p_v[0] = p_v[1] = modulus;
  return VolumeField(std::move(p_v), mesh);
}

}  // namespace ho

// ---- Convex / Mesh (oracle, test infrastructure) ------------------------------------------------------------------------
//   make_convex_mesh.cc:17-60                MakeConvexVolumeMesh (CalcCentroidOfEnclosedVolume)
//   calc_distance_to_surface_mesh.cc:30-120  CalcSquaredDistanceToLineSegment / ToTriangle (the Bvh<Obb> of :122-200 only
//                                            prunes; the minimum over all triangles is the same number)
//   make_mesh_field.cc:45-96                 MakeVolumeMeshPressureField
// Pinned by make_mesh_field_test.cc (non_convex_mesh.vtk: vertex 0 carries the modulus, 1-5 zero; WithMargin;
// NoInteriorVertex), calc_distance_to_surface_mesh_test.cc (the seven Voronoi regions of a posed triangle) and
// make_convex_mesh_test.cc (CoherentVertices), restated in tests/test_oracle_cylinder_capsule.py.
namespace ho {
namespace detail {

inline double CalcSquaredDistanceToLineSegment(const Vec3& Q, const Vec3& A, const Vec3& B) {
  const Vec3 AB = B - A;
  const Vec3 AQ = Q - A;
  if (dot(AQ, AB) <= 0.0) return dot(AQ, AQ);
  const Vec3 BQ = Q - B;
  if (dot(BQ, AB) >= 0.0) return dot(BQ, BQ);
  const Vec3 n = cross(AB, AQ);
  return dot(n, n) / dot(AB, AB);
}

inline double CalcSquaredDistanceToTriangle(const Vec3& p_WQ, const Vec3& A, const Vec3& B, const Vec3& C) {
  const Vec3 p_AB_W = B - A;
  const Vec3 p_AC_W = C - A;
  const Vec3 Tx_W = p_AB_W / norm(p_AB_W);
  const Vec3 N = cross(p_AB_W, p_AC_W);
  const Vec3 Tz_W = N / norm(N);
  const Vec3 Ty_W = cross(Tz_W, Tx_W);
  auto to_T = [&](const Vec3& v) { return Vec3(dot(Tx_W, v), dot(Ty_W, v), dot(Tz_W, v)); };
  const Vec3 p_TB = to_T(p_AB_W);
  const Vec3 p_TC = to_T(p_AC_W);
  const Vec3 p_TQ = to_T(p_WQ - A);
  // IsInTriangle: Q = b0 B + b1 C in the x-y plane of T (Cramer's rule; the reference solves the same 2 x 2 by QR).
  const double det = p_TB.x * p_TC.y - p_TC.x * p_TB.y;
  const double b0 = (p_TQ.x * p_TC.y - p_TC.x * p_TQ.y) / det;
  const double b1 = (p_TB.x * p_TQ.y - p_TQ.x * p_TB.y) / det;
  if (b0 >= 0.0 && b1 >= 0.0 && b0 + b1 <= 1.0) return p_TQ.z * p_TQ.z;
  const Vec3 p_TA(0, 0, 0);
  return std::min({CalcSquaredDistanceToLineSegment(p_TQ, p_TA, p_TB), CalcSquaredDistanceToLineSegment(p_TQ, p_TA, p_TC),
                   CalcSquaredDistanceToLineSegment(p_TQ, p_TB, p_TC)});
}

}  // namespace detail

inline double CalcDistanceToSurfaceMesh(const Vec3& p_WQ, const TriMesh& mesh_W) {
  double min_squared = std::numeric_limits<double>::infinity();
  for (const Tri& t : mesh_W.tris)
    min_squared = std::min(min_squared, detail::CalcSquaredDistanceToTriangle(p_WQ, mesh_W.verts[t[0]], mesh_W.verts[t[1]],
                                                                             mesh_W.verts[t[2]]));
  return std::sqrt(min_squared);
}

inline VolumeMesh MakeConvexVolumeMesh(const TriMesh& surface) {
  double six_total_volume = 0;
  Vec3 centroid(0, 0, 0);
  for (const Tri& tri : surface.tris) {
    const Vec3& p = surface.verts[tri[0]];
    const Vec3& q = surface.verts[tri[1]];
    const Vec3& r = surface.verts[tri[2]];
    const double six_volume = dot(cross(p, q), r);
    six_total_volume += six_volume;
    centroid = centroid + six_volume * ((p + q) + r);
  }
  if (!(six_total_volume > 0)) throw std::invalid_argument("convex mesh: no enclosed volume");
  std::vector<Vec3> V(surface.verts.begin(), surface.verts.end());
  V.push_back(centroid / (4 * six_total_volume));
  const int c = static_cast<int>(V.size()) - 1;
  std::vector<Tet> E;
  for (const Tri& e : surface.tris) E.push_back({c, e[0], e[1], e[2]});
  return VolumeMesh(std::move(E), std::move(V));
}

inline VolumeField MakeVolumeMeshPressureField(const VolumeMesh& mesh, double modulus, double margin = 0.0) {
  if (!(modulus > 0)) throw std::invalid_argument("mesh field");
  const TriMesh surface = ConvertVolumeToSurfaceMesh(mesh);
  std::vector<char> on_boundary(mesh.verts.size(), 0);
  {
    std::map<std::array<int, 3>, int> face_count;
    static const int kFaces[4][3] = {{1, 2, 3}, {3, 2, 0}, {1, 3, 0}, {2, 1, 0}};
    for (const Tet& t : mesh.tets)
      for (const auto& f : kFaces) {
        std::array<int, 3> key{t[f[0]], t[f[1]], t[f[2]]};
        std::sort(key.begin(), key.end());
        ++face_count[key];
      }
    for (const auto& [key, count] : face_count)
      if (count == 1)
        for (int bv : key) on_boundary[bv] = 1;
  }
  std::vector<double> values;
  double max_value = std::numeric_limits<double>::lowest();
  for (size_t v = 0; v < mesh.verts.size(); ++v) {
    if (on_boundary[v]) {
      values.push_back(0);
      continue;
    }
    const double distance = CalcDistanceToSurfaceMesh(mesh.verts[v], surface) + margin;
    values.push_back(distance);
    max_value = std::max(distance, max_value);
  }
  if (max_value <= 0)
    throw std::runtime_error("MakeVolumeMeshPressureField(): the computed max distance to boundary among all mesh vertices is "
                             "non-positive. Perhaps the mesh lacks interior vertices.");
  for (double& p : values) p = modulus * (p - margin) / (max_value - margin);
  return VolumeField(std::move(values), mesh);
}

}  // namespace ho
