// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// Restatement of the per-element-pair clipping, the mesh builders, the derived
// mesh quantities and the ContactSurface assembly:
//   geometry/proximity/mesh_intersection.cc, field_intersection.cc,
//   mesh_plane_intersection.cc, contact_surface_utility.cc, plane.{h,cc},
//   posed_half_space.h, polygon_surface_mesh.cc, triangle_surface_mesh.h,
//   mesh_field_linear.h, query_results/contact_surface.cc.
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <utility>
#include <set>
#include <map>
#include <vector>

#include "ho_mesh.h"

namespace ho {

enum class Representation { kTriangle = 0, kPolygon = 1 };  // contact_surface.h:24

// Plane<double> (plane.cc:12-35): stores unit normal + displacement d = n̂·p.
struct Plane {
  Vec3 nhat;
  double displacement{0};
  Plane() = default;
  Plane(const Vec3& normal, const Vec3& point, bool already_normalized = false) {
    if (!already_normalized) {
      const double magnitude = norm(normal);
      if (magnitude < 1e-10) throw std::runtime_error("Plane: normal magnitude too small");
      nhat = normal / magnitude;
    } else {
      nhat = normal;
    }
    displacement = dot(nhat, point);
  }
  double CalcHeight(const Vec3& q) const { return dot(nhat, q) - displacement; }  // plane.h:69-71
};
using PosedHalfSpace = Plane;  // posed_half_space.h:35-70: CalcSignedDistance == CalcHeight

// mesh_intersection.cc:49-58.
inline Vec3 CalcIntersection(const Vec3& p_FA, const Vec3& p_FB, const PosedHalfSpace& H_F) {
  const double a = H_F.CalcHeight(p_FA);
  const double b = H_F.CalcHeight(p_FB);
  const double wa = b / (b - a);
  const double wb = 1.0 - wa;
  return wa * p_FA + wb * p_FB;
}

// mesh_intersection.cc:76-119.
inline void ClipPolygonByHalfSpace(const std::vector<Vec3>& in, const PosedHalfSpace& H_F, std::vector<Vec3>* out) {
  out->clear();
  const int size = static_cast<int>(in.size());
  bool contained[16];
  if (size > 16) throw std::logic_error("ClipPolygonByHalfSpace: too many vertices");
  for (int i = 0; i < size; ++i) contained[i] = H_F.CalcHeight(in[i]) <= 0;
  int previous_index = size - 1;
  for (int i = 0; i < size; ++i) {
    const Vec3& previous = in[previous_index];
    const Vec3& current = in[i];
    const bool previous_contained = contained[previous_index];
    const bool current_contained = contained[i];
    if (current_contained) {
      if (!previous_contained) out->push_back(CalcIntersection(current, previous, H_F));
      out->push_back(current);
    } else if (previous_contained) {
      out->push_back(CalcIntersection(current, previous, H_F));
    }
    previous_index = i;
  }
}

// mesh_intersection.cc:121-170.
inline void RemoveNearlyDuplicateVertices(std::vector<Vec3>* polygon, std::vector<int>* companion = nullptr) {
  // `companion` is NOT touched by the reference (field_intersection.cc:192
  // dedupes only the polygon and returns the face list unchanged); it is
  // accepted here only so callers can document that fact.
  (void)companion;
  if (polygon->size() <= 1) return;
  auto near = [](const Vec3& p, const Vec3& q) {
    const double kEpsSquared(1e-14 * 1e-14);
    return squared_norm(p - q) < kEpsSquared;
  };
  // std::unique semantics: keep first of each run, comparing against the last kept.
  size_t result = 0;
  for (size_t first = 1; first < polygon->size(); ++first) {
    if (!near((*polygon)[result], (*polygon)[first])) {
      ++result;
      if (result != first) (*polygon)[result] = (*polygon)[first];
    }
  }
  polygon->resize(result + 1);
  if (polygon->size() >= 3) {
    if (near((*polygon)[0], polygon->back())) polygon->pop_back();
  }
}

// mesh_intersection.cc:176-264. Returns the polygon in frame M (empty if < 3 vertices).
inline std::vector<Vec3> ClipTriangleByTetrahedron(int element, const VolumeMesh& volume_M, int face,
                                                   const TriMesh& surface_N, const Xform& X_MN) {
  std::vector<Vec3> buf[2];
  std::vector<Vec3>* in_M = &buf[0];
  std::vector<Vec3>* out_M = &buf[1];
  for (int i = 0; i < 3; ++i) in_M->push_back(X_MN * surface_N.verts[surface_N.tris[face][i]]);
  Vec3 p_MVs[4];
  for (int i = 0; i < 4; ++i) p_MVs[i] = volume_M.verts[volume_M.tets[element][i]];
  static const int faces[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
  for (const auto& fv : faces) {
    const Vec3& p_MA = p_MVs[fv[0]];
    const Vec3& p_MB = p_MVs[fv[1]];
    const Vec3& p_MC = p_MVs[fv[2]];
    const Vec3 normal_M = cross(p_MB - p_MA, p_MC - p_MA);
    const PosedHalfSpace half_space_M(normal_M, p_MA);
    ClipPolygonByHalfSpace(*in_M, half_space_M, out_M);
    std::swap(in_M, out_M);
  }
  RemoveNearlyDuplicateVertices(in_M);
  if (in_M->size() < 3) in_M->clear();
  if (in_M->size() > 7) throw std::logic_error("ClipTriangleByTetrahedron: more than 7 vertices");
  return *in_M;
}

inline double CosAlpha() {
  constexpr double kAlpha = 5. * M_PI / 8.;
  static const double kCosAlpha = std::cos(kAlpha);
  return kCosAlpha;
}

// contact_surface_utility.cc:336-354.
inline bool IsFaceNormalInNormalDirection(const Vec3& normal_F, const TriMesh& surface_M, int tri_index,
                                          const Mat3& R_FM) {
  const Vec3 face_normal_F = R_FM * surface_M.normals[tri_index];
  const double cos_theta = dot(normal_F, normalized(face_normal_F));
  return cos_theta > CosAlpha();
}

// mesh_intersection.cc:266-288.
inline bool IsFaceNormalAlongPressureGradient(const VolumeField& field_M, const TriMesh& surface_N, const Xform& X_MN,
                                              int tet_index, int tri_index) {
  const Vec3 grad_p_M = field_M.gradients[tet_index];
  return IsFaceNormalInNormalDirection(normalized(grad_p_M), surface_N, tri_index, X_MN.R);
}

// contact_surface_utility.cc:218-286.
inline Vec3 CalcPolygonCentroid(const std::vector<int>& polygon, const Vec3& n_F, const std::vector<Vec3>& vertices_F) {
  const int v_count = static_cast<int>(polygon.size());
  auto triangle_centroid = [&](int v0, int v1, int v2) { return (vertices_F[v0] + vertices_F[v1] + vertices_F[v2]) / 3; };
  if (v_count == 3) return triangle_centroid(polygon[0], polygon[1], polygon[2]);
  auto triangle_weight = [&](int v0, int v1, int v2) {
    const Vec3& r0 = vertices_F[v0];
    const Vec3& r1 = vertices_F[v1];
    const Vec3& r2 = vertices_F[v2];
    return dot(cross(r1 - r0, r2 - r0), n_F);
  };
  Vec3 accum{0, 0, 0};
  double total_weight = 0;
  const int v0 = polygon[0];
  int v2 = polygon[1];
  for (int i = 2; i < v_count; ++i) {
    const int v1 = v2;
    v2 = polygon[i];
    const double weight = triangle_weight(v0, v1, v2);
    accum += weight * triangle_centroid(v0, v1, v2);
    total_weight += weight;
  }
  if (total_weight == 0) {
    accum = Vec3{0, 0, 0};
    for (int i = 0; i < v_count; ++i) accum += vertices_F[polygon[i]];
    total_weight = v_count;
  }
  return accum / total_weight;
}

// Mesh builders (contact_surface_utility.{h,cc}). One struct serves both
// representations; `rep` selects TriMeshBuilder or PolyMeshBuilder behaviour.
struct MeshBuilder {
  Representation rep;
  std::vector<Vec3> vertices_B;
  std::vector<double> pressures;
  std::vector<int> face_data;          // polygon rep: |n|v0|..|vn-1| ...
  std::vector<Tri> faces;              // triangle rep
  std::vector<Vec3> grad_e_per_face;   // polygon rep only (contact_surface_utility.cc:203)
  int num_faces_{0};

  explicit MeshBuilder(Representation r) : rep(r) {}
  int AddVertex(const Vec3& p_BV, double field_value) {
    vertices_B.push_back(p_BV);
    pressures.push_back(field_value);
    return static_cast<int>(vertices_B.size()) - 1;
  }
  // Returns the number of faces added.
  int AddPolygon(const std::vector<int>& polygon_vertices, const Vec3& nhat_B, const Vec3& grad_e_MN_B) {
    if (rep == Representation::kPolygon) {
      // contact_surface_utility.cc:186-205, 318-328.
      face_data.push_back(static_cast<int>(polygon_vertices.size()));
      for (int v : polygon_vertices) face_data.push_back(v);
      grad_e_per_face.push_back(grad_e_MN_B);
      ++num_faces_;
      return 1;
    }
    // contact_surface_utility.cc:137-164, 288-316.
    const Vec3 p_FC = CalcPolygonCentroid(polygon_vertices, nhat_B, vertices_B);
    const int centroid_index = static_cast<int>(vertices_B.size());
    vertices_B.push_back(p_FC);
    int v2 = polygon_vertices.back();
    const int n = static_cast<int>(polygon_vertices.size());
    for (int i = 0; i < n; ++i) {
      const int v1 = v2;
      v2 = polygon_vertices[i];
      faces.push_back({v1, v2, centroid_index});
    }
    const Vec3& p_BC = vertices_B.back();
    const Vec3& p_BN = vertices_B[polygon_vertices[0]];
    const double pressure_at_N = pressures[polygon_vertices[0]];
    pressures.push_back(dot(grad_e_MN_B, p_BC - p_BN) + pressure_at_N);
    num_faces_ += n;
    return n;
  }
  int num_faces() const { return num_faces_; }
};

// The flattened result: everything a ContactSurface<double> exposes
// (query_results/contact_surface.h:160-415) plus the emitting element pair per
// face (the reference's own notion of equality, field_intersection_test.cc:397-421).
struct SurfaceResult {
  int64_t id_M{0}, id_N{0};
  Representation rep{Representation::kPolygon};
  std::vector<Vec3> vertices_W;
  std::vector<double> pressures;   // field value at each vertex
  std::vector<int> face_data;      // polygon rep
  std::vector<Tri> tris;           // triangle rep
  std::vector<double> areas;       // per face
  std::vector<Vec3> normals_W;     // per face
  std::vector<Vec3> centroids_W;   // per face
  std::vector<double> e_centroid;  // field value at the face centroid (discrete_update_manager.cc:1011-1015)
  std::vector<Vec3> field_grad_W;  // polygon rep: per-face field gradient
  std::vector<double> field_at_Wo; // polygon rep: per-face value at World origin
  bool has_grad_eM{false}, has_grad_eN{false};
  std::vector<Vec3> grad_eM_W, grad_eN_W;  // per face
  double total_area{0};
  Vec3 centroid_W;
  std::vector<int> elem_first, elem_second;  // per face: (tet, tri) or (tet0, tet1)
  int num_faces() const { return static_cast<int>(areas.size()); }
};

// Builds mesh-derived quantities in the builder's frame B (= M), then
// re-expresses in World and fills `out` (everything but ids / constituent
// gradients). Follows PolygonSurfaceMesh ctor + TransformVertices
// (polygon_surface_mesh.cc:25-37,97-236), MeshFieldLinear ctor + Transform
// (mesh_field_linear.h:218-229,377-383,502-509), TriangleSurfaceMesh
// (triangle_surface_mesh.h:190-198,452-487).
// `already_in_world`: the builder's frame B is World (the half-space paths build their mesh in
// World directly, mesh_plane_intersection.cc:236-249, mesh_half_space_intersection.cc:84-86): no
// TransformVertices / field Transform step, so normals are not re-normalised either.
inline void MakeMeshAndFieldInWorld(MeshBuilder&& b, const Xform& X_WM, SurfaceResult* out,
                                    bool already_in_world = false) {
  out->rep = b.rep;
  std::vector<Vec3>& verts = b.vertices_B;
  if (b.rep == Representation::kPolygon) {
    // --- PolygonSurfaceMesh::ComputePositionDependentQuantities (frame M).
    std::vector<int> poly_indices;
    double total_area = 0;
    Vec3 p_MSc{0, 0, 0};
    std::vector<Vec3> centroids_M;
    int i = 0;
    while (i < static_cast<int>(b.face_data.size())) {
      poly_indices.push_back(i);
      const int v_count = b.face_data[i];
      const int off = i + 1;
      const Vec3& r_MA = verts[b.face_data[off]];
      Vec3 p_MTc_scaled{0, 0, 0};
      Vec3 normal_M{0, 0, 0};
      double double_poly_area = 0;
      for (int v = 1; v < v_count - 1; ++v) {
        const Vec3& r_MB = verts[b.face_data[off + v]];
        const Vec3& r_MC = verts[b.face_data[off + v + 1]];
        const Vec3 c = cross(r_MB - r_MA, r_MC - r_MA);
        normal_M += c;
        const double cross_magnitude = norm(c);
        double_poly_area += cross_magnitude;
        p_MTc_scaled += cross_magnitude * ((r_MA + r_MB) + r_MC);
      }
      const double poly_area = double_poly_area / 2;
      out->areas.push_back(poly_area);
      out->normals_W.push_back(normalized(normal_M));  // still in M here
      const Vec3 p_MTc_area_scaled = p_MTc_scaled / 6;
      if (poly_area != 0.0) {
        centroids_M.push_back(p_MTc_area_scaled / poly_area);
      } else {
        Vec3 acc = r_MA;
        for (int v = 1; v < v_count; ++v) acc += verts[b.face_data[off + v]];
        centroids_M.push_back(acc / v_count);
      }
      const double old_area = total_area;
      total_area += poly_area;
      if (total_area != 0.0) {
        p_MSc = (p_MSc * old_area + p_MTc_area_scaled) / total_area;
      } else {
        p_MSc = centroids_M.back();
      }
      i += v_count + 1;
    }
    // --- MeshFieldLinear(values, mesh, gradients): value at Mo per face.
    const int nf = static_cast<int>(poly_indices.size());
    std::vector<double> values_at_Mo(nf);
    for (int f = 0; f < nf; ++f) {
      const int v0 = b.face_data[poly_indices[f] + 1];
      values_at_Mo[f] = b.pressures[v0] - dot(b.grad_e_per_face[f], verts[v0]);
    }
    // --- TransformVertices(X_WM) + field.Transform(X_WM).
    if (!already_in_world) {
      for (Vec3& v : verts) v = X_WM * v;
      for (Vec3& n : out->normals_W) n = normalized(X_WM.R * n);
    }
    out->centroids_W.reserve(nf);
    for (const Vec3& c : centroids_M) out->centroids_W.push_back(already_in_world ? c : X_WM * c);
    out->centroid_W = already_in_world ? p_MSc : X_WM * p_MSc;
    out->total_area = total_area;
    out->field_grad_W.resize(nf);
    out->field_at_Wo.resize(nf);
    out->e_centroid.resize(nf);
    for (int f = 0; f < nf; ++f) {
      const Vec3 g = already_in_world ? b.grad_e_per_face[f] : X_WM.R * b.grad_e_per_face[f];
      out->field_grad_W[f] = g;
      out->field_at_Wo[f] = already_in_world ? values_at_Mo[f] : values_at_Mo[f] - dot(g, X_WM.p);
      out->e_centroid[f] = dot(g, out->centroids_W[f]) + out->field_at_Wo[f];  // EvaluateCartesian(f, p_WC)
    }
    out->face_data = std::move(b.face_data);
  } else {
    // --- TriangleSurfaceMesh::ComputePositionDependentQuantities (frame M).
    const int nf = static_cast<int>(b.faces.size());
    out->areas.resize(nf);
    out->normals_W.resize(nf);
    double total_area = 0;
    Vec3 p_MSc{0, 0, 0};
    for (int f = 0; f < nf; ++f) {
      const Vec3& A = verts[b.faces[f][0]];
      const Vec3& B = verts[b.faces[f][1]];
      const Vec3& C = verts[b.faces[f][2]];
      const Vec3 c = cross(B - A, C - A);
      const double n = norm(c);
      const double face_area = 0.5 * n;
      out->areas[f] = face_area;
      total_area += face_area;
      out->normals_W[f] = (n != 0.0) ? c / n : c;
      p_MSc += face_area * ((A + B) + C);
    }
    if (total_area != 0.) p_MSc = p_MSc / (3. * total_area);
    // --- TransformVertices(X_WM): normals are NOT renormalised for triangles.
    if (!already_in_world) {
      for (Vec3& v : verts) v = X_WM * v;
      for (Vec3& n : out->normals_W) n = X_WM.R * n;
    }
    out->centroid_W = already_in_world ? p_MSc : X_WM * p_MSc;
    out->total_area = total_area;
    out->centroids_W.resize(nf);
    out->e_centroid.resize(nf);
    const double third = 1.0 / 3.0;
    for (int f = 0; f < nf; ++f) {
      const Tri& t = b.faces[f];
      out->centroids_W[f] = ((verts[t[0]] + verts[t[1]]) + verts[t[2]]) / 3;  // element_centroid()
      // tri_e_MN().Evaluate(f, (1/3,1/3,1/3)) (mesh_field_linear.h:300-312).
      double value = third * b.pressures[t[0]];
      value += third * b.pressures[t[1]];
      value += third * b.pressures[t[2]];
      out->e_centroid[f] = value;
    }
    out->tris = std::move(b.faces);
  }
  out->vertices_W = std::move(verts);
  out->pressures = std::move(b.pressures);
}

// ContactSurface ctor + SwapMAndN (contact_surface.cc:12-59,
// polygon_surface_mesh.cc:39-53, triangle_surface_mesh.cc:39-47).
inline void EnforceIdOrder(SurfaceResult* s) {
  if (!(s->id_N < s->id_M)) return;
  std::swap(s->id_M, s->id_N);
  if (s->rep == Representation::kPolygon) {
    int i = 0;
    while (i < static_cast<int>(s->face_data.size())) {
      const int v_count = s->face_data[i];
      int f_0 = i;
      int f_n = i + v_count + 1;
      for (int k = 0; k < v_count / 2; ++k) std::swap(s->face_data[++f_0], s->face_data[--f_n]);
      i += v_count + 1;
    }
  } else {
    for (Tri& t : s->tris) std::swap(t[0], t[1]);
  }
  for (Vec3& n : s->normals_W) n = -n;
  std::swap(s->grad_eM_W, s->grad_eN_W);
  std::swap(s->has_grad_eM, s->has_grad_eN);
}

// ---------------------------------------------------------------------------
// Soft–rigid (mesh_intersection.cc:290-440).

// mesh_intersection.cc:321-382. Returns number of faces added.
inline int CalcContactPolygonTetTri(const VolumeMesh& mesh_M, const VolumeField& field_M, const TriMesh& surface_N,
                                    const Xform& X_MN, MeshBuilder* builder_M, int tet_index, int tri_index,
                                    std::vector<Vec3>* grad_eM_Ms,
                                    bool filter_face_normal_along_field_gradient = true) {
  if (filter_face_normal_along_field_gradient) {
    if (!IsFaceNormalAlongPressureGradient(field_M, surface_N, X_MN, tet_index, tri_index)) return 0;
  }
  const std::vector<Vec3> polygon_vertices_M = ClipTriangleByTetrahedron(tet_index, mesh_M, tri_index, surface_N, X_MN);
  if (polygon_vertices_M.size() < 3) return 0;
  std::vector<int> idx;
  for (const Vec3& p_MV : polygon_vertices_M)
    idx.push_back(builder_M->AddVertex(p_MV, field_M.EvaluateCartesian(tet_index, p_MV)));
  const Vec3 nhat_M = X_MN.R * surface_N.normals[tri_index];
  const Vec3 grad_e_MN_M = field_M.gradients[tet_index];
  const int num_new_faces = builder_M->AddPolygon(idx, nhat_M, grad_e_MN_M);
  for (int i = 0; i < num_new_faces; ++i) grad_eM_Ms->push_back(grad_e_MN_M);
  return num_new_faces;
}

// ComputeContactSurfaceFromCompliantVolumeRigidSurface (mesh_intersection.cc:384-440)
// over an explicit candidate list. Returns false when the surface is null.
inline bool ContactSurfaceFromCompliantVolumeRigidSurface(int64_t id_S, const VolumeMesh& mesh_S,
                                                          const VolumeField& field_S, const Xform& X_WS, int64_t id_R,
                                                          const TriMesh& mesh_R, const Xform& X_WR,
                                                          const std::vector<std::pair<int, int>>& candidates,
                                                          Representation rep, SurfaceResult* out,
                                                          bool filter_face_normal_along_field_gradient = true) {
  const Xform X_SR = InvertAndCompose(X_WS, X_WR);
  MeshBuilder builder(rep);
  std::vector<Vec3> grad_eS_S;
  std::vector<int> e0, e1;
  for (const auto& [tet, tri] : candidates) {
    const int n = CalcContactPolygonTetTri(mesh_S, field_S, mesh_R, X_SR, &builder, tet, tri, &grad_eS_S,
                                           filter_face_normal_along_field_gradient);
    for (int i = 0; i < n; ++i) {
      e0.push_back(tet);
      e1.push_back(tri);
    }
  }
  if (builder.num_faces() == 0) return false;
  *out = SurfaceResult{};
  MakeMeshAndFieldInWorld(std::move(builder), X_WS, out);
  out->grad_eM_W.reserve(grad_eS_S.size());
  for (const Vec3& g : grad_eS_S) out->grad_eM_W.push_back(X_WS.R * g);
  out->has_grad_eM = true;
  out->has_grad_eN = false;
  out->id_M = id_S;
  out->id_N = id_R;
  out->elem_first = std::move(e0);
  out->elem_second = std::move(e1);
  EnforceIdOrder(out);
  return true;
}

// ---------------------------------------------------------------------------
// Soft–soft (field_intersection.cc).

// field_intersection.cc:30-83.
inline bool CalcEquilibriumPlane(int element0, const VolumeField& field0_M, int element1, const VolumeField& field1_N,
                                 const Xform& X_MN, Plane* plane_M) {
  const Vec3 grad_f0_M = field0_M.gradients[element0];
  const double f0_Mo = field0_M.values_at_Mo[element0];
  const Vec3 grad_f1_N = field1_N.gradients[element1];
  const Vec3 grad_f1_M = X_MN.R * grad_f1_N;
  const Vec3 p_NMo = -(transpose(X_MN.R) * X_MN.p);
  const double f1_Mo = field1_N.EvaluateCartesian(element1, p_NMo);
  const Vec3 n_M = grad_f0_M - grad_f1_M;
  const double magnitude = norm(n_M);
  if (magnitude <= 0.0) return false;
  const Vec3 nhat_M = n_M / magnitude;
  const Vec3 p_MQ = ((f1_Mo - f0_Mo) / magnitude) * nhat_M;
  *plane_M = Plane(nhat_M, p_MQ, /*already_normalized=*/true);
  return true;
}

// mesh_plane_intersection.cc:19-180 (cut_edges == nullptr path).
inline void SliceTetrahedronWithPlane(int tet_index, const VolumeMesh& mesh_M, const Plane& plane_M,
                                      std::vector<Vec3>* polygon_vertices, std::vector<int>* faces) {
  static const int kTetEdges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  static const int kEdgeTable[16][4] = {{-1, -1, -1, -1}, {0, 3, 2, -1}, {0, 1, 4, -1}, {4, 3, 2, 1},
                                        {1, 2, 5, -1},    {0, 3, 5, 1},  {0, 2, 5, 4},  {3, 5, 4, -1},
                                        {3, 4, 5, -1},    {4, 5, 2, 0},  {1, 5, 3, 0},  {1, 5, 2, -1},
                                        {1, 2, 3, 4},     {0, 4, 1, -1}, {0, 2, 3, -1}, {-1, -1, -1, -1}};
  static const int kFaceTable[16][4] = {{-1, -1, -1, -1}, {2, 1, 3, -1}, {3, 0, 2, -1}, {2, 1, 3, 0},
                                        {3, 1, 0, -1},    {2, 1, 0, 3},  {3, 1, 0, 2},  {1, 0, 2, -1},
                                        {2, 0, 1, -1},    {0, 1, 3, 2},  {0, 1, 2, 3},  {0, 1, 3, -1},
                                        {3, 1, 2, 0},     {2, 0, 3, -1}, {3, 1, 2, -1}, {-1, -1, -1, -1}};
  if (faces) faces->clear();
  double distance[4];
  int code = 0;
  for (int i = 0; i < 4; ++i) {
    distance[i] = plane_M.CalcHeight(mesh_M.verts[mesh_M.tets[tet_index][i]]);
    if (distance[i] > 0) code |= 1 << i;
  }
  if (kEdgeTable[code][0] == -1) return;
  for (int i = 0; i < 4; ++i) {
    const int edge_index = kEdgeTable[code][i];
    if (edge_index == -1) break;
    const int e0 = kTetEdges[edge_index][0], e1 = kTetEdges[edge_index][1];
    const Vec3& p_MV0 = mesh_M.verts[mesh_M.tets[tet_index][e0]];
    const Vec3& p_MV1 = mesh_M.verts[mesh_M.tets[tet_index][e1]];
    const double d_v0 = distance[e0];
    const double d_v1 = distance[e1];
    const double t = d_v0 / (d_v0 - d_v1);
    polygon_vertices->push_back(p_MV0 + t * (p_MV1 - p_MV0));
    if (faces) faces->push_back(kFaceTable[code][i]);
  }
}

// field_intersection.cc:85-198.
inline std::pair<std::vector<Vec3>, std::vector<int>> IntersectTetrahedra(int element0, const VolumeMesh& mesh0_M,
                                                                          int element1, const VolumeMesh& mesh1_N,
                                                                          const Xform& X_MN,
                                                                          const Plane& equilibrium_plane_M) {
  std::vector<Vec3> polygon_buffer[2];
  std::vector<int> face_buffer[2];
  double distances[16];
  SliceTetrahedronWithPlane(element0, mesh0_M, equilibrium_plane_M, &polygon_buffer[0], &face_buffer[0]);
  RemoveNearlyDuplicateVertices(&polygon_buffer[0]);
  if (polygon_buffer[0].size() < 3) return {};
  Vec3 p_MVs[4];
  for (int i = 0; i < 4; ++i) p_MVs[i] = X_MN * mesh1_N.verts[mesh1_N.tets[element1][i]];
  std::vector<Vec3>* current_polygon = &polygon_buffer[0];
  std::vector<Vec3>* clipped_polygon = &polygon_buffer[1];
  std::vector<int>* current_faces = &face_buffer[0];
  std::vector<int>* clipped_faces = &face_buffer[1];
  for (int face = 0; face < 4; ++face) {
    clipped_polygon->clear();
    clipped_faces->clear();
    const Vec3& p_MA = p_MVs[(face + 1) % 4];
    const Vec3 triangle_outward_normal_M = X_MN.R * (-mesh1_N.inward_normals[element1][face]);
    const PosedHalfSpace half_space_M(triangle_outward_normal_M, p_MA);
    const int size = static_cast<int>(current_polygon->size());
    if (size > 16) throw std::logic_error("IntersectTetrahedra: too many vertices");
    for (int i = 0; i < size; ++i) distances[i] = half_space_M.CalcHeight((*current_polygon)[i]);
    for (int i = 0; i < size; ++i) {
      const int j = (i + 1) % size;
      if (distances[i] <= 0) {
        clipped_polygon->push_back((*current_polygon)[i]);
        clipped_faces->push_back((*current_faces)[i]);
        if (distances[j] > 0) {
          const double wa = distances[j] / (distances[j] - distances[i]);
          const double wb = 1.0 - wa;
          clipped_polygon->push_back(wa * (*current_polygon)[i] + wb * (*current_polygon)[j]);
          clipped_faces->push_back(face + 4);
        }
      } else if (distances[j] <= 0) {
        const double wa = distances[j] / (distances[j] - distances[i]);
        const double wb = 1.0 - wa;
        clipped_polygon->push_back(wa * (*current_polygon)[i] + wb * (*current_polygon)[j]);
        clipped_faces->push_back((*current_faces)[i]);
      }
    }
    std::swap(current_polygon, clipped_polygon);
    std::swap(current_faces, clipped_faces);
    if (current_polygon->empty()) return {};
  }
  RemoveNearlyDuplicateVertices(current_polygon);
  if (current_polygon->size() < 3) return {};
  return {*current_polygon, *current_faces};
}

// field_intersection.cc:200-216.
inline bool IsPlaneNormalAlongPressureGradient(const Vec3& nhat_M, int tetrahedron, const VolumeField& field_M) {
  const Vec3 grad_p_M = field_M.gradients[tetrahedron];
  const double cos_theta = dot(nhat_M, normalized(grad_p_M));
  return cos_theta > CosAlpha();
}

// field_intersection.cc:377-428. Returns number of faces added.
inline int CalcContactPolygonTetTet(const VolumeMesh& mesh0_M, const VolumeField& field0_M, const VolumeMesh& mesh1_N,
                                    const VolumeField& field1_N, const Xform& X_MN, const Mat3& R_NM, int tet0,
                                    int tet1, MeshBuilder* builder_M) {
  Plane equilibrium_plane_M;
  if (!CalcEquilibriumPlane(tet0, field0_M, tet1, field1_N, X_MN, &equilibrium_plane_M)) return 0;
  const Vec3 polygon_nhat_M = equilibrium_plane_M.nhat;
  if (!IsPlaneNormalAlongPressureGradient(polygon_nhat_M, tet0, field0_M)) return 0;
  const Vec3 reverse_polygon_nhat_N = R_NM * (-polygon_nhat_M);
  if (!IsPlaneNormalAlongPressureGradient(reverse_polygon_nhat_N, tet1, field1_N)) return 0;
  const auto [polygon_vertices_M, faces] = IntersectTetrahedra(tet0, mesh0_M, tet1, mesh1_N, X_MN, equilibrium_plane_M);
  if (polygon_vertices_M.size() < 3) return 0;
  std::vector<int> idx;
  for (const Vec3& p_MV : polygon_vertices_M)
    idx.push_back(builder_M->AddVertex(p_MV, field0_M.EvaluateCartesian(tet0, p_MV)));
  return builder_M->AddPolygon(idx, polygon_nhat_M, field0_M.gradients[tet0]);
}

// ComputeContactSurfaceFromCompliantVolumes (field_intersection.cc:430-520,
// volume-BVH path) over an explicit candidate list.
inline bool ContactSurfaceFromCompliantVolumes(int64_t id_M, const VolumeMesh& mesh_M, const VolumeField& field_M,
                                               const Xform& X_WM, int64_t id_N, const VolumeMesh& mesh_N,
                                               const VolumeField& field_N, const Xform& X_WN,
                                               const std::vector<std::pair<int, int>>& candidates, Representation rep,
                                               SurfaceResult* out) {
  const Xform X_MN = InvertAndCompose(X_WM, X_WN);
  MeshBuilder builder(rep);
  const Mat3 R_NM = transpose(X_MN.R);  // RotationMatrix::inverse()
  std::vector<int> tet0_of_face, tet1_of_face;
  for (const auto& [tet0, tet1] : candidates) {
    const int n = CalcContactPolygonTetTet(mesh_M, field_M, mesh_N, field_N, X_MN, R_NM, tet0, tet1, &builder);
    for (int i = 0; i < n; ++i) {
      tet0_of_face.push_back(tet0);
      tet1_of_face.push_back(tet1);
    }
  }
  if (builder.num_faces() == 0) return false;
  *out = SurfaceResult{};
  MakeMeshAndFieldInWorld(std::move(builder), X_WM, out);
  const int nf = out->num_faces();
  out->grad_eM_W.reserve(nf);
  out->grad_eN_W.reserve(nf);
  for (int i = 0; i < nf; ++i) out->grad_eM_W.push_back(X_WM.R * field_M.gradients[tet0_of_face[i]]);
  for (int i = 0; i < nf; ++i) out->grad_eN_W.push_back(X_WN.R * field_N.gradients[tet1_of_face[i]]);
  out->has_grad_eM = out->has_grad_eN = true;
  out->id_M = id_M;
  out->id_N = id_N;
  out->elem_first = std::move(tet0_of_face);
  out->elem_second = std::move(tet1_of_face);
  EnforceIdOrder(out);
  return true;
}

// ---------------------------------------------------------------------------
// Soft-soft through the surfaces: the second IntersectFields overload (field_intersection.cc:254-375), which the
// reference builds and tests (field_intersection_test.cc:305-344,423-476) but does not call from
// ComputeContactSurfaceFromCompliantVolumes (use_surfaces = false, :511-517).

// VolumeMeshTopology::Initialize (volume_mesh_topology.cc:13-68): neighbor(e, f) = the tetrahedron sharing face f of e
// (face f is the one opposite vertex f), -1 on the boundary.
inline std::vector<std::array<int, 4>> VolumeMeshNeighbors(const VolumeMesh& mesh) {
  std::vector<std::array<int, 4>> neighbors(mesh.num_elements(), std::array<int, 4>{-1, -1, -1, -1});
  auto get_face = [&mesh](int e, int f) {
    std::array<int, 3> t{mesh.tets[e][(f + 1) % 4], mesh.tets[e][(f + 2) % 4], mesh.tets[e][(f + 3) % 4]};
    std::sort(t.begin(), t.end());
    return t;
  };
  std::map<std::array<int, 3>, int> face_to_tet;
  for (int tet = 0; tet < mesh.num_elements(); ++tet) {
    for (int f = 0; f < 4; ++f) {
      const std::array<int, 3> face = get_face(tet, f);
      auto it = face_to_tet.find(face);
      if (it == face_to_tet.end()) {
        face_to_tet.emplace(face, tet);
        continue;
      }
      const int other = it->second;
      neighbors[tet][f] = other;
      for (int g = 0; g < 4; ++g)
        if (get_face(other, g) == face) {
          neighbors[other][g] = tet;
          break;
        }
    }
  }
  return neighbors;
}

// CalcContactPolygon as the BFS sees it (field_intersection.cc:377-428): the intersected faces of the two tetrahedra
// (0-3: faces of tet0, 4-7: faces of tet1; one entry per vertex of the polygon BEFORE the last
// RemoveNearlyDuplicateVertices), empty when no polygon was added.
inline std::vector<int> CalcContactPolygonTetTetFaces(const VolumeMesh& mesh0_M, const VolumeField& field0_M,
                                                      const VolumeMesh& mesh1_N, const VolumeField& field1_N,
                                                      const Xform& X_MN, const Mat3& R_NM, int tet0, int tet1) {
  Plane equilibrium_plane_M;
  if (!CalcEquilibriumPlane(tet0, field0_M, tet1, field1_N, X_MN, &equilibrium_plane_M)) return {};
  const Vec3 polygon_nhat_M = equilibrium_plane_M.nhat;
  if (!IsPlaneNormalAlongPressureGradient(polygon_nhat_M, tet0, field0_M)) return {};
  const Vec3 reverse_polygon_nhat_N = R_NM * (-polygon_nhat_M);
  if (!IsPlaneNormalAlongPressureGradient(reverse_polygon_nhat_N, tet1, field1_N)) return {};
  const auto [polygon_vertices_M, faces] = IntersectTetrahedra(tet0, mesh0_M, tet1, mesh1_N, X_MN, equilibrium_plane_M);
  if (polygon_vertices_M.size() < 3) return {};
  return faces;
}

// The tetrahedron pairs the BFS hands to CalcContactPolygon, in queue order (field_intersection.cc:254-375):
// seeds = the tetrahedra behind every pair of surface triangles in overlapping leaves of the two surface BVHs,
// then the face neighbours of every pair that produced a polygon. `surface_pairs` = Bvh::Collide of the two surface
// hierarchies, (triangle of M's surface, triangle of N's surface) in traversal order.
inline std::vector<std::pair<int, int>> SurfaceBfsTetPairs(const VolumeMesh& mesh0_M, const VolumeField& field0_M,
                                                           const std::vector<int>& tri_to_tet_M,
                                                           const std::vector<std::array<int, 4>>& neighbors_M,
                                                           const VolumeMesh& mesh1_N, const VolumeField& field1_N,
                                                           const std::vector<int>& tri_to_tet_N,
                                                           const std::vector<std::array<int, 4>>& neighbors_N,
                                                           const Xform& X_MN,
                                                           const std::vector<std::pair<int, int>>& surface_pairs) {
  std::set<std::pair<int, int>> visited;
  std::vector<std::pair<int, int>> queue;  // never popped: `head` walks it, so it doubles as the visit order
  for (const auto& [tri0, tri1] : surface_pairs) {
    const std::pair<int, int> tet_pair(tri_to_tet_M[tri0], tri_to_tet_N[tri1]);
    if (visited.insert(tet_pair).second) queue.push_back(tet_pair);
  }
  const Mat3 R_NM = transpose(X_MN.R);
  for (size_t head = 0; head < queue.size(); ++head) {
    const auto [tetM, tetN] = queue[head];
    const std::vector<int> faces =
        CalcContactPolygonTetTetFaces(mesh0_M, field0_M, mesh1_N, field1_N, X_MN, R_NM, tetM, tetN);
    for (int face : faces) {
      std::pair<int, int> next;
      if (face < 4) {
        const int neighborM = neighbors_M[tetM][face];
        if (neighborM < 0) continue;
        next = {neighborM, tetN};
      } else {
        const int neighborN = neighbors_N[tetN][face - 4];
        if (neighborN < 0) continue;
        next = {tetM, neighborN};
      }
      if (visited.insert(next).second) queue.push_back(next);
    }
  }
  return queue;
}

}  // namespace ho
