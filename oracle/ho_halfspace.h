// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// Restatement of the two half-space special cases of CalcRigidCompliant
// (hydroelastic_calculator.cc:25-41):
//   compliant volume mesh x rigid half space   mesh_plane_intersection.cc:182-342
//   compliant half space  x rigid surface mesh mesh_half_space_intersection.cc:17-430
// Both build their mesh in World directly and weld vertices that come from the same mesh edge /
// mesh vertex through hash maps, like the reference. The element order is the mesh's index order
// (the reference visits elements in the order its BVH reports them; the emitted face set does not
// depend on that order, a welded vertex's last bits do).
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "ho_contact.h"

namespace ho {

// SliceTetWithPlane (mesh_plane_intersection.cc:182-249). Returns the number of faces added.
inline int SliceTetWithPlane(int tet_index, const VolumeMesh& mesh_M, const VolumeField& field_M, const Plane& plane_M,
                             const Xform& X_WM, MeshBuilder* builder_W, std::map<std::pair<int, int>, int>* cut_edges) {
  static const int kTetEdges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  static const int kEdgeTable[16][4] = {{-1, -1, -1, -1}, {0, 3, 2, -1}, {0, 1, 4, -1}, {4, 3, 2, 1},
                                        {1, 2, 5, -1},    {0, 3, 5, 1},  {0, 2, 5, 4},  {3, 5, 4, -1},
                                        {3, 4, 5, -1},    {4, 5, 2, 0},  {1, 5, 3, 0},  {1, 5, 2, -1},
                                        {1, 2, 3, 4},     {0, 4, 1, -1}, {0, 2, 3, -1}, {-1, -1, -1, -1}};
  double distance[4];
  int code = 0;
  for (int i = 0; i < 4; ++i) {
    distance[i] = plane_M.CalcHeight(mesh_M.verts[mesh_M.tets[tet_index][i]]);
    if (distance[i] > 0) code |= 1 << i;
  }
  const int* intersected_edges = kEdgeTable[code];
  if (intersected_edges[0] == -1) return 0;
  std::vector<int> face_verts;
  for (int e = 0; e < 4; ++e) {
    const int edge_index = intersected_edges[e];
    if (edge_index == -1) break;
    const int a = kTetEdges[edge_index][0], b = kTetEdges[edge_index][1];
    const int v0 = mesh_M.tets[tet_index][a];
    const int v1 = mesh_M.tets[tet_index][b];
    const std::pair<int, int> mesh_edge{std::min(v0, v1), std::max(v0, v1)};  // SortedPair<int>
    auto iter = cut_edges->find(mesh_edge);
    if (iter != cut_edges->end()) {
      face_verts.push_back(iter->second);
    } else {
      const Vec3& p_MV0 = mesh_M.verts[v0];
      const Vec3& p_MV1 = mesh_M.verts[v1];
      const double d_v0 = distance[a];
      const double d_v1 = distance[b];
      const double t = d_v0 / (d_v0 - d_v1);
      const Vec3 p_MC = p_MV0 + t * (p_MV1 - p_MV0);
      const double e0 = field_M.values[v0];
      const double e1 = field_M.values[v1];
      const int new_index = builder_W->AddVertex(X_WM * p_MC, e0 + t * (e1 - e0));
      (*cut_edges)[mesh_edge] = new_index;
      face_verts.push_back(new_index);
    }
  }
  const Vec3 nhat_W = X_WM.R * plane_M.nhat;
  const Vec3 grad_e_MN_W = X_WM.R * field_M.gradients[tet_index];
  return builder_W->AddPolygon(face_verts, nhat_W, grad_e_MN_W);
}

// ComputeContactSurfaceFromCompliantVolumeRigidHalfSpace (mesh_plane_intersection.cc:251-342).
// Returns false when the surface is null.
inline bool ContactSurfaceFromCompliantVolumeRigidHalfSpace(int64_t id_S, const VolumeMesh& mesh_S,
                                                            const VolumeField& field_S, const Xform& X_WS, int64_t id_R,
                                                            const Xform& X_WR, Representation rep, SurfaceResult* out) {
  const Xform X_SR = Inverse(X_WS) * X_WR;
  const Vec3 Rz_S{X_SR.R(0, 2), X_SR.R(1, 2), X_SR.R(2, 2)};
  const Plane plane_S(Rz_S, X_SR.p, /*already_normalized=*/true);
  MeshBuilder builder_W(rep);
  std::map<std::pair<int, int>, int> cut_edges;
  std::vector<Vec3> grad_eM_W;
  std::vector<int> e0;
  for (int tet = 0; tet < mesh_S.num_elements(); ++tet) {
    const int n = SliceTetWithPlane(tet, mesh_S, field_S, plane_S, X_WS, &builder_W, &cut_edges);
    const Vec3 grad_eMi_W = X_WS.R * field_S.gradients[tet];
    for (int i = 0; i < n; ++i) {
      grad_eM_W.push_back(grad_eMi_W);
      e0.push_back(tet);
    }
  }
  if (builder_W.num_faces() == 0) return false;
  *out = SurfaceResult{};
  MakeMeshAndFieldInWorld(std::move(builder_W), Xform{}, out, /*already_in_world=*/true);
  out->grad_eM_W = std::move(grad_eM_W);
  out->has_grad_eM = true;
  out->has_grad_eN = false;
  out->id_M = id_S;
  out->id_N = id_R;
  out->elem_second.assign(e0.size(), -1);  // the half space has no elements
  out->elem_first = std::move(e0);
  EnforceIdOrder(out);
  return true;
}

// ConstructTriangleHalfspaceIntersectionPolygon (mesh_half_space_intersection.cc:128-262) with its
// two GetVertexAddIfNeeded helpers (:52-124) and CalcEdgePlaneIntersection (:36-50).
struct HalfSpaceClipper {
  const TriMesh& mesh_F;
  const PosedHalfSpace& half_space_F;
  double pressure_scale;
  double margin;
  const Xform& X_WF;
  MeshBuilder* builder_W;
  std::map<int, int> vertices_to_new;
  std::map<std::pair<int, int>, int> edges_to_new;

  static int sgn(double x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
  double pressure_in_F(const Vec3& p_FV) const {  // calc_pressure_R (:392-396)
    const double phi_V = half_space_F.CalcHeight(p_FV);
    return (-phi_V - margin) * pressure_scale;
  }
  int EdgeVertex(int a, int b, double s_a, double s_b) {
    if (sgn(s_a) == sgn(s_b)) throw std::logic_error("GetVertexAddIfNeeded: DRAKE_DEMAND(sgn(s_a) != sgn(s_b))");
    const std::pair<int, int> edge{std::min(a, b), std::max(a, b)};
    auto it = edges_to_new.find(edge);
    if (it == edges_to_new.end()) {
      const double t = std::fabs(s_a) / (std::fabs(s_a) + std::fabs(s_b));
      const Vec3 p_FV = mesh_F.verts[a] + t * (mesh_F.verts[b] - mesh_F.verts[a]);
      const int index = builder_W->AddVertex(X_WF * p_FV, pressure_in_F(p_FV));
      it = edges_to_new.insert({edge, index}).first;
    }
    return it->second;
  }
  int MeshVertex(int v_index) {
    auto it = vertices_to_new.find(v_index);
    if (it == vertices_to_new.end()) {
      const Vec3 p_FV = mesh_F.verts[v_index];
      const int index = builder_W->AddVertex(X_WF * p_FV, pressure_in_F(p_FV));
      it = vertices_to_new.insert({v_index, index}).first;
    }
    return it->second;
  }
  // Returns the number of faces added.
  int Clip(int tri_index, const Vec3& grad_pressure_in_W) {
    const Tri& triangle = mesh_F.tris[tri_index];
    double s[3];
    int num_positive = 0;
    for (int i = 0; i < 3; ++i) {
      s[i] = half_space_F.CalcHeight(mesh_F.verts[triangle[i]]);
      if (s[i] > 0) ++num_positive;
    }
    if (num_positive == 3) return 0;
    const Vec3 nhat_W = X_WF.R * mesh_F.normals[tri_index];
    if (num_positive == 0) {
      const int v0 = MeshVertex(triangle[0]);
      const int v1 = MeshVertex(triangle[1]);
      const int v2 = MeshVertex(triangle[2]);
      return builder_W->AddPolygon({v0, v1, v2}, nhat_W, grad_pressure_in_W);
    }
    if (num_positive == 1) {
      for (int i0 = 0; i0 < 3; ++i0) {
        if (s[i0] >= 0) {
          const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
          const int v0 = triangle[i0], v1 = triangle[i1], v2 = triangle[i2];
          const int e01 = EdgeVertex(v0, v1, s[i0], s[i1]);
          const int e02 = EdgeVertex(v0, v2, s[i0], s[i2]);
          const int n1 = MeshVertex(v1);
          const int n2 = MeshVertex(v2);
          return builder_W->AddPolygon({n1, n2, e02, e01}, nhat_W, grad_pressure_in_W);
        }
      }
    }
    for (int i0 = 0; i0 < 3; ++i0) {  // num_positive == 2
      if (s[i0] <= 0) {
        const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
        const int v0 = triangle[i0], v1 = triangle[i1], v2 = triangle[i2];
        const int n0 = MeshVertex(v0);
        const int e01 = EdgeVertex(v0, v1, s[i0], s[i1]);
        const int e02 = EdgeVertex(v0, v2, s[i0], s[i2]);
        return builder_W->AddPolygon({n0, e01, e02}, nhat_W, grad_pressure_in_W);
      }
    }
    throw std::logic_error("ConstructTriangleHalfspaceIntersectionPolygon: unreachable");
  }
};

// ComputeContactSurfaceFromCompliantHalfSpaceRigidMesh (mesh_half_space_intersection.cc:338-418).
// Returns false when the surface is null.
inline bool ContactSurfaceFromCompliantHalfSpaceRigidMesh(int64_t id_S, const Xform& X_WS, double pressure_scale,
                                                          int64_t id_R, const TriMesh& mesh_R, const Xform& X_WR,
                                                          Representation rep, double margin, SurfaceResult* out,
                                                          bool filter_normals = true) {
  const Vec3 unit_grad_p_W = -Vec3{X_WS.R(0, 2), X_WS.R(1, 2), X_WS.R(2, 2)};
  Xform X_RS = InvertAndCompose(X_WR, X_WS);
  const Vec3 Sz_R{X_RS.R(0, 2), X_RS.R(1, 2), X_RS.R(2, 2)};
  X_RS.p = X_RS.p + Sz_R * margin;
  const PosedHalfSpace hs_R(Sz_R, X_RS.p, /*already_normalized=*/false);
  const Vec3 grad_p_W = (-pressure_scale) * Vec3{X_WS.R(0, 2), X_WS.R(1, 2), X_WS.R(2, 2)};
  MeshBuilder builder_W(rep);
  HalfSpaceClipper clipper{mesh_R, hs_R, pressure_scale, margin, X_WR, &builder_W, {}, {}};
  std::vector<int> e1;
  for (int tri = 0; tri < mesh_R.num_elements(); ++tri) {
    // The BVH callback's filter (:349-358); the BVH itself only drops triangles with no vertex inside.
    // filter_normals = false: ComputeContactSurface (:264-336) over every triangle, as the reference's
    // MeshHalfSpaceValueTest cases call it.
    if (filter_normals && !IsFaceNormalInNormalDirection(unit_grad_p_W, mesh_R, tri, X_WR.R)) continue;
    const int n = clipper.Clip(tri, grad_p_W);
    for (int i = 0; i < n; ++i) e1.push_back(tri);
  }
  if (builder_W.num_faces() == 0) return false;
  *out = SurfaceResult{};
  MakeMeshAndFieldInWorld(std::move(builder_W), Xform{}, out, /*already_in_world=*/true);
  out->grad_eM_W.assign(out->num_faces(), grad_p_W);  // grad_eS_W: M is the half space (:322-329)
  out->has_grad_eM = true;
  out->has_grad_eN = false;
  out->id_M = id_S;
  out->id_N = id_R;
  out->elem_first.assign(e1.size(), -1);
  out->elem_second = std::move(e1);
  EnforceIdOrder(out);
  return true;
}

}  // namespace ho
