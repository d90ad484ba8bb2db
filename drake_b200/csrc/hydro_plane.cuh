// Per-element arithmetic of the two half-space special cases of CalcRigidCompliant
// (hydroelastic_calculator.cc:25-41):
//   compliant volume mesh x rigid half space   (mesh_plane_intersection.cc:182-342)
//   compliant half space  x rigid surface mesh (mesh_half_space_intersection.cc:17-430)
// and the World-frame mesh / field quantities of their polygons (the reference builds these two
// surfaces in World directly: PolyMeshBuilder / TriMeshBuilder over X_WM * p, then the
// PolygonSurfaceMesh / TriangleSurfaceMesh constructors, with no TransformVertices step).
//
// Everything here is __host__ __device__ over plain records so that tests/cpp/plane_host_check.cc can
// run the very same expressions under g++ in the CPU test suite; the product only ever calls them
// from k_plane_* (hydro_kernels.cuh). Operation order follows the reference (hydro_math.cuh).
#pragma once
#include "hydro_math.cuh"
#include "hydro_types.h"

namespace dbh {

constexpr double kCosAlphaPlane = -0.38268343236508973;  // cos(5*pi/8) (contact_surface_utility.cc:347-353)

// One contact polygon of a half-space pair, measured and expressed in World.
struct WorldPoly {
  int n;         // 0 (none), 3 or 4
  int status;    // 0, or the dbh_status this element raises
  D3 v[4];
  double e[4];
  // Where each vertex comes from in the mesh, so that the host can weld the vertices the reference shares between
  // polygons (cut_edges, mesh_plane_intersection.cc:205-231; vertices_to_newly_created_vertices /
  // edges_to_newly_created_vertices, mesh_half_space_intersection.cc:52-101): i = a copy of the element's local vertex i;
  // 16 + 4 a + b = cut from the edge between the element's local vertices a and b.
  int origin[4];
  D3 nhat_W;     // the normal handed to the builder (TriMeshBuilder's centroid weights)
  D3 grad_W;     // grad_e_MN_W
};

// Rigid half space (z <= 0 of R) as a plane in the compliant mesh's frame S
// (mesh_plane_intersection.cc:304-311): X_SR = X_WS.inverse() * X_WR (rigid_transform.h:738-745 and
// ComposeXX), normal = Rz_S taken as already normalised, displacement = Rz_S . p_SRo (plane.cc:29).
__host__ __device__ __forceinline__ void rigid_half_space_in_mesh_frame(const Xf& X_WS, const Xf& X_WR, D3* nhat_S,
                                                                        double* displacement) {
  const D3 Rz_S = rot_t(X_WS, make_d3(X_WR.m[6], X_WR.m[7], X_WR.m[8]));
  const D3 p_SW = rot_t(X_WS, -make_d3(X_WS.m[9], X_WS.m[10], X_WS.m[11]));
  const D3 p_SRo = p_SW + rot_t(X_WS, make_d3(X_WR.m[9], X_WR.m[10], X_WR.m[11]));
  *nhat_S = Rz_S;
  *displacement = dot(Rz_S, p_SRo);
}

// kMarchingTetsEdgeTable (mesh_plane_intersection.cc:47-64): the up to four tetrahedron edges cut by the
// plane for a sign code (bit i set: vertex i strictly above), one edge index per nibble, 0xF = none.
__host__ __device__ __forceinline__ unsigned int marching_tets_row(int code) {
  const unsigned short kEdgeTable[16] = {0xFFFF, 0xF230, 0xF410, 0x1234, 0xF521, 0x1530, 0x4520, 0xF453,
                                         0xF543, 0x0254, 0x0351, 0xF251, 0x4321, 0xF140, 0xF320, 0xFFFF};
  return kEdgeTable[code];
}

// Only the size (0, 3 or 4) of the polygon slice_tet_with_plane would produce: the same heights and
// sign code, nothing else (the count pass needs no pose, pressure or gradient).
__host__ __device__ __forceinline__ int count_tet_slice(const TetRec& tet, D3 nhat_S, double displacement) {
  int code = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dot(nhat_S, load3(tet.v[i])) - displacement > 0) code |= 1 << i;
  }
  const unsigned int row = marching_tets_row(code);
  if ((row & 0xFu) == 0xFu) return 0;
  return ((row >> 12) & 0xFu) == 0xFu ? 3 : 4;
}

// SliceTetWithPlane (mesh_plane_intersection.cc:182-249) for one tetrahedron: marching-tetrahedra
// tables (:20-117, four edge indices per sign code, packed one per nibble, 0xF = none), cut point
// p_MV0 + t (p_MV1 - p_MV0) with t = d0 / (d0 - d1), pressure e0 + t (e1 - e0) from the VERTEX values,
// vertices re-expressed in World before they reach the builder. vp: the four vertex pressures.
__host__ __device__ __forceinline__ void slice_tet_with_plane(const TetRec& tet, const double* vp, D3 nhat_S,
                                                              double displacement, const Xf& X_WS, WorldPoly* out) {
  const unsigned int kEdgeFirst = 0x210210u;   // kTetEdges[k].first, one per nibble
  const unsigned int kEdgeSecond = 0x333021u;  // kTetEdges[k].second
  out->n = 0;
  out->status = 0;
  double distance[4];
  int code = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    distance[i] = dot(nhat_S, load3(tet.v[i])) - displacement;  // Plane::CalcHeight (plane.h:69-71)
    if (distance[i] > 0) code |= 1 << i;
  }
  const unsigned int row = marching_tets_row(code);
  if ((row & 0xFu) == 0xFu) return;
  int n = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned int edge = (row >> (4 * k)) & 0xFu;
    if (edge == 0xFu) break;
    const int a = (kEdgeFirst >> (4 * edge)) & 0xF, b = (kEdgeSecond >> (4 * edge)) & 0xF;
    const D3 p0 = load3(tet.v[a]), p1 = load3(tet.v[b]);
    const double t = distance[a] / (distance[a] - distance[b]);
    const D3 p_MC = p0 + t * (p1 - p0);
    out->v[n] = xform_point(X_WS, p_MC);
    out->e[n] = vp[a] + t * (vp[b] - vp[a]);
    out->origin[n] = 16 + 4 * a + b;
    ++n;
  }
  out->n = n;
  out->nhat_W = rot(X_WS, nhat_S);
  out->grad_W = rot(X_WS, load3(tet.grad));
}

// Compliant half space (z <= 0 of S, pushed out by its margin) as a PosedHalfSpace in the rigid
// mesh's frame R (mesh_half_space_intersection.cc:360-388): X_RS = X_WR.InvertAndCompose(X_WS),
// p_RS += Sz_R * margin, Plane(Sz_R, p_RS) normalises (plane.cc:16-29). Returns a dbh_status.
__host__ __device__ __forceinline__ int compliant_half_space_in_mesh_frame(const Xf& X_WR, const Xf& X_WS, double margin,
                                                                           D3* nhat_R, double* displacement) {
  const Xf X_RS = invert_and_compose(X_WR, X_WS);
  const D3 Sz_R = make_d3(X_RS.m[6], X_RS.m[7], X_RS.m[8]);
  const D3 p_RS = make_d3(X_RS.m[9] + Sz_R.x * margin, X_RS.m[10] + Sz_R.y * margin, X_RS.m[11] + Sz_R.z * margin);
  const double magnitude = norm(Sz_R);
  if (magnitude < 1e-10) return 9;  // DBH_ERR_PLANE_NORMAL
  *nhat_R = Sz_R / magnitude;
  *displacement = dot(*nhat_R, p_RS);
  return 0;
}

__host__ __device__ __forceinline__ int sgn_of(double x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }

// The BVH callback's filter (mesh_half_space_intersection.cc:349-358, IsFaceNormalInNormalDirection,
// contact_surface_utility.cc:336-354) followed by ConstructTriangleHalfspaceIntersectionPolygon
// (:128-262) for one triangle of the rigid mesh; pressure (-phi - margin) * pressure_scale (:392-396)
// evaluated at the mesh-frame point, positions X_WR * p (:84-86).
__host__ __device__ __forceinline__ void clip_tri_by_half_space(const TriRec& tri, D3 nhat_R, double displacement,
                                                                double pressure_scale, double margin, const Xf& X_WR,
                                                                const Xf& X_WS, WorldPoly* out,
                                                                bool filter_normals = true) {
  out->n = 0;
  out->status = 0;
  const D3 Sz_W = make_d3(X_WS.m[6], X_WS.m[7], X_WS.m[8]);
  const D3 face_normal_W = rot(X_WR, load3(tri.n));
  // filter_normals is only ever false in the host check of the reference's single-triangle cases, which call
  // ConstructTriangleHalfspaceIntersectionPolygon without the BVH callback.
  if (filter_normals && !(dot(-Sz_W, normalized(face_normal_W)) > kCosAlphaPlane)) return;
  D3 p[3];
  double s[3];
  int num_positive = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p[i] = load3(tri.v[i]);
    s[i] = dot(nhat_R, p[i]) - displacement;
    if (s[i] > 0) ++num_positive;
  }
  if (num_positive == 3) return;
  int n = 0;
  auto add = [&](D3 p_RV, int origin) {
    const double phi_V = dot(nhat_R, p_RV) - displacement;
    out->v[n] = xform_point(X_WR, p_RV);
    out->e[n] = (-phi_V - margin) * pressure_scale;
    out->origin[n] = origin;
    ++n;
  };
  // CalcEdgePlaneIntersection (:36-50); GetVertexAddIfNeeded demands different signs (:66).
  auto add_edge = [&](int a, int b) {
    if (sgn_of(s[a]) == sgn_of(s[b])) out->status = 13;  // DBH_ERR_REFERENCE_DEMAND
    const double t = fabs(s[a]) / (fabs(s[a]) + fabs(s[b]));
    add(p[a] + t * (p[b] - p[a]), 16 + 4 * a + b);
  };
  if (num_positive == 0) {
    add(p[0], 0);
    add(p[1], 1);
    add(p[2], 2);
  } else if (num_positive == 1) {
    int i0 = 0;  // the first vertex with s >= 0 (:188-189)
    while (i0 < 2 && !(s[i0] >= 0)) ++i0;
    const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    // Builder order: edge (i0,i1), edge (i0,i2), v1, v2; polygon {v1, v2, e02, e01} (:196-215).
    add(p[i1], i1);
    add(p[i2], i2);
    add_edge(i0, i2);
    add_edge(i0, i1);
  } else {
    int i0 = 0;  // the first vertex with s <= 0 (:226-227)
    while (i0 < 2 && !(s[i0] <= 0)) ++i0;
    const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    add(p[i0], i0);
    add_edge(i0, i1);
    add_edge(i0, i2);
  }
  out->n = n;
  out->nhat_W = face_normal_W;
  out->grad_W = (-pressure_scale) * Sz_W;  // grad_p_W (:398)
}

// Which half-space pair owns work item w (plane_elem_off[q] <= w < plane_elem_off[q + 1])? The pairs of a
// batch mostly have meshes of similar size, so an interpolated guess plus a short walk finds it in O(1)
// loads; a binary search (a chain of ~log2(pairs) dependent loads) is only the fallback.
__host__ __device__ __forceinline__ int find_plane_pair(const long long* off, int q_begin, int q_end, long long w, long long w0,
                                               long long count) {
  int q = q_begin + (int)(((w - w0) * (long long)(q_end - q_begin)) / count);  // in [q_begin, q_end - 1]
  int steps = 0;
  while (steps < 4 && off[q] > w) {
    --q;
    ++steps;
  }
  while (steps < 4 && off[q + 1] <= w) {
    ++q;
    ++steps;
  }
  if (off[q] <= w && w < off[q + 1]) return q;
  int lo = q_begin, hi = q_end;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= w) lo = mid;
    else hi = mid;
  }
  return lo;
}

// What one polygon contributes to the packed results.
struct PackedPoly {
  int nv;           // vertices: n (polygons) or n + 1 (triangle fan about the centroid)
  int nf;           // faces: 1 or n
  int nfd;          // face_data ints: n + 1 or 3 n
  D3 v[5];
  double e[5];
  double area[4];
  D3 normal[4];     // out of N into M for the (S, R) construction order; the caller negates when swapped
  D3 centroid[4];
  double e_c[4];
  int fd[12];       // vertex indices local to the surface
};

// Mesh / field quantities of one World-frame polygon. vloc: surface-local index of its first vertex.
// Polygons: PolyMeshBuilder::AddPolygon (contact_surface_utility.cc:186-205), PolygonSurfaceMesh
// CalcAreaNormalAndCentroid (polygon_surface_mesh.cc:117-218), MeshFieldLinear value at the origin and
// EvaluateCartesian (mesh_field_linear.h:349,502-509), ReverseFaceWinding (polygon_surface_mesh.cc:39-49).
// Triangles: TriMeshBuilder::AddPolygon + CalcPolygonCentroid + AddPolygonToTriangleMeshData
// (contact_surface_utility.cc:137-164,218-316), TriangleSurfaceMesh derived quantities
// (triangle_surface_mesh.h:452-487), Evaluate at (1/3,1/3,1/3) (mesh_field_linear.h:300-312),
// SurfaceTriangle::ReverseWinding (triangle_surface_mesh.h:53).
__host__ __device__ __forceinline__ void pack_world_polygon(const WorldPoly& P, int representation, bool swapped,
                                                            int vloc, PackedPoly* out) {
  const int n = P.n;
  const D3 v0 = P.v[0];
  const double e0 = P.e[0];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < n) {
      out->v[k] = P.v[k];
      out->e[k] = P.e[k];
    }
  }
  if (representation == 1 /*DBH_POLYGON*/) {
    out->nv = n;
    out->nf = 1;
    out->nfd = n + 1;
    out->fd[0] = n;
    D3 p_scaled = make_d3(0, 0, 0), normal = make_d3(0, 0, 0), vsum = v0;
    double double_area = 0;
    out->fd[1] = vloc + (swapped ? (n - 1) : 0);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      if (k < n) {
        out->fd[1 + k] = vloc + (swapped ? (n - 1 - k) : k);
        vsum = vsum + P.v[k];
        if (k >= 2) {  // triangle (0, k-1, k)
          const D3 cr = cross(P.v[k - 1] - v0, P.v[k] - v0);
          normal = normal + cr;
          const double mag = norm(cr);
          double_area += mag;
          p_scaled = p_scaled + mag * ((v0 + P.v[k - 1]) + P.v[k]);
        }
      }
    }
    const double area = double_area / 2;
    const D3 p_as = p_scaled / 6;
    const D3 c = area != 0.0 ? p_as / area : vsum / (double)n;
    out->area[0] = area;
    out->normal[0] = normalized(normal);
    out->centroid[0] = c;
    const double at_Wo = e0 - dot(P.grad_W, v0);
    out->e_c[0] = dot(P.grad_W, c) + at_Wo;
    return;
  }
  out->nv = n + 1;
  out->nf = n;
  out->nfd = 3 * n;
  D3 accum = make_d3(0, 0, 0), vsum = v0;
  double total_weight = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    if (k < n) {
      vsum = vsum + P.v[k];
      if (k >= 2) {
        const double weight = dot(cross(P.v[k - 1] - v0, P.v[k] - v0), P.nhat_W);
        accum = accum + weight * (((v0 + P.v[k - 1]) + P.v[k]) / 3.0);
        total_weight += weight;
      }
    }
  }
  D3 p_c;
  if (n == 3) {
    p_c = vsum / 3.0;
  } else {
    if (total_weight == 0) {
      accum = vsum;
      total_weight = n;
    }
    p_c = accum / total_weight;
  }
  const double e_cv = dot(P.grad_W, p_c - v0) + e0;
  out->v[n] = p_c;
  out->e[n] = e_cv;
  const double third = 1.0 / 3.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < n) {
      const int k1 = k == 0 ? n - 1 : k - 1;  // triangle (v_{k-1}, v_k, centroid)
      const D3 A = P.v[k1], B = P.v[k];
      const D3 cr = cross(B - A, p_c - A);
      const double nm = norm(cr);
      out->area[k] = 0.5 * nm;
      out->normal[k] = (nm != 0.0) ? cr / nm : cr;
      out->centroid[k] = ((A + B) + p_c) / 3.0;
      double e_c = third * P.e[k1];
      e_c += third * P.e[k];
      e_c += third * e_cv;
      out->e_c[k] = e_c;
      out->fd[3 * k + 0] = vloc + (swapped ? k : k1);
      out->fd[3 * k + 1] = vloc + (swapped ? k1 : k);
      out->fd[3 * k + 2] = vloc + n;
    }
  }
}

}  // namespace dbh
