// TEST INFRASTRUCTURE ONLY. Compiles the __host__ __device__ per-element arithmetic of the half-space
// kernels (drake_b200/csrc/hydro_plane.cuh) with g++ so that tests/test_plane_math_cpu.py can compare the
// very expressions k_plane runs on the device with the oracle, on a machine without a GPU. This is
// not a CPU path of the product: nothing under drake_b200/ builds, loads or calls this file, and the
// C-ABI library has no entry point that reaches these functions without a CUDA device.
#include <math.h>

#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __forceinline__ inline
#include "../../drake_b200/csrc/hydro_clip.cuh"
#include "../../drake_b200/csrc/hydro_plane.cuh"
#include "../../drake_b200/csrc/hydro_wrench.cuh"
#include "../../drake_b200/csrc/hydro_derive.cuh"
#include "../../drake_b200/csrc/hydro_bfs.cuh"

using namespace dbh;

namespace {

struct Out {
  int* elem;
  int* face_n;       // polygon size behind each face
  double* areas;
  double* normals;
  double* centroids;
  double* e_c;
  double* grad;
  double* verts;     // 3 per vertex
  double* pressures;
  int* face_data;
  int nf, nv, nfd, npoly;
};

void Append(const WorldPoly& P, int rep, bool swapped, int elem, Out* o) {
  PackedPoly K;
  // Surface-local index of the polygon's first vertex (k_plane: vj for polygons, vj + pj for fans).
  pack_world_polygon(P, rep, swapped, o->nv, &K);
  for (int k = 0; k < K.nv; ++k) {
    o->verts[3 * (o->nv + k) + 0] = K.v[k].x;
    o->verts[3 * (o->nv + k) + 1] = K.v[k].y;
    o->verts[3 * (o->nv + k) + 2] = K.v[k].z;
    o->pressures[o->nv + k] = K.e[k];
  }
  for (int f = 0; f < K.nf; ++f) {
    const int i = o->nf + f;
    o->elem[i] = elem;
    o->face_n[i] = P.n;
    o->areas[i] = K.area[f];
    const D3 nn = swapped ? -K.normal[f] : K.normal[f];
    o->normals[3 * i] = nn.x, o->normals[3 * i + 1] = nn.y, o->normals[3 * i + 2] = nn.z;
    o->centroids[3 * i] = K.centroid[f].x, o->centroids[3 * i + 1] = K.centroid[f].y, o->centroids[3 * i + 2] = K.centroid[f].z;
    o->e_c[i] = K.e_c[f];
    o->grad[3 * i] = P.grad_W.x, o->grad[3 * i + 1] = P.grad_W.y, o->grad[3 * i + 2] = P.grad_W.z;
  }
  for (int k = 0; k < K.nfd; ++k) o->face_data[o->nfd + k] = K.fd[k];
  o->nf += K.nf;
  o->nv += K.nv;
  o->nfd += K.nfd;
  ++o->npoly;
}

}  // namespace

extern "C" {

// Compliant mesh x rigid half space over a whole mesh. tet_v [nt*12], tet_grad [nt*3], tet_vp [nt*4].
// counts4 = faces, vertices, face_data ints, polygons. Returns 0 or the first status raised.
int phc_mesh_vs_rigid_half_space(int nt, const double* tet_v, const double* tet_grad, const double* tet_vp,
                                 const double* X_WS, const double* X_WR, int rep, int swapped, int* elem, int* face_n,
                                 double* areas, double* normals, double* centroids, double* e_c, double* grad,
                                 double* verts, double* pressures, int* face_data, int* counts4) {
  Xf XS, XR;
  std::memcpy(XS.m, X_WS, sizeof(XS.m));
  std::memcpy(XR.m, X_WR, sizeof(XR.m));
  D3 nhat_S;
  double disp;
  rigid_half_space_in_mesh_frame(XS, XR, &nhat_S, &disp);
  Out o{elem, face_n, areas, normals, centroids, e_c, grad, verts, pressures, face_data, 0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) {
    TetRec rec;
    std::memcpy(rec.v, tet_v + 12 * t, sizeof(rec.v));
    std::memcpy(rec.grad, tet_grad + 3 * t, sizeof(rec.grad));
    rec.p_mo = 0;
    WorldPoly P;
    slice_tet_with_plane(rec, tet_vp + 4 * t, nhat_S, disp, XS, &P);
    if (P.status) return P.status;
    if (count_tet_slice(rec, nhat_S, disp) != P.n) return -1;  // the count pass's shortcut must agree
    if (P.n) Append(P, rep, swapped != 0, t, &o);
  }
  counts4[0] = o.nf, counts4[1] = o.nv, counts4[2] = o.nfd, counts4[3] = o.npoly;
  return 0;
}

// Compliant half space x rigid mesh over a whole mesh. tri_v [nt*9], tri_n [nt*3].
int phc_half_space_vs_rigid_mesh(int nt, const double* tri_v, const double* tri_n, const double* X_WS,
                                 double pressure_scale, double margin, const double* X_WR, int rep, int swapped,
                                 int filter_normals, int* elem, int* face_n, double* areas, double* normals, double* centroids, double* e_c,
                                 double* grad, double* verts, double* pressures, int* face_data, int* counts4) {
  Xf XS, XR;
  std::memcpy(XS.m, X_WS, sizeof(XS.m));
  std::memcpy(XR.m, X_WR, sizeof(XR.m));
  D3 nhat_R;
  double disp;
  if (int st = compliant_half_space_in_mesh_frame(XR, XS, margin, &nhat_R, &disp)) return st;
  Out o{elem, face_n, areas, normals, centroids, e_c, grad, verts, pressures, face_data, 0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) {
    TriRec rec;
    std::memcpy(rec.v, tri_v + 9 * t, sizeof(rec.v));
    std::memcpy(rec.n, tri_n + 3 * t, sizeof(rec.n));
    WorldPoly P;
    clip_tri_by_half_space(rec, nhat_R, disp, pressure_scale, margin, XR, XS, &P, filter_normals != 0);
    if (P.status) return P.status;
    if (P.n) Append(P, rep, swapped != 0, t, &o);
  }
  counts4[0] = o.nf, counts4[1] = o.nv, counts4[2] = o.nfd, counts4[3] = o.npoly;
  return 0;
}

// The work item -> pair map of k_plane (interpolated guess + short walk + binary-search fallback) against a
// plain linear scan, for every work item of the range [q_begin, q_end). Returns the number of mismatches.
int phc_check_find_plane_pair(int num_pairs, const long long* off, int q_begin, int q_end) {
  (void)num_pairs;
  const long long w0 = off[q_begin], count = off[q_end] - w0;
  int bad = 0, q_lin = q_begin;
  for (long long w = w0; w < w0 + count; ++w) {
    while (off[q_lin + 1] <= w) ++q_lin;
    if (find_plane_pair(off, q_begin, q_end, w, w0, count) != q_lin) ++bad;
  }
  return bad;
}

// The per-face traction arithmetic of k_wrench summed face by face in order (the oracle's order, so the
// comparison can be exact): rep 1: polygons (areas, normals, centroids, e_c per face); rep 0: triangles
// (vertices [nv*3], pressures [nv], tris [nf*3] besides areas and normals). bodies22 = p_WAo, p_WBo, w_WA,
// v_WA, w_WB, v_WB, p_WC (7 x 3) then dissipation. out18 = F_Ac_W, F_Ao_W, F_Bo_W as (tau, f).
void phc_surface_wrench(int rep, int nf, const double* areas, const double* normals, const double* centroids,
                        const double* e_c, const double* verts, const double* pressures, const int* tris,
                        const double* bodies22, double mu, double vslip, double* out18) {
  TractionBodies d;
  const double* b = bodies22;
  d.p_WAo = load3(b), d.p_WBo = load3(b + 3), d.w_WA = load3(b + 6), d.v_WA = load3(b + 9), d.w_WB = load3(b + 12);
  d.v_WB = load3(b + 15), d.p_WC = load3(b + 18);
  d.dissipation = b[21];
  d.mu_coulomb = mu;
  d.vslip_regularizer = vslip;
  Spatial6 F{make_d3(0, 0, 0), make_d3(0, 0, 0)};
  for (int f = 0; f < nf; ++f) {
    if (rep == 1) {
      F = F + polygon_face_force(d, areas[f], load3(normals + 3 * f), load3(centroids + 3 * f), e_c[f]);
    } else {
      const int* t = tris + 3 * f;
      F = F + triangle_face_force(d, areas[f], load3(normals + 3 * f), load3(verts + 3 * t[0]), load3(verts + 3 * t[1]),
                                  load3(verts + 3 * t[2]), pressures[t[0]], pressures[t[1]], pressures[t[2]]);
    }
  }
  Spatial6 Fa, Fb;
  shift_to_body_origins(d, F, &Fa, &Fb);
  const Spatial6* all[3] = {&F, &Fa, &Fb};
  for (int k = 0; k < 3; ++k) {
    out18[6 * k + 0] = all[k]->rot.x, out18[6 * k + 1] = all[k]->rot.y, out18[6 * k + 2] = all[k]->rot.z;
    out18[6 * k + 3] = all[k]->trans.x, out18[6 * k + 4] = all[k]->trans.y, out18[6 * k + 5] = all[k]->trans.z;
  }
}
// deformable_face_contact (hydro_wrench.cuh) for nf faces: per face normal, centroid, pressure, rotated gradient and the
// deformable triangle's three vertices (9 doubles); 22 doubles out per face, the layout of k_deformable.
void phc_deformable_contact(int nf, const double* mesh_normals, const double* centroids, const double* pressures,
                            const double* grads, const double* X_WR, const double* X_WD, const double* tri_verts, double* out) {
  Xf XR, XD;
  std::memcpy(XR.m, X_WR, sizeof(XR.m));
  std::memcpy(XD.m, X_WD, sizeof(XD.m));
  for (int f = 0; f < nf; ++f) {
    const DeformableFaceContact c =
        deformable_face_contact(load3(mesh_normals + 3 * f), load3(centroids + 3 * f), pressures[f], load3(grads + 3 * f), XR, XD,
                                load3(tri_verts + 9 * f), load3(tri_verts + 9 * f + 3), load3(tri_verts + 9 * f + 6));
    double* o = out + 22 * f;
    o[0] = c.p_WC.x, o[1] = c.p_WC.y, o[2] = c.p_WC.z;
    o[3] = c.nhat_W.x, o[4] = c.nhat_W.y, o[5] = c.nhat_W.z;
    for (int k = 0; k < 9; ++k) o[6 + k] = c.R_WC[k];
    o[15] = c.pressure;
    o[16] = c.grad_W.x, o[17] = c.grad_W.y, o[18] = c.grad_W.z;
    o[19] = c.barycentric.x, o[20] = c.barycentric.y, o[21] = c.barycentric.z;
  }
}
// deformable_pair_contact (hydro_wrench.cuh) for nf faces of a tet-tet surface: per face normal, centroid, field value
// and the two tetrahedra's vertices (12 doubles each); 24 doubles out per face, the layout of k_deformable.
void phc_deformable_pair_contact(int nf, const double* normals, const double* centroids, const double* values,
                                 const double* X_WM, const double* X_WN, const double* tet0_verts, const double* tet1_verts,
                                 double* out) {
  Xf XM, XN;
  std::memcpy(XM.m, X_WM, sizeof(XM.m));
  std::memcpy(XN.m, X_WN, sizeof(XN.m));
  for (int f = 0; f < nf; ++f) {
    double t0[4][3], t1[4][3];
    std::memcpy(t0, tet0_verts + 12 * f, sizeof(t0));
    std::memcpy(t1, tet1_verts + 12 * f, sizeof(t1));
    const DeformablePairContact c = deformable_pair_contact(load3(normals + 3 * f), load3(centroids + 3 * f), values[f], XM, XN, t0, t1);
    double* o = out + 24 * f;
    o[0] = c.p_WC.x, o[1] = c.p_WC.y, o[2] = c.p_WC.z;
    o[3] = c.nhat_W.x, o[4] = c.nhat_W.y, o[5] = c.nhat_W.z;
    for (int k = 0; k < 9; ++k) o[6 + k] = c.R_WC[k];
    o[15] = c.signed_distance;
    for (int k = 0; k < 4; ++k) o[16 + k] = c.bary0[k], o[20 + k] = c.bary1[k];
  }
}
double phc_combined_dissipation(double kA, double kB, double dA, double dB) {
  return combined_hunt_crossley_dissipation(kA, kB, dA, dB);
}
double phc_combined_friction(double a, double b) { return combined_dynamic_friction(a, b); }

// ---- the exact stage's polygon primitives (hydro_clip.cuh) on the exceptional path's storage ------------------
// PosedHalfSpace(normal, point) = Plane ctor (plane.cc:16-29): normalise, displacement = nhat . point.
static bool MakeHalfSpace(D3 normal, D3 point, D3* nh, double* d) {
  const double mag = norm(normal);
  if (mag < 1e-10) return false;
  *nh = normal / mag;
  *d = dot(*nh, point);
  return true;
}
// ClipPolygonByHalfSpace (mesh_intersection.cc:76-119). Returns the vertex count, -1 for a bad normal.
int phc_clip_polygon_by_half_space(int n, const double* verts, const double* normal, const double* point, double* out) {
  LocalPoly P;
  for (int i = 0; i < n; ++i) P.set(0, i, load3(verts + 3 * i));
  D3 nh;
  double d;
  if (!MakeHalfSpace(load3(normal), load3(point), &nh, &d)) return -1;
  const int m = n > 0 ? clip_by_half_space(P, 0, n, nh, d) : 0;
  for (int i = 0; i < m && i < LocalPoly::kCap; ++i) {
    const D3 v = P.get(1, i);
    out[3 * i] = v.x, out[3 * i + 1] = v.y, out[3 * i + 2] = v.z;
  }
  return m;
}
// RemoveNearlyDuplicateVertices (mesh_intersection.cc:121-170).
int phc_remove_nearly_duplicate(int n, const double* verts, double* out) {
  LocalPoly P;
  for (int i = 0; i < n; ++i) P.set(0, i, load3(verts + 3 * i));
  const int m = remove_nearly_duplicate(P, 0, n);
  for (int i = 0; i < m; ++i) {
    const D3 v = P.get(0, i);
    out[3 * i] = v.x, out[3 * i + 1] = v.y, out[3 * i + 2] = v.z;
  }
  return m;
}
// IntersectTetrahedra (field_intersection.cc:85-198) for a given equilibrium plane, with the sequence
// tet_stage_b runs: slice tet0 (marching tetrahedra), dedupe, clip by the four faces of tet1 (outward normal =
// R_MN * -inward, through vertex (face + 1) % 4; PosedHalfSpace normalises), dedupe. tet1_inward: 4 x 3 unit
// inward normals (VolumeMesh::inward_normal). Returns the vertex count (0: no polygon).
int phc_intersect_tetrahedra(const double* tet0_v, const double* tet1_v, const double* tet1_inward, const double* X_MN,
                             const double* nhat3, double displacement, double* out) {
  Xf X;
  std::memcpy(X.m, X_MN, sizeof(X.m));
  const D3 nhat_M = load3(nhat3);
  LocalPoly P;
  double dist[4];
  int code = 0;
  for (int k = 0; k < 4; ++k) {
    dist[k] = dot(nhat_M, load3(tet0_v + 3 * k)) - displacement;
    if (dist[k] > 0) code |= 1 << k;
  }
  const unsigned row = marching_tets_row(code);
  int n = 0;
  for (int k = 0; k < 4; ++k) {
    const unsigned edge = (row >> (4 * k)) & 0xFu;
    if (edge == 0xFu) break;
    const int e0 = (0x210210u >> (4 * edge)) & 0xF, e1 = (0x333021u >> (4 * edge)) & 0xF;
    const D3 V0 = load3(tet0_v + 3 * e0), V1 = load3(tet0_v + 3 * e1);
    const double t = dist[e0] / (dist[e0] - dist[e1]);
    P.set(0, n++, V0 + t * (V1 - V0));
  }
  n = remove_nearly_duplicate(P, 0, n);
  if (n < 3) return 0;
  int buf = 0;
  for (int face = 0; face < 4; ++face) {
    const D3 p_MA = xform_point(X, load3(tet1_v + 3 * ((face + 1) % 4)));
    D3 nh;
    double d;
    if (!MakeHalfSpace(rot(X, -load3(tet1_inward + 3 * face)), p_MA, &nh, &d)) return -1;
    n = clip_by_half_space_edges(P, buf, n, nh, d);
    if (n > LocalPoly::kCap) return -2;
    buf = 1 - buf;
    if (n == 0) return 0;
  }
  n = remove_nearly_duplicate(P, buf, n);
  if (n < 3) return 0;
  for (int i = 0; i < n; ++i) {
    const D3 v = P.get(buf, i);
    out[3 * i] = v.x, out[3 * i + 1] = v.y, out[3 * i + 2] = v.z;
  }
  return n;
}
// CalcContactPolygon as the surface-seeded intersector's k_bfs_wave runs it (hydro_bfs.cuh): the records registration
// derives (derive_tet), then tet_pair_polygon_with_faces on the exceptional path's storage. tet*_v: 4 x 3 vertices,
// tet*_p: the 4 vertex field values. out_faces: IntersectTetrahedra's returned face codes (one per vertex before the last
// dedupe); info2 = {number of codes, bit mask of the codes}. Returns the polygon size, or -status.
int phc_tet_pair_polygon_faces(const double* tet0_v, const double* tet0_p, const double* tet1_v, const double* tet1_p,
                               const double* X_MN, double* out_poly, int* out_faces, int* info2) {
  Xf X;
  std::memcpy(X.m, X_MN, sizeof(X.m));
  TetRec r0, r1;
  TetAux a0, a1;
  D3 v0[4], v1[4];
  for (int k = 0; k < 4; ++k) {
    v0[k] = load3(tet0_v + 3 * k);
    v1[k] = load3(tet1_v + 3 * k);
  }
  if (!derive_tet(v0, tet0_p, &r0, &a0) || !derive_tet(v1, tet1_p, &r1, &a1)) return -100;
  LocalPoly P;
  TetPairPolygon r;
  tet_pair_polygon_with_faces(X, r0, r1, a0, a1, P, &r);
  if (r.status) return -r.status;
  for (int i = 0; i < r.n; ++i) {
    const D3 v = P.get(r.buf, i);
    out_poly[3 * i] = v.x, out_poly[3 * i + 1] = v.y, out_poly[3 * i + 2] = v.z;
  }
  for (int i = 0; i < r.codes_n; ++i) out_faces[i] = face_code_at(r.codes, i);
  info2[0] = r.codes_n;
  info2[1] = static_cast<int>(r.faces);
  return r.n;
}
// ClipTriangleByTetrahedron (mesh_intersection.cc:176-264) with the sequence tri_stage_b runs: the triangle's
// vertices into frame M, four clips by the tetrahedron's face half spaces (built the way registration builds
// TetAux::plane_nh / plane_d: n = (B-A)x(C-A) over the face table {1,2,3},{0,3,2},{0,1,3},{0,2,1}, Plane ctor,
// d = nhat . A), dedupe. Returns the vertex count.
int phc_clip_triangle_by_tetrahedron(const double* tet_v, const double* tri_v, const double* X_MN, double* out) {
  Xf X;
  std::memcpy(X.m, X_MN, sizeof(X.m));
  LocalPoly P;
  for (int k = 0; k < 3; ++k) P.set(0, k, xform_point(X, load3(tri_v + 3 * k)));
  static const int kFaces[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
  int n = 3, buf = 0;
  for (int f = 0; f < 4 && n > 0; ++f) {
    const D3 A = load3(tet_v + 3 * kFaces[f][0]), B = load3(tet_v + 3 * kFaces[f][1]), Cc = load3(tet_v + 3 * kFaces[f][2]);
    D3 nh;
    double d;
    if (!MakeHalfSpace(cross(B - A, Cc - A), A, &nh, &d)) return -1;
    n = clip_by_half_space(P, buf, n, nh, d);
    if (n > LocalPoly::kCap) return -2;
    buf = 1 - buf;
  }
  n = remove_nearly_duplicate(P, buf, n);
  if (n < 3) return 0;
  for (int i = 0; i < n; ++i) {
    const D3 v = P.get(buf, i);
    out[3 * i] = v.x, out[3 * i + 1] = v.y, out[3 * i + 2] = v.z;
  }
  return n;
}
// The exact fp64 box test of the descent's undecided pairs (hydro_math.cuh::boxes_overlap = BoxesOverlap,
// boxes_overlap.cc:197-255): half widths a, b; X_AB as 12 doubles (R column-major, then p).
int phc_boxes_overlap(const double* a, const double* b, const double* X_AB) {
  Xf X;
  for (int k = 0; k < 12; ++k) X.m[k] = X_AB[k];
  return boxes_overlap(a, b, X, D3{X_AB[9], X_AB[10], X_AB[11]}) ? 1 : 0;
}
// The registration-time derivation of one tetrahedron's record (hydro_derive.cuh::derive_tet: what
// VolumeMesh::CalcGradBarycentric / MeshFieldLinear's gradient and value at the mesh origin / the inward face normals
// become on the device): v [12], u [4] -> grad [3], p_mo [1], inward [12]. Returns 0 for a degenerate tetrahedron.
int phc_derive_tet(const double* v, const double* u, double* grad, double* p_mo, double* inward) {
  const D3 V[4] = {load3(v), load3(v + 3), load3(v + 6), load3(v + 9)};
  TetRec rec;
  TetAux aux;
  if (!derive_tet(V, u, &rec, &aux)) return 0;
  for (int k = 0; k < 3; ++k) grad[k] = rec.grad[k];
  *p_mo = rec.p_mo;
  for (int f = 0; f < 4; ++f)
    for (int k = 0; k < 3; ++k) inward[3 * f + k] = aux.inward[f][k];
  return 1;
}
int phc_face_normal_along_gradient(const double* ngrad, const double* v) {
  return face_normal_along_gradient(load3(ngrad), load3(v)) ? 1 : 0;
}

}  // extern "C"
