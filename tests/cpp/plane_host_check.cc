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
#include "../../drake_b200/csrc/hydro_plane.cuh"
#include "../../drake_b200/csrc/hydro_wrench.cuh"

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
double phc_combined_dissipation(double kA, double kB, double dA, double dB) {
  return combined_hunt_crossley_dissipation(kA, kB, dA, dB);
}
double phc_combined_friction(double a, double b) { return combined_dynamic_friction(a, b); }

}  // extern "C"
