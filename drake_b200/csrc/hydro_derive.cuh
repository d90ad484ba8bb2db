// Per-tetrahedron records derived from the four vertex positions and the four vertex field values. One function, used
// by the host at registration (dbh_register_compliant_mesh) and by the device when a volume mesh deforms
// (k_refit_tets, dbh_context_update_vertices): the same expressions in the same order, compiled without FMA
// contraction on both sides, so a mesh re-posed on the device to the positions it was registered with gets the bits it
// was registered with (tests/test_refit.py).
//   gradient + value at the frame origin   MeshFieldLinear::CalcGradientField / CalcValueAtMeshOrigin
//                                          (mesh_field_linear.h:440-509, volume_mesh.h:275-289,361-448)
//   inward face normals                    VolumeMesh::ComputePositionDependentQuantities (volume_mesh.cc:91-138)
//   face half spaces                       ClipTriangleByTetrahedron (mesh_intersection.cc:225-238) + Plane ctor
//                                          (plane.cc:16-29)
#pragma once
#include <limits>

#include "hydro_math.cuh"
#include "hydro_types.h"

namespace dbh {

__host__ __device__ inline double derive_quiet_nan() {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(0x7ff8000000000000ll);
#else
  return std::numeric_limits<double>::quiet_NaN();
#endif
}

// false: the tetrahedron is degenerate (a zero signed volume: the reference's field constructor throws).
__host__ __device__ inline bool derive_tet(const D3 v[4], const double u[4], TetRec* rec, TetAux* aux) {
  const double kEps = 2.220446049250313e-16;   // std::numeric_limits<double>::epsilon()
  TetRec& r = *rec;
  for (int i = 0; i < 4; ++i) {
    r.v[i][0] = v[i].x;
    r.v[i][1] = v[i].y;
    r.v[i][2] = v[i].z;
  }
  D3 g = make_d3(0, 0, 0);
  for (int i = 0; i < 4; ++i) {
    const D3 V = v[i], A = v[(i + 1) % 4], B = v[(i + 2) % 4], C = v[(i + 3) % 4];
    const D3 area_vector = cross(B - A, C - A);
    const double signed_volume = dot(area_vector, V - A);
    if (fabs(signed_volume) <= kEps) return false;
    g = g + u[i] * (area_vector / signed_volume);
  }
  r.grad[0] = g.x;
  r.grad[1] = g.y;
  r.grad[2] = g.z;
  r.p_mo = u[0] - dot(g, v[0]);  // CalcValueAtMeshOrigin (mesh_field_linear.h:502-509)
  // Inward normals (volume_mesh.cc:91-138).
  const D3 ab = v[1] - v[0], ac = v[2] - v[0], ad = v[3] - v[0], bc = v[2] - v[1], bd = v[3] - v[1];
  const D3 n[4] = {normalized(cross(bd, bc)), normalized(cross(ac, ad)), normalized(cross(ad, ab)),
                   normalized(cross(ab, ac))};
  TetAux& x = *aux;
  for (int f = 0; f < 4; ++f) {
    x.inward[f][0] = n[f].x;
    x.inward[f][1] = n[f].y;
    x.inward[f][2] = n[f].z;
  }
  // Face half spaces of ClipTriangleByTetrahedron (mesh_intersection.cc:225-238 + plane.cc:16-29): faces
  // {1,2,3}, {0,3,2}, {0,1,3}, {0,2,1}.
  const int kFaces[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
  for (int f = 0; f < 4; ++f) {
    const D3 A = v[kFaces[f][0]], B = v[kFaces[f][1]], C = v[kFaces[f][2]];
    const D3 normal = cross(B - A, C - A);
    const double mag = norm(normal);
    if (mag < 1e-10) {
      x.plane_nh[f][0] = x.plane_nh[f][1] = x.plane_nh[f][2] = 0.0;
      x.plane_d[f] = derive_quiet_nan();   // the Plane ctor throws when reached
    } else {
      const D3 nh = normal / mag;
      x.plane_nh[f][0] = nh.x;
      x.plane_nh[f][1] = nh.y;
      x.plane_nh[f][2] = nh.z;
      x.plane_d[f] = dot(nh, A);
    }
  }
  const D3 ng = normalized(g);
  x.ngrad[0] = ng.x;
  x.ngrad[1] = ng.y;
  x.ngrad[2] = ng.z;
  // Shape factor for the cull margins: the tetrahedron bounded by the four face planes offset outward by e is this
  // one scaled about its incentre by (1 + e / inradius).
  {
    const D3 cd = v[3] - v[2];
    const double diam2 = fmax(fmax(fmax(dot(ab, ab), dot(ac, ac)), fmax(dot(ad, ad), dot(bc, bc))), fmax(dot(bd, bd), dot(cd, cd)));
    const double vol6 = fabs(dot(cross(ab, ac), ad));
    const double area2 = norm(cross(bd, bc)) + norm(cross(ac, ad)) + norm(cross(ad, ab)) + norm(cross(ab, ac));
    const double inradius = area2 > 0 ? vol6 / area2 : 0.0;  // 3V / sum(A) = (vol6 / 2) / (area2 / 2)
    x.amp = inradius > 0 ? sqrt(diam2) / inradius : 1e300;
    if (!(x.amp < 1e300)) x.amp = 1e300;
  }
  return true;
}

// The fp32 copies the conservative stage reads.
__host__ __device__ inline void derive_tet_cull(const TetRec& t, const TetAux& x, TetCullA* ca, TetCullB* cb) {
  for (int k = 0; k < 4; ++k)
    for (int d = 0; d < 3; ++d) {
      ca->v[k][d] = static_cast<float>(t.v[k][d]);
      // Outward normal of the face opposite vertex k (a degenerate face gives 0: never culls).
      cb->nh[k][d] = static_cast<float>(-x.inward[k][d]);
    }
  for (int d = 0; d < 3; ++d) ca->grad[d] = static_cast<float>(t.grad[d]);
  ca->p_mo = static_cast<float>(t.p_mo);
  const double len = sqrt(t.grad[0] * t.grad[0] + t.grad[1] * t.grad[1] + t.grad[2] * t.grad[2]);
  cb->inv_grad_len = len > 0 ? static_cast<float>(1.0 / len) : 0.f;
  const double m = fmax(2e-5, 1e-11 * (1.0 + x.amp));
  cb->margin = m < 1e30 ? static_cast<float>(m) : 1.0f / 0.0f;
  cb->pad[0] = cb->pad[1] = 0.f;
}

}  // namespace dbh
