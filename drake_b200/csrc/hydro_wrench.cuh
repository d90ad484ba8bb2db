// Per-face arithmetic of the contact-surface consumer SURVEY §8(f) ranks third: the continuous
// hydroelastic traction integration of multibody/plant/hydroelastic_traction_calculator.cc (one spatial
// force per contact surface, so that only 18 doubles per surface have to leave the GPU instead of the
// whole mesh), with the combination rules that feed it (hydroelastic_engine.cc:63-104,
// coulomb_friction.h:122-139). __host__ __device__ over plain values, like hydro_plane.cuh: the CPU test
// suite compiles these expressions with g++ (tests/cpp/plane_host_check.cc); the product only calls them
// from k_wrench. Operation order follows the reference.
#pragma once
#include "hydro_math.cuh"

namespace dbh {

// multibody/math spatial vectors: rotational part first, then translational.
struct Spatial6 {
  D3 rot, trans;
};
__host__ __device__ __forceinline__ Spatial6 operator+(Spatial6 a, Spatial6 b) {
  return Spatial6{a.rot + b.rot, a.trans + b.trans};
}
__host__ __device__ __forceinline__ Spatial6 scaled(Spatial6 a, double s) { return Spatial6{s * a.rot, s * a.trans}; }
// SpatialForce::Shift: tau_Bq = tau_Bp - p_BpBq x f.
__host__ __device__ __forceinline__ Spatial6 shift_force(Spatial6 F, D3 p) {
  return Spatial6{F.rot - cross(p, F.trans), F.trans};
}

// CalcAtanXOverXFromXSquared (hydroelastic_traction_calculator.cc:281-315).
__host__ __device__ __forceinline__ double atan_x_over_x_from_x_squared(double x2) {
  if (x2 <= 0.12 * 0.12) {
    return 1. - x2 * (1. / 3. - x2 * (1. / 5. - x2 * (1. / 7. - x2 * (1. / 9. - x2 * (1. / 11. - x2 * (1. / 13. -
           x2 * (1. / 15. - x2 / 17.)))))));
  }
  const double x = sqrt(x2);
  return atan(x) / x;
}

// HydroelasticEngine::CalcCombinedElasticModulus / CalcCombinedDissipation
// (multibody/hydroelastics/hydroelastic_engine.cc:63-104): what MultibodyPlant's continuous path hands to the
// traction calculator (multibody_plant.cc:2645-2654). E = hydroelastic modulus, +infinity for a rigid geometry.
__host__ __device__ __forceinline__ double combined_hunt_crossley_dissipation(double E_A, double E_B, double d_A,
                                                                              double d_B) {
  const bool rigid_A = isinf(E_A), rigid_B = isinf(E_B);
  if (rigid_A && rigid_B) return 0.5 * (d_A + d_B);
  const double Estar = rigid_A ? E_B : (rigid_B ? E_A : E_A * E_B / (E_A + E_B));
  double d_star = 0;
  if (!rigid_A) d_star += Estar / E_A * d_A;
  if (!rigid_B) d_star += Estar / E_B * d_B;
  return d_star;
}
// CalcContactFrictionFromSurfaceProperties (coulomb_friction.h:122-139), dynamic coefficient.
__host__ __device__ __forceinline__ double combined_dynamic_friction(double mu_A, double mu_B) {
  const double denom = mu_A + mu_B;
  return denom == 0.0 ? 0.0 : (2 * mu_A * mu_B) / denom;
}

// HydroelasticTractionCalculator::Data (hydroelastic_traction_calculator.h:40-80) with the scalars of one
// surface: A is the body of geometry M, B the body of geometry N, p_WC the surface's centroid.
struct TractionBodies {
  D3 p_WAo, p_WBo;
  D3 w_WA, v_WA, w_WB, v_WB;
  D3 p_WC;
  double dissipation, mu_coulomb, vslip_regularizer;
};

// CalcTractionAtQHelper (:218-279) then ComputeSpatialTractionAtAcFromTractionAtAq (:85-96): the traction
// on A at Q as a spatial traction about C.
__host__ __device__ __forceinline__ Spatial6 spatial_traction_at_ac(const TractionBodies& d, double e, D3 nhat_W,
                                                                    D3 p_WQ) {
  const D3 v_WAq = d.v_WA + cross(d.w_WA, p_WQ - d.p_WAo);  // SpatialVelocity::Shift
  const D3 v_WBq = d.v_WB + cross(d.w_WB, p_WQ - d.p_WBo);
  const D3 v_BqAq_W = v_WAq - v_WBq;
  const double vn_BqAq_W = dot(v_BqAq_W, nhat_W);
  const double c = d.dissipation * e;
  const double raw = e - vn_BqAq_W * c;
  const double normal_traction = raw > 0.0 ? raw : 0.0;  // max(., 0)
  const D3 vt_BqAq_W = v_BqAq_W - vn_BqAq_W * nhat_W;
  const double vs_squared = d.vslip_regularizer * d.vslip_regularizer;
  const double x_squared = dot(vt_BqAq_W, vt_BqAq_W) / vs_squared;
  const double regularized_friction = (2.0 / 3.14159265358979323846) * d.mu_coulomb * normal_traction *
                                      atan_x_over_x_from_x_squared(x_squared) / d.vslip_regularizer;
  const D3 traction_Aq_W = normal_traction * nhat_W + (-regularized_friction) * vt_BqAq_W;
  return shift_force(Spatial6{make_d3(0, 0, 0), traction_Aq_W}, d.p_WC - p_WQ);
}

// One polygon: a single quadrature point at its centroid (:50-58).
__host__ __device__ __forceinline__ Spatial6 polygon_face_force(const TractionBodies& d, double area, D3 nhat_W,
                                                                D3 centroid_W, double e_centroid) {
  return scaled(spatial_traction_at_ac(d, e_centroid, nhat_W, centroid_W), area);
}

// One triangle: order-2 Gaussian rule (gaussian_triangle_quadrature_rule.h:44-51: barycentric points
// (1/6,1/6), (1/6,2/3), (2/3,1/6), weights 1/3), TriangleQuadrature::Integrate (sum of f(b_i) w_i, times the
// area), CalcCartesianFromBarycentric (triangle_surface_mesh.h:266-281), MeshFieldLinear::Evaluate
// (mesh_field_linear.h:300-312).
__host__ __device__ __forceinline__ Spatial6 triangle_face_force(const TractionBodies& d, double area, D3 nhat_W,
                                                                 D3 va, D3 vb, D3 vc, double ea, double eb, double ec) {
  const double kB0[3] = {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0};
  const double kB1[3] = {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0};
  const double kW = 1.0 / 3.0;
  Spatial6 integral{make_d3(0, 0, 0), make_d3(0, 0, 0)};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double b0 = kB0[k], b1 = kB1[k], b2 = 1.0 - kB0[k] - kB1[k];
    const D3 p_WQ = (b0 * va + b1 * vb) + b2 * vc;
    double e = b0 * ea;
    e += b1 * eb;
    e += b2 * ec;
    const Spatial6 f = scaled(spatial_traction_at_ac(d, e, nhat_W, p_WQ), kW);
    integral = k == 0 ? f : integral + f;
  }
  return scaled(integral, area);
}

// ShiftSpatialForcesAtCentroidToBodyOrigins (:64-83).
__host__ __device__ __forceinline__ void shift_to_body_origins(const TractionBodies& d, Spatial6 F_Ac_W, Spatial6* F_Ao_W,
                                                               Spatial6* F_Bo_W) {
  *F_Ao_W = shift_force(F_Ac_W, d.p_WAo - d.p_WC);
  const Spatial6 s = shift_force(F_Ac_W, d.p_WBo - d.p_WC);
  *F_Bo_W = Spatial6{-s.rot, -s.trans};
}

// The geometry-only half of DiscreteUpdateManager::AppendDiscreteContactPairsForHydroelasticContact
// (multibody/plant/discrete_update_manager.cc:896-1036): what the discrete solvers read from one face of a contact
// surface -- first-order quadrature, one point per face at its centroid. The other half (Jacobians, vn0, friction
// and dissipation look-ups) needs the plant's kinematic tree and stays on the host.
struct DiscreteFacePair {
  D3 p_WC;        // quadrature point = face centroid                                  (:965-967)
  D3 nhat_BA_W;   // face normal, out of N into M                                      (:926-929)
  double phi0;    // -p0 / g, negative in penetration                                  (:1028-1029)
  double fn0;     // Ae * p0                                                           (:1017-1018)
  double stiffness;  // k = Ae * g; 0 marks a face the reference skips                  (:1025)
};
// has_M / has_N: HasGradE_M / HasGradE_N of the surface. Returns false for the faces the reference skips:
// Ae <= 1e-14 (:909-919) or a pressure gradient along the normal below 1e-14 on either side (:941-949).
__host__ __device__ __forceinline__ bool discrete_face_pair(double Ae, D3 nhat_BA_W, D3 p_WC, double p0, D3 grad_eM_W,
                                                            D3 grad_eN_W, bool has_M, bool has_N, DiscreteFacePair* out) {
  out->p_WC = p_WC;
  out->nhat_BA_W = nhat_BA_W;
  out->phi0 = 0.0;
  out->fn0 = 0.0;
  out->stiffness = 0.0;
  if (!(Ae > 1.0e-14)) return false;
  const double kInf = 1.0 / 0.0;
  const double gM = has_M ? dot(grad_eM_W, nhat_BA_W) : kInf;
  const double gN = has_N ? -dot(grad_eN_W, nhat_BA_W) : kInf;
  const double kGradientEpsilon = 1.0e-14;
  if (gM < kGradientEpsilon || gN < kGradientEpsilon) return false;
  const double g = 1.0 / (1.0 / gM + 1.0 / gN);
  out->fn0 = Ae * p0;
  out->stiffness = Ae * g;
  out->phi0 = -p0 / g;
  return true;
}

// ---- deformable x rigid contact data ------------------------------------------------------------
// What the reference derives per polygon of a tet-triangle contact surface whose triangle mesh is a DEFORMABLE body's
// surface and whose volume mesh + pressure field belong to a rigid geometry (AddDeformableRigidContactSurface,
// geometry/proximity/deformable_mesh_intersection.cc:85-183; DeformableContactSurface ctor,
// geometry/query_results/deformable_contact.cc:35-66):
struct DeformableFaceContact {
  D3 p_WC;          // contact point = polygon centroid                                   (deformable_contact.cc:59)
  D3 nhat_W;        // from the rigid body into the deformable one: -face normal          (:58)
  double R_WC[9];   // MakeFromOneUnitVector(-nhat_W, z), column-major                    (:60-61)
  double pressure;  // rigid pressure field at the centroid                               (deformable_mesh_intersection.cc:112-121)
  D3 grad_W;        // X_WR * grad_R: the POINT transform, as the reference writes it     (:172-174)
  D3 barycentric;   // of the centroid in the deformable triangle                         (:69-79)
};

// RotationMatrix::MakeFromOneUnitVector (math/rotation_matrix.cc:133-242), column-major into R.
__host__ __device__ __forceinline__ void make_from_one_unit_vector(D3 u, int axis_index, double* R) {
  const double uu[3] = {u.x, u.y, u.z};
  int i = 0;   // minCoeff of |u|: the first of equal minima
  if (fabs(uu[1]) < fabs(uu[i])) i = 1;
  if (fabs(uu[2]) < fabs(uu[i])) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  const double mag_a_x_u = sqrt(1 - uu[i] * uu[i]);
  const double r = 1 / mag_a_x_u;
  const double s = -r * uu[i];
  double v[3], w[3];
  v[i] = 0;
  v[j] = -r * uu[k];
  v[k] = r * uu[j];
  w[i] = mag_a_x_u;
  w[j] = s * uu[j];
  w[k] = s * uu[k];
  const int cv = (axis_index + 1) % 3, cw = (axis_index + 2) % 3;
  for (int q = 0; q < 3; ++q) {
    R[axis_index * 3 + q] = uu[q];
    R[cv * 3 + q] = v[q];
    R[cw * 3 + q] = w[q];
  }
}

// TriangleSurfaceMesh::CalcBarycentric (triangle_surface_mesh.cc:53-87): the least-squares solution of
// [v1 - v0, v2 - v0] b = p - v0 -- there through Eigen's colPivHouseholderQr, here through the 2 x 2 normal equations
// (the same solution up to rounding: the parity tests compare it with a tolerance).
__host__ __device__ __forceinline__ D3 calc_barycentric(D3 p, D3 v0, D3 v1, D3 v2) {
  const D3 u1 = v1 - v0, u2 = v2 - v0, r = p - v0;
  const double a11 = dot(u1, u1), a12 = dot(u1, u2), a22 = dot(u2, u2);
  const double c1 = dot(u1, r), c2 = dot(u2, r);
  const double det = a11 * a22 - a12 * a12;
  const double b1 = (a22 * c1 - a12 * c2) / det;
  const double b2 = (a11 * c2 - a12 * c1) / det;
  return D3{1.0 - b1 - b2, b1, b2};
}

// VolumeMesh::CalcBarycentric (volume_mesh.cc:24-52) solves the 4 x 4 system [1 ... ; v0 v1 v2 v3] b = [1; p] by LU;
// here Cramer's rule on signed volumes (the same solution up to rounding: compared with a tolerance).
__host__ __device__ __forceinline__ void calc_barycentric_tet(D3 p, D3 v0, D3 v1, D3 v2, D3 v3, double b[4]) {
  const double vol = dot(cross(v1 - v0, v2 - v0), v3 - v0);
  b[0] = dot(cross(v1 - p, v2 - p), v3 - p) / vol;
  b[1] = dot(cross(p - v0, v2 - v0), v3 - v0) / vol;
  b[2] = dot(cross(v1 - v0, p - v0), v3 - v0) / vol;
  b[3] = dot(cross(v1 - v0, v2 - v0), p - v0) / vol;
}

// Deformable x deformable (AddDeformableDeformableContactSurface, deformable_field_intersection.cc:46-124, and the second
// DeformableContactSurface constructor, deformable_contact.cc:68-100): the contact mesh is the equilibrium surface of the
// two signed-distance fields; per polygon the contact point (centroid), the mesh's face normal, R_WC, the signed
// distance there and the barycentric coordinates of the point in the two tetrahedra that made the polygon.
struct DeformablePairContact {
  D3 p_WC;
  D3 nhat_W;        // contact_mesh_W.face_normal(i)                                     (deformable_contact.cc:96)
  double R_WC[9];   // MakeFromOneUnitVector(-nhat_W, z), column-major                   (:98-99)
  double signed_distance;
  double bary0[4];  // in the tetrahedron of the intersector's first field (field0, its mesh M)
  double bary1[4];  // in the tetrahedron of the second field
};
__host__ __device__ __forceinline__ DeformablePairContact deformable_pair_contact(D3 normal_W, D3 centroid_W, double value,
                                                                                  const Xf& X_WM, const Xf& X_WN,
                                                                                  const double tet0[4][3], const double tet1[4][3]) {
  DeformablePairContact c;
  c.p_WC = centroid_W;
  c.nhat_W = normal_W;
  make_from_one_unit_vector(-normal_W, 2, c.R_WC);
  c.signed_distance = value;
  const D3 p_M = rot_t(X_WM, centroid_W - make_d3(X_WM.m[9], X_WM.m[10], X_WM.m[11]));
  const D3 p_N = rot_t(X_WN, centroid_W - make_d3(X_WN.m[9], X_WN.m[10], X_WN.m[11]));
  calc_barycentric_tet(p_M, load3(tet0[0]), load3(tet0[1]), load3(tet0[2]), load3(tet0[3]), c.bary0);
  calc_barycentric_tet(p_N, load3(tet1[0]), load3(tet1[1]), load3(tet1[2]), load3(tet1[3]), c.bary1);
  return c;
}

// mesh_normal_W: the contact mesh's face normal as the intersector built it (out of the triangle's front side: the
// deformable surface's outward normal); grad_W: the rigid field's gradient rotated into World; X_WR / X_WD: poses of the
// rigid geometry and of the deformable surface mesh (identity in the reference); tri: the triangle's vertices in D.
__host__ __device__ __forceinline__ DeformableFaceContact deformable_face_contact(D3 mesh_normal_W, D3 centroid_W,
                                                                                  double pressure, D3 grad_W, const Xf& X_WR,
                                                                                  const Xf& X_WD, D3 t0, D3 t1, D3 t2) {
  DeformableFaceContact c;
  c.p_WC = centroid_W;
  c.nhat_W = -mesh_normal_W;
  make_from_one_unit_vector(mesh_normal_W, 2, c.R_WC);
  c.pressure = pressure;
  c.grad_W = D3{X_WR.m[9] + grad_W.x, X_WR.m[10] + grad_W.y, X_WR.m[11] + grad_W.z};
  const D3 p_DC = rot_t(X_WD, centroid_W - make_d3(X_WD.m[9], X_WD.m[10], X_WD.m[11]));
  c.barycentric = calc_barycentric(p_DC, t0, t1, t2);
  return c;
}

}  // namespace dbh
