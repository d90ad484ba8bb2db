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

}  // namespace dbh
