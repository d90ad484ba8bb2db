// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// Restatement of the contact-surface consumer that SURVEY §8(f) ranks third: the continuous hydroelastic
// traction integration of multibody/plant/hydroelastic_traction_calculator.{h,cc} (per-face quadrature of
// the normal + regularised friction traction into one spatial force per surface) with the combination
// rules it is fed by (multibody/hydroelastics/hydroelastic_engine.cc:63-104, coulomb_friction.h:122-139) and the
// quadrature it uses (multibody/triangle_quadrature/{triangle_quadrature.h,gaussian_triangle_quadrature_rule.h}).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "ho_contact.h"

namespace ho {

// multibody/math/spatial_force.h / spatial_velocity.h: rotational part first, then translational.
struct Spatial {
  Vec3 rot, trans;
};
inline Spatial operator+(const Spatial& a, const Spatial& b) { return {a.rot + b.rot, a.trans + b.trans}; }
inline Spatial operator*(const Spatial& a, double s) { return {a.rot * s, a.trans * s}; }
inline Spatial operator-(const Spatial& a) { return {-a.rot, -a.trans}; }
// SpatialForce::Shift (spatial_force.h): tau_Bq = tau_Bp - p_BpBq x f.
inline Spatial ShiftForce(const Spatial& F, const Vec3& p) { return {F.rot - cross(p, F.trans), F.trans}; }
// SpatialVelocity::Shift (spatial_velocity.h): v_Bq = v_Bp + w x p_BpBq.
inline Spatial ShiftVelocity(const Spatial& V, const Vec3& p) { return {V.rot, V.trans + cross(V.rot, p)}; }

// hydroelastic_traction_calculator.cc:281-315.
inline double CalcAtanXOverXFromXSquared(double x2) {
  constexpr double x_cuttoff = 0.12;
  constexpr double x_cutoff_squared = x_cuttoff * x_cuttoff;
  if (x2 <= x_cutoff_squared) {
    return 1. - x2 * (1. / 3. - x2 * (1. / 5. - x2 * (1. / 7. - x2 * (1. / 9. - x2 * (1. / 11. - x2 * (1. / 13. -
           x2 * (1. / 15. - x2 / 17.)))))));
  }
  const double x = std::sqrt(x2);
  return std::atan(x) / x;
}

// HydroelasticEngine::CalcCombinedElasticModulus / CalcCombinedDissipation
// (multibody/hydroelastics/hydroelastic_engine.cc:63-104), the rule MultibodyPlant's continuous path feeds the
// traction calculator with (multibody_plant.cc:2645-2654). E = hydroelastic modulus, +infinity for a rigid geometry.
inline double CombinedElasticModulus(double E_A, double E_B) {
  const double kInf = std::numeric_limits<double>::infinity();
  if (E_A == kInf) return E_B;
  if (E_B == kInf) return E_A;
  return E_A * E_B / (E_A + E_B);
}
inline double CombinedHuntCrossleyDissipation(double E_A, double E_B, double d_A, double d_B) {
  const double kInf = std::numeric_limits<double>::infinity();
  const double Estar = CombinedElasticModulus(E_A, E_B);
  if (Estar == kInf) return 0.5 * (d_A + d_B);
  double d_star = 0;
  if (E_A != kInf) d_star += Estar / E_A * d_A;
  if (E_B != kInf) d_star += Estar / E_B * d_B;
  return d_star;
}
// coulomb_friction.h:122-139 (dynamic coefficient).
inline double CombinedDynamicFriction(double mu_A, double mu_B) {
  const double denom = mu_A + mu_B;
  return denom == 0.0 ? 0.0 : (2 * mu_A * mu_B) / denom;
}

// HydroelasticTractionCalculator::Data (hydroelastic_traction_calculator.h:40-80): A is the body of
// geometry M, B the body of geometry N; p_WC is the surface's centroid.
struct TractionData {
  Vec3 p_WAo, p_WBo;
  Spatial V_WA, V_WB;
  Vec3 p_WC;
};

// CalcTractionAtQHelper (hydroelastic_traction_calculator.cc:218-279): traction on A at Q, in World.
inline Vec3 CalcTractionAtQ(const TractionData& data, double e, const Vec3& nhat_W, double dissipation,
                            double mu_coulomb, double vslip_regularizer, const Vec3& p_WQ) {
  const Vec3 p_AoAq_W = p_WQ - data.p_WAo;
  const Spatial V_WAq = ShiftVelocity(data.V_WA, p_AoAq_W);
  const Vec3 p_BoBq_W = p_WQ - data.p_WBo;
  const Spatial V_WBq = ShiftVelocity(data.V_WB, p_BoBq_W);
  const Vec3 v_BqAq_W = V_WAq.trans - V_WBq.trans;
  const double vn_BqAq_W = dot(v_BqAq_W, nhat_W);
  const double c = dissipation * e;
  const double normal_traction = std::max(e - vn_BqAq_W * c, 0.0);
  const Vec3 vt_BqAq_W = v_BqAq_W - nhat_W * vn_BqAq_W;
  const double vs_squared = vslip_regularizer * vslip_regularizer;
  const double x_squared = squared_norm(vt_BqAq_W) / vs_squared;
  const double regularized_friction =
      (2.0 / M_PI) * mu_coulomb * normal_traction * CalcAtanXOverXFromXSquared(x_squared) / vslip_regularizer;
  const Vec3 ft_Aq_W = (-regularized_friction) * vt_BqAq_W;
  return nhat_W * normal_traction + ft_Aq_W;
}

// ComputeSpatialTractionAtAcFromTractionAtAq (:85-96).
inline Spatial TractionAtAc(const TractionData& data, const Vec3& p_WQ, const Vec3& traction_Aq_W) {
  const Vec3 p_QC_W = data.p_WC - p_WQ;
  return ShiftForce(Spatial{Vec3{0, 0, 0}, traction_Aq_W}, p_QC_W);
}

// ComputeSpatialForcesAtCentroidFromHydroelasticModel (:24-62): triangles by the order-2 Gaussian rule
// (3 points, weights 1/3, gaussian_triangle_quadrature_rule.h:44-51; TriangleQuadrature::Integrate sums
// f(b_i) w_i then multiplies by the area), polygons by one point at the centroid.
inline Spatial SpatialForceAtCentroid(const SurfaceResult& s, const TractionData& data, double dissipation,
                                      double mu_coulomb, double vslip_regularizer) {
  Spatial F_Ac_W{Vec3{0, 0, 0}, Vec3{0, 0, 0}};
  const double kB[3][2] = {{1.0 / 6.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}, {2.0 / 3.0, 1.0 / 6.0}};
  const double kW = 1.0 / 3.0;
  for (int i = 0; i < s.num_faces(); ++i) {
    if (s.rep == Representation::kTriangle) {
      const Tri& t = s.tris[i];
      auto f = [&](int k) {
        const double b0 = kB[k][0], b1 = kB[k][1], b2 = 1.0 - kB[k][0] - kB[k][1];
        // CalcCartesianFromBarycentric (triangle_surface_mesh.h:266-281): T_CB * b.
        const Vec3 p_WQ = (s.vertices_W[t[0]] * b0 + s.vertices_W[t[1]] * b1) + s.vertices_W[t[2]] * b2;
        // MeshFieldLinear::Evaluate (mesh_field_linear.h:300-312).
        double e = b0 * s.pressures[t[0]];
        e += b1 * s.pressures[t[1]];
        e += b2 * s.pressures[t[2]];
        const Vec3 traction = CalcTractionAtQ(data, e, s.normals_W[i], dissipation, mu_coulomb, vslip_regularizer, p_WQ);
        return TractionAtAc(data, p_WQ, traction);
      };
      Spatial integral = f(0) * kW;
      integral = integral + f(1) * kW;
      integral = integral + f(2) * kW;
      F_Ac_W = F_Ac_W + integral * s.areas[i];
    } else {
      const Vec3& p_WQ = s.centroids_W[i];
      const Vec3 traction =
          CalcTractionAtQ(data, s.e_centroid[i], s.normals_W[i], dissipation, mu_coulomb, vslip_regularizer, p_WQ);
      F_Ac_W = F_Ac_W + TractionAtAc(data, p_WQ, traction) * s.areas[i];
    }
  }
  return F_Ac_W;
}

// ShiftSpatialForcesAtCentroidToBodyOrigins (:64-83).
inline void ShiftToBodyOrigins(const TractionData& data, const Spatial& F_Ac_W, Spatial* F_Ao_W, Spatial* F_Bo_W) {
  *F_Ao_W = ShiftForce(F_Ac_W, data.p_WAo - data.p_WC);
  *F_Bo_W = -ShiftForce(F_Ac_W, data.p_WBo - data.p_WC);
}

}  // namespace ho
