// drake_b200_point_pair.h — the point-pair half of ComputeContactSurfacesWithFallback, host-side C++17.
//
// The reference hands every pair whose hydroelastic surface cannot be made (rigid-rigid, a geometry with no
// hydroelastic representation, the rejected half-space combinations) to penetration_as_point_pair::MaybeMakePointPair
// (geometry/proximity_engine.cc:958-970). That narrow phase has two halves:
//   * Drake's own closed forms, used whenever one of the two shapes is a sphere
//     (penetration_as_point_pair_callback.cc:92-118,203-300: sphere x {sphere, box, half space, cylinder, capsule})
//     through point_distance::DistanceToPoint (distance_to_point_callback.cc:15-77,160-178,406-482);
//   * FCL's collide() for everything else (:145-207) -- an external dependency of the reference, absent from its tree
//     (tools/workspace/fcl_internal/repository.bzl: RobotLocomotion/fcl at commit e5efcc41b57b). Inside FCL that call is
//     again three things: closed-form primitives for a Halfspace against Box, Capsule, Cylinder, Ellipsoid and Convex
//     (fcl/narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h), box x box (box_box-inl.h), and libccd's
//     Minkowski Portal Refinement for every other convex pair (convexity_based_algorithm/gjk_libccd-inl.h: GJKCollide).
// This header restates the first half completely -- Sphere against Sphere, Box, HalfSpace, Cylinder, Capsule, every shape
// the reference has a closed form for (it sends Ellipsoid and Convex to FCL, :273-278) -- and the second half from FCL's
// and libccd's published algorithms, with the reference's post-processing of an FCL contact (:164-207). The FCL
// sources are not available here: the half-space functions are pinned by the reference's BoxPenetrationTest TangentPlane scenes
// (penetration_as_point_pair_callback_test.cc:575-910, 1e-12) and by the shapes' analytic support functions
// (tests/test_point_pair_cpu.py), not against FCL itself; likewise FCL's box x box (ODE's dBoxBox, below), pinned by the
// TangentBox scenes, and libccd's Minkowski Portal Refinement for every other convex pair (MprPenetration, below), pinned
// by the same test's cylinder / capsule / convex scenes at the tolerances the reference states. The batched
// query (dbh_batch_query, DBH_QUERY_WITH_FALLBACK) reports WHICH pairs need it (dbh_batch_copy_fallback_pairs): a
// handful per scene, so the arithmetic stays on the host next to the shape descriptions the C-ABI never sees.
// No Eigen: 3-vectors are std::array<double, 3>; operation order follows the reference.
#ifndef DRAKE_B200_POINT_PAIR_H_
#define DRAKE_B200_POINT_PAIR_H_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <limits>
#include <optional>

namespace drake_b200 {
namespace point_pair {

using Vec3 = std::array<double, 3>;
// RigidTransform<double> as the C-ABI lays it out: rotation column-major (9), then translation (3).
using Pose = std::array<double, 12>;

inline Vec3 Rotate(const Pose& X, const Vec3& v) {
  return {(X[0] * v[0] + X[3] * v[1]) + X[6] * v[2], (X[1] * v[0] + X[4] * v[1]) + X[7] * v[2],
          (X[2] * v[0] + X[5] * v[1]) + X[8] * v[2]};
}
inline Vec3 RotateT(const Pose& X, const Vec3& v) {
  return {(X[0] * v[0] + X[1] * v[1]) + X[2] * v[2], (X[3] * v[0] + X[4] * v[1]) + X[5] * v[2],
          (X[6] * v[0] + X[7] * v[1]) + X[8] * v[2]};
}
inline Vec3 Apply(const Pose& X, const Vec3& v) {  // X * v
  const Vec3 r = Rotate(X, v);
  return {X[9] + r[0], X[10] + r[1], X[11] + r[2]};
}
inline Vec3 ApplyInverse(const Pose& X, const Vec3& p) {  // X.inverse() * p = R^T (p - t)
  return RotateT(X, {p[0] - X[9], p[1] - X[10], p[2] - X[11]});
}
inline double Dot(const Vec3& a, const Vec3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline double Norm(const Vec3& a) { return std::sqrt(Dot(a, a)); }

// DistanceToPointRelativeTolerance (proximity_utilities.h:29-31).
inline double RelativeTolerance(double size) { return 1e-14 * std::max(1.0, size); }
// DistanceToPoint::Sign: Sign(0) = 1 (distance_to_point_callback.h:185-191).
inline double Sign(double x) { return x < 0.0 ? -1.0 : 1.0; }

// SignedDistanceToPoint<double> (query_results/signed_distance_to_point.h): nearest point N on the boundary of G in
// G's frame, signed distance (negative inside), gradient in World.
struct SignedDistanceToPoint {
  Vec3 p_GN;
  double distance;
  Vec3 grad_W;
};

// DistanceToPoint::operator()(Sphere) (distance_to_point_callback.cc:15-58,361-369).
inline SignedDistanceToPoint DistanceToSphere(double radius, const Pose& X_WG, const Vec3& p_WQ) {
  const Vec3 p_GQ = ApplyInverse(X_WG, p_WQ);
  const double dist = Norm(p_GQ);
  const double tolerance = RelativeTolerance(radius);
  const Vec3 grad_G = dist > tolerance ? Vec3{p_GQ[0] / dist, p_GQ[1] / dist, p_GQ[2] / dist} : Vec3{1.0, 0.0, 0.0};
  const Vec3 p_GN{radius * grad_G[0], radius * grad_G[1], radius * grad_G[2]};
  const double distance = Dot({p_GQ[0] - p_GN[0], p_GQ[1] - p_GN[1], p_GQ[2] - p_GN[2]}, grad_G);
  return {p_GN, distance, Rotate(X_WG, grad_G)};
}

// DistanceToPoint::operator()(HalfSpace) (distance_to_point_callback.cc:60-77,350-359): the half space is z <= 0.
inline SignedDistanceToPoint DistanceToHalfSpace(const Pose& X_WG, const Vec3& p_WQ) {
  const Vec3 p_GQ = ApplyInverse(X_WG, p_WQ);
  return {{p_GQ[0], p_GQ[1], 0.0}, p_GQ[2], Rotate(X_WG, {0.0, 0.0, 1.0})};
}

// DistanceToPoint::ComputeDistanceToBox<dim> (distance_to_point_callback.cc:406-482) with ExtremalAxis (:387-403): the
// nearest point on the boundary of the box [-h, h]^dim and the gradient, both in the box frame, for a query point p.
// Three regions: outside (clamp; gradient along p - N), within a relative tolerance of the boundary (normalised sum of
// the signed axes it touches), inside (the nearest face; ties go to +x, -x, +y, -y, ...).
template <int dim>
inline void ComputeDistanceToBox(const std::array<double, dim>& h, const std::array<double, dim>& p,
                                 std::array<double, dim>* p_GN, std::array<double, dim>* grad_G) {
  enum Location { kInside, kBoundary, kOutside };
  Location location[dim];
  std::array<double, dim> c{};
  bool any_outside = false, any_boundary = false;
  for (int i = 0; i < dim; ++i) {
    const double tolerance = RelativeTolerance(h[i]);
    if (std::abs(p[i]) > h[i] + tolerance) {
      location[i] = kOutside;
      c[i] = Sign(p[i]) * h[i];
      any_outside = true;
    } else if (std::abs(p[i]) >= h[i] - tolerance) {
      location[i] = kBoundary;
      c[i] = Sign(p[i]) * h[i];
      any_boundary = true;
    } else {
      location[i] = kInside;
      c[i] = p[i];
    }
  }
  *p_GN = c;
  grad_G->fill(0.0);
  if (any_outside) {
    double n2 = 0;
    for (int i = 0; i < dim; ++i) n2 += (p[i] - c[i]) * (p[i] - c[i]);
    const double distance = std::sqrt(n2);
    for (int i = 0; i < dim; ++i) (*grad_G)[i] = (p[i] - c[i]) / distance;
  } else if (any_boundary) {
    double n2 = 0;
    for (int i = 0; i < dim; ++i) {
      if (location[i] == kBoundary) (*grad_G)[i] = Sign(c[i]);
      n2 += (*grad_G)[i] * (*grad_G)[i];
    }
    const double n = std::sqrt(n2);
    for (int i = 0; i < dim; ++i) (*grad_G)[i] /= n;
  } else {
    double min_dist = std::numeric_limits<double>::infinity();
    int axis = -1;
    for (int i = 0; i < dim; ++i)
      for (const double bound : {h[i], -h[i]}) {
        const double dist = std::abs(bound - p[i]);
        if (dist < min_dist) {
          min_dist = dist;
          axis = i;
        }
      }
    const double sign = Sign(p[axis]);
    (*p_GN)[axis] = sign * h[axis];
    (*grad_G)[axis] = sign;
  }
}

// DistanceToPoint::operator()(Box) (distance_to_point_callback.cc:160-178); size = full side lengths.
inline SignedDistanceToPoint DistanceToBox(const Vec3& size, const Pose& X_WG, const Vec3& p_WQ) {
  const Vec3 p_GQ = ApplyInverse(X_WG, p_WQ);
  Vec3 p_GN, grad_G;
  ComputeDistanceToBox<3>({size[0] / 2.0, size[1] / 2.0, size[2] / 2.0}, p_GQ, &p_GN, &grad_G);
  const Vec3 grad_W = Rotate(X_WG, grad_G);
  const Vec3 p_WN = Apply(X_WG, p_GN);
  const double distance = Dot(grad_W, {p_WQ[0] - p_WN[0], p_WQ[1] - p_WN[1], p_WQ[2] - p_WN[2]});
  return {p_GN, distance, grad_W};
}

// DistanceToPoint::operator()(Cylinder) (distance_to_point_callback.cc:194-288): the cylinder (axis z, radius r, length
// L) becomes the 2-D box [-r, r] x [-L/2, L/2] of its cross section through the query point and the axis; a point within
// a relative tolerance of the axis takes the cross section through +x.
inline SignedDistanceToPoint DistanceToCylinder(double radius, double length, const Pose& X_WG, const Vec3& p_WQ) {
  const Vec3 p_GQ = ApplyInverse(X_WG, p_WQ);
  const double r_Q = std::sqrt(p_GQ[0] * p_GQ[0] + p_GQ[1] * p_GQ[1]);
  const bool near_center_line = r_Q < RelativeTolerance(radius);
  const std::array<double, 2> v_GBr = near_center_line ? std::array<double, 2>{1.0, 0.0}
                                                       : std::array<double, 2>{p_GQ[0] / r_Q, p_GQ[1] / r_Q};
  std::array<double, 2> p_BN, grad_B;
  ComputeDistanceToBox<2>({radius, length / 2.0}, {r_Q, p_GQ[2]}, &p_BN, &grad_B);
  const Vec3 p_GN{p_BN[0] * v_GBr[0], p_BN[0] * v_GBr[1], p_BN[1]};
  const Vec3 grad_G{grad_B[0] * v_GBr[0], grad_B[0] * v_GBr[1], grad_B[1]};
  const Vec3 grad_W = Rotate(X_WG, grad_G);
  const Vec3 p_WN = Apply(X_WG, p_GN);
  const double distance = Dot(grad_W, {p_WQ[0] - p_WN[0], p_WQ[1] - p_WN[1], p_WQ[2] - p_WN[2]});
  return {p_GN, distance, grad_W};
}

// SphereDistanceInSphereFrame (distance_to_point_callback.cc:15-58): nearest point, distance and gradient of a sphere
// centred on the origin of the frame the query point is given in.
inline void SphereDistanceInSphereFrame(double radius, const Vec3& p_SQ, Vec3* p_SN, double* distance, Vec3* grad_S) {
  const double dist = Norm(p_SQ);
  const double tolerance = RelativeTolerance(radius);
  *grad_S = dist > tolerance ? Vec3{p_SQ[0] / dist, p_SQ[1] / dist, p_SQ[2] / dist} : Vec3{1.0, 0.0, 0.0};
  *p_SN = {radius * (*grad_S)[0], radius * (*grad_S)[1], radius * (*grad_S)[2]};
  *distance = Dot({p_SQ[0] - (*p_SN)[0], p_SQ[1] - (*p_SN)[1], p_SQ[2] - (*p_SN)[2]}, *grad_S);
}

// ComputeDistanceToPrimitive(Capsule) (distance_to_point_callback.cc:79-160): axis z, radius r, the spine of length L.
// Beyond the spine's ends the nearest cap is a sphere; beside the spine the problem is that of the query point's
// projection onto the x-y plane against a sphere at the origin, shifted back along z.
inline SignedDistanceToPoint DistanceToCapsule(double radius, double length, const Pose& X_WG, const Vec3& p_WQ) {
  const double half_length = length / 2;
  const Vec3 p_GQ = ApplyInverse(X_WG, p_WQ);
  Vec3 p_GN, grad_G;
  double distance;
  if (p_GQ[2] >= half_length || p_GQ[2] <= -half_length) {
    const Vec3 p_GS{0.0, 0.0, p_GQ[2] >= half_length ? half_length : -half_length};
    Vec3 p_SN;
    SphereDistanceInSphereFrame(radius, {p_GQ[0] - p_GS[0], p_GQ[1] - p_GS[1], p_GQ[2] - p_GS[2]}, &p_SN, &distance, &grad_G);
    p_GN = {p_GS[0] + p_SN[0], p_GS[1] + p_SN[1], p_GS[2] + p_SN[2]};
  } else {
    Vec3 p_GM;
    SphereDistanceInSphereFrame(radius, {p_GQ[0], p_GQ[1], 0.0}, &p_GM, &distance, &grad_G);
    p_GN = {p_GM[0], p_GM[1], p_GQ[2]};
  }
  return {p_GN, distance, Rotate(X_WG, grad_G)};
}

struct Result {  // PenetrationAsPointPair<double> without the ids
  Vec3 p_WCa, p_WCb, nhat_BA_W;
  double depth;
};

// SphereShapePenetration (penetration_as_point_pair_callback.cc:92-118): sphere A of radius r at X_WA against the
// signed distance of the other shape B to A's centre. Empty when depth < 0.
inline std::optional<Result> SphereShapePenetration(double radius_A, const Pose& X_WA, const Pose& X_WB,
                                                    const SignedDistanceToPoint& shape_B_to_point_Ao) {
  const double depth = radius_A - shape_B_to_point_Ao.distance;
  if (!(depth >= 0)) return std::nullopt;
  Result r;
  r.depth = depth;
  r.p_WCb = Apply(X_WB, shape_B_to_point_Ao.p_GN);
  r.nhat_BA_W = shape_B_to_point_Ao.grad_W;
  r.p_WCa = {X_WA[9] - radius_A * r.nhat_BA_W[0], X_WA[10] - radius_A * r.nhat_BA_W[1],
             X_WA[11] - radius_A * r.nhat_BA_W[2]};
  return r;
}

// PenetrationAsPointPair::SwapAAndB (query_results/penetration_as_point_pair.h:26-31).
inline void SwapAAndB(Result* r) {
  std::swap(r->p_WCa, r->p_WCb);
  r->nhat_BA_W = {-r->nhat_BA_W[0], -r->nhat_BA_W[1], -r->nhat_BA_W[2]};
}

// ---------------------------------------------------------------------------------------------------------------------
// FCL's half-space primitives (fcl/narrowphase/detail/primitive_shape_algorithm/halfspace-inl.h at the commit the
// reference pins), as the reference reaches them: CalcDistanceFallback<double> (penetration_as_point_pair_callback.cc:
// 145-207) calls fcl::collide() with one contact requested on a pair (Halfspace, X) or (X, Halfspace), X not a sphere.
// Drake's HalfSpace is fcl::Halfspaced(0, 0, 1, 0) (proximity_engine.cc:667): the solid is z <= 0 of its frame.

// fcl::Contact as CalcDistanceFallback reads it: the position between the two surfaces and the depth. FCL's normal is
// implied: it points from the shape into the half space, i.e. it is -n_W of the half space.
struct FclHalfSpaceContact {
  Vec3 pos;
  double depth;
};

// fcl::detail::halfspaceIntersectTolerance<double>().
inline constexpr double kHalfspaceIntersectTolerance = 0.0000001;

// The half space of frame H posed in World: outward unit normal n_W (the frame's +z) and offset d of n_W . x <= d
// (fcl::transform(Halfspace, tf): n' = R n, d' = d + n' . t).
struct WorldHalfSpace {
  Vec3 n;
  double d;
  double SignedDistance(const Vec3& p) const { return Dot(n, p) - d; }
};
inline WorldHalfSpace HalfSpaceInWorld(const Pose& X_WH) {
  const Vec3 n{X_WH[6], X_WH[7], X_WH[8]};
  return {n, Dot(n, {X_WH[9], X_WH[10], X_WH[11]})};
}
inline Vec3 AddScaled(const Vec3& p, const Vec3& v, double s) { return {p[0] + v[0] * s, p[1] + v[1] * s, p[2] + v[2] * s}; }

// boxHalfspaceIntersect: depth from the box's support function; the deepest point is a face centre when the plane is
// parallel to a face (within the tolerance), else the lowest corner.
inline std::optional<FclHalfSpaceContact> BoxHalfSpace(const Vec3& size, const Pose& X_WB, const Pose& X_WH) {
  const WorldHalfSpace hs = HalfSpaceInWorld(X_WH);
  const Vec3 T{X_WB[9], X_WB[10], X_WB[11]};
  const Vec3 Q = RotateT(X_WB, hs.n);  // the plane normal in the box frame
  const Vec3 A{Q[0] * size[0], Q[1] * size[1], Q[2] * size[2]};
  const double depth = 0.5 * ((std::abs(A[0]) + std::abs(A[1])) + std::abs(A[2])) - hs.SignedDistance(T);
  if (depth < 0) return std::nullopt;
  Vec3 p = T;
  int parallel = -1;
  for (int i = 0; i < 3 && parallel < 0; ++i)
    if (std::abs(Q[i] - 1) < kHalfspaceIntersectTolerance || std::abs(Q[i] + 1) < kHalfspaceIntersectTolerance) parallel = i;
  for (int i = 0; i < 3; ++i) {
    if (parallel >= 0 && i != parallel) continue;
    const int sign = A[i] > 0 ? -1 : 1;
    p = AddScaled(p, {X_WB[3 * i], X_WB[3 * i + 1], X_WB[3 * i + 2]}, 0.5 * size[i] * sign);
  }
  return FclHalfSpaceContact{AddScaled(p, hs.n, depth * 0.5), depth};
}

// capsuleHalfspaceIntersect: axis parallel to the plane -> the point under the centre; else the cap centre at the lower
// end of the spine.
inline std::optional<FclHalfSpaceContact> CapsuleHalfSpace(double radius, double length, const Pose& X_WC, const Pose& X_WH) {
  const WorldHalfSpace hs = HalfSpaceInWorld(X_WH);
  const Vec3 T{X_WC[9], X_WC[10], X_WC[11]};
  const Vec3 dir_z{X_WC[6], X_WC[7], X_WC[8]};
  const double cosa = Dot(dir_z, hs.n);
  if (std::abs(cosa) < kHalfspaceIntersectTolerance) {
    const double depth = radius - hs.SignedDistance(T);
    if (depth < 0) return std::nullopt;
    return FclHalfSpaceContact{AddScaled(T, hs.n, 0.5 * depth - radius), depth};
  }
  const int sign = cosa > 0 ? -1 : 1;
  const Vec3 p = AddScaled(T, dir_z, length * 0.5 * sign);
  const double depth = radius - hs.SignedDistance(p);
  if (depth < 0) return std::nullopt;
  const Vec3 c = AddScaled(p, hs.n, -radius);  // the deepest point
  return FclHalfSpaceContact{AddScaled(c, hs.n, 0.5 * depth), depth};
}

// cylinderHalfspaceIntersect: axis parallel to the plane -> the point under the centre; else the lowest point of the
// lower rim (the disc's centre when the axis is along the normal).
inline std::optional<FclHalfSpaceContact> CylinderHalfSpace(double radius, double length, const Pose& X_WC, const Pose& X_WH) {
  const WorldHalfSpace hs = HalfSpaceInWorld(X_WH);
  const Vec3 T{X_WC[9], X_WC[10], X_WC[11]};
  const Vec3 dir_z{X_WC[6], X_WC[7], X_WC[8]};
  const double cosa = Dot(dir_z, hs.n);
  if (std::abs(cosa) < kHalfspaceIntersectTolerance) {
    const double depth = radius - hs.SignedDistance(T);
    if (depth < 0) return std::nullopt;
    return FclHalfSpaceContact{AddScaled(T, hs.n, 0.5 * depth - radius), depth};
  }
  Vec3 C{dir_z[0] * cosa - hs.n[0], dir_z[1] * cosa - hs.n[1], dir_z[2] * cosa - hs.n[2]};  // -n projected into the disc
  if (std::abs(cosa + 1) < kHalfspaceIntersectTolerance || std::abs(cosa - 1) < kHalfspaceIntersectTolerance) {
    C = {0.0, 0.0, 0.0};
  } else {
    const double s = radius / Norm(C);
    C = {C[0] * s, C[1] * s, C[2] * s};
  }
  const int sign = cosa > 0 ? -1 : 1;
  const Vec3 axis_end = AddScaled(T, dir_z, length * 0.5 * sign);
  const Vec3 p{axis_end[0] + C[0], axis_end[1] + C[1], axis_end[2] + C[2]};  // the deepest point
  const double depth = -hs.SignedDistance(p);
  if (depth < 0) return std::nullopt;
  return FclHalfSpaceContact{AddScaled(p, hs.n, 0.5 * depth), depth};
}

// ellipsoidHalfspaceIntersect: computed in the ellipsoid's frame. With the plane normal n there, the centre is
// sqrt(sum (r_i n_i)^2) above the tangent plane; the reported position lies on the ray from the centre through the
// deepest point, half the depth above it along the normal.
inline std::optional<FclHalfSpaceContact> EllipsoidHalfSpace(const Vec3& radii, const Pose& X_WE, const Pose& X_WH) {
  const WorldHalfSpace hs = HalfSpaceInWorld(X_WH);
  const Vec3 n = RotateT(X_WE, hs.n);
  const double d = hs.d - Dot(hs.n, {X_WE[9], X_WE[10], X_WE[11]});  // n . x <= d in the ellipsoid's frame
  const Vec3 radii2{radii[0] * radii[0], radii[1] * radii[1], radii[2] * radii[2]};
  const double center_to_contact_plane = std::sqrt(Dot({n[0] * n[0], n[1] * n[1], n[2] * n[2]}, radii2));
  const double depth = center_to_contact_plane + d;
  if (!(depth >= 0)) return std::nullopt;
  const double k = 1.0 / center_to_contact_plane;
  const Vec3 support_vector{k * (radii2[0] * n[0]), k * (radii2[1] * n[1]), k * (radii2[2] * n[2])};
  const double f = 0.5 * depth / Dot(n, support_vector) - 1.0;
  return FclHalfSpaceContact{Apply(X_WE, {support_vector[0] * f, support_vector[1] * f, support_vector[2] * f}), depth};
}

// convexHalfspaceIntersect: the first vertex of smallest signed distance. Drake gives FCL the convex hull of a Mesh or
// Convex; a linear function takes its minimum over the hull at a vertex of the mesh itself, so the mesh's own vertices
// give the same depth (and the same vertex except at exact ties). Touching (distance 0) counts as intersecting.
inline std::optional<FclHalfSpaceContact> VerticesHalfSpace(const double* vertices, std::size_t num_vertices, const Pose& X_WC,
                                                            const Pose& X_WH) {
  const WorldHalfSpace hs = HalfSpaceInWorld(X_WH);
  const Vec3 n = RotateT(X_WC, hs.n);
  const double d = hs.d - Dot(hs.n, {X_WC[9], X_WC[10], X_WC[11]});
  double min_signed_distance = std::numeric_limits<double>::max();
  Vec3 deepest{0.0, 0.0, 0.0};
  for (std::size_t v = 0; v < num_vertices; ++v) {
    const Vec3 p{vertices[3 * v], vertices[3 * v + 1], vertices[3 * v + 2]};
    const double signed_distance = Dot(n, p) - d;
    if (signed_distance < min_signed_distance) {
      min_signed_distance = signed_distance;
      deepest = p;
    }
  }
  if (!(min_signed_distance <= 0)) return std::nullopt;
  const double depth = -min_signed_distance;
  return FclHalfSpaceContact{AddScaled(Apply(X_WC, deepest), hs.n, 0.5 * depth), depth};
}

// CalcDistanceFallback<double> (penetration_as_point_pair_callback.cc:164-207) on such a contact. FCL's normal points
// from the first object into the second and Drake's nhat_BA_W is its negative; osculation (depth <= epsilon) is no
// contact; the two witness points sit half the depth either side of FCL's position.
inline std::optional<Result> HalfSpaceContactToPointPair(const std::optional<FclHalfSpaceContact>& contact, const Pose& X_WH,
                                                         bool half_space_is_A) {
  if (!contact) return std::nullopt;
  if (contact->depth <= std::numeric_limits<double>::epsilon()) return std::nullopt;
  const Vec3 n{X_WH[6], X_WH[7], X_WH[8]};
  Result r;
  r.depth = contact->depth;
  r.nhat_BA_W = half_space_is_A ? Vec3{-n[0], -n[1], -n[2]} : n;
  r.p_WCa = AddScaled(contact->pos, r.nhat_BA_W, -0.5 * r.depth);
  r.p_WCb = AddScaled(contact->pos, r.nhat_BA_W, 0.5 * r.depth);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Box x box: FCL's boxBox2 (fcl/narrowphase/detail/primitive_shape_algorithm/box_box-inl.h, the ODE dBoxBox algorithm),
// which is what fcl::collide() runs for two fcl::Boxd, as CalcDistanceFallback<double> reads its deepest contact
// (fcl::collide keeps the contacts of largest depth when fewer are requested than found; Drake requests one).
//   1. 15 separating axes (3 + 3 face normals, 9 edge x edge), each giving the overlap along it; any positive
//      separation ends the test. The axis of smallest overlap gives normal and depth; an edge x edge axis only wins when
//      its overlap is more than 5 % smaller than the best face axis' (fudge_factor 1.05), and parallel edges are skipped.
//   2. edge x edge: the closest points of the two supporting edges; the contact is their midpoint.
//   3. face axis: the face of the other box most anti-parallel to the normal (the incident face) is clipped against
//      the rectangle of the reference face; the deepest of the clipped points is the contact, reported half its depth
//      back towards the reference face.
// Pinned by the reference's BoxPenetrationTest TangentBox1 / TangentBox2 scenes (penetration_as_point_pair_callback_test.cc:
// 914-920: 1e-12) and by an exact 15-axis oracle (oracle/ho_point_pair.py); FCL's choice among EQUALLY deep points (two
// faces resting flat on each other) is not pinned -- the first of the clipped polygon is taken.

struct FclContact {   // normal: unit, from the first shape into the second
  Vec3 normal, pos;
  double depth;
};

namespace box_box_detail {
inline Vec3 Col(const Pose& X, int j) { return {X[3 * j], X[3 * j + 1], X[3 * j + 2]}; }
// Parameters of the closest points pa + alpha ua, pb + beta ub of two lines (ODE's dLineClosestApproach).
inline void LineClosestApproach(const Vec3& pa, const Vec3& ua, const Vec3& pb, const Vec3& ub, double* alpha, double* beta) {
  const Vec3 p{pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  const double uaub = Dot(ua, ub), q1 = Dot(ua, p), q2 = -Dot(ub, p);
  double d = 1 - uaub * uaub;
  if (d <= 0.0001) {
    *alpha = 0;
    *beta = 0;
  } else {
    d = 1.0 / d;
    *alpha = (q1 + uaub * q2) * d;
    *beta = (uaub * q1 + q2) * d;
  }
}
// The quadrilateral q (4 points x, y) clipped by the rectangle |x| <= h[0], |y| <= h[1]: up to 8 points (ODE's
// intersectRectQuad: Sutherland-Hodgman over the four sides in the order -x, +x, -y, +y).
inline int IntersectRectQuad(const double h[2], const double q_in[8], double ret[16]) {
  double buffer[16];
  double* q = const_cast<double*>(q_in);
  double* r = ret;
  int nq = 4, nr = 0;
  for (int dir = 0; dir <= 1; ++dir) {
    for (int sign = -1; sign <= 1; sign += 2) {
      double* pq = q;
      double* pr = r;
      nr = 0;
      for (int i = nq; i > 0; --i) {
        if (sign * pq[dir] < h[dir]) {   // this point is inside the chopping line
          pr[0] = pq[0];
          pr[1] = pq[1];
          pr += 2;
          ++nr;
          if (nr & 8) {
            q = r;
            goto done;
          }
        }
        double* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {   // this line crosses the chopping line
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          pr[dir] = sign * h[dir];
          pr += 2;
          ++nr;
          if (nr & 8) {
            q = r;
            goto done;
          }
        }
        pq += 2;
      }
      q = r;
      r = (q == ret) ? buffer : ret;
      nq = nr;
    }
  }
done:
  if (q != ret)
    for (int i = 0; i < 2 * nr; ++i) ret[i] = q[i];
  return nr;
}
}  // namespace box_box_detail

inline std::optional<FclContact> BoxBox(const Vec3& side1, const Pose& X_W1, const Vec3& side2, const Pose& X_W2) {
  using namespace box_box_detail;
  const double fudge_factor = 1.05;
  const Vec3 T1{X_W1[9], X_W1[10], X_W1[11]}, T2{X_W2[9], X_W2[10], X_W2[11]};
  const Vec3 p{T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
  const Vec3 pp = RotateT(X_W1, p);
  const double A[3] = {side1[0] * 0.5, side1[1] * 0.5, side1[2] * 0.5}, B[3] = {side2[0] * 0.5, side2[1] * 0.5, side2[2] * 0.5};
  double R[3][3], Q[3][3];   // R = R1^T R2
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      R[i][j] = Dot(Col(X_W1, i), Col(X_W2, j));
      Q[i][j] = std::abs(R[i][j]);
    }
  double s = -std::numeric_limits<double>::max();
  int code = 0;
  bool invert_normal = false;
  Vec3 normal_face{0, 0, 0}, normalC{0, 0, 0};
  bool normal_is_face = true;
  for (int i = 0; i < 3; ++i) {   // u1, u2, u3: the face normals of box 1
    const double tmp = pp[i];
    const double s2 = std::abs(tmp) - (((Q[i][0] * B[0] + Q[i][1] * B[1]) + Q[i][2] * B[2]) + A[i]);
    if (s2 > 0) return std::nullopt;
    if (s2 > s) {
      s = s2;
      normal_face = Col(X_W1, i);
      normal_is_face = true;
      invert_normal = tmp < 0;
      code = 1 + i;
    }
  }
  for (int i = 0; i < 3; ++i) {   // v1, v2, v3: the face normals of box 2
    const double tmp = Dot(Col(X_W2, i), p);
    const double s2 = std::abs(tmp) - (((Q[0][i] * A[0] + Q[1][i] * A[1]) + Q[2][i] * A[2]) + B[i]);
    if (s2 > 0) return std::nullopt;
    if (s2 > s) {
      s = s2;
      normal_face = Col(X_W2, i);
      normal_is_face = true;
      invert_normal = tmp < 0;
      code = 4 + i;
    }
  }
  const double fudge2 = 1.0e-6, eps = std::numeric_limits<double>::epsilon();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Q[i][j] += fudge2;
  for (int i = 0; i < 3; ++i) {   // u_i x v_j
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double tmp = pp[i2] * R[i1][j] - pp[i1] * R[i2][j];
      const double s2_raw = std::abs(tmp) - (((A[i1] * Q[i2][j] + A[i2] * Q[i1][j]) + B[j1] * Q[i][j2]) + B[j2] * Q[i][j1]);
      if (s2_raw > 0) return std::nullopt;
      double n[3] = {0, 0, 0};
      n[i1] = -R[i2][j];
      n[i2] = R[i1][j];
      const double l = std::sqrt(n[i1] * n[i1] + n[i2] * n[i2]);
      if (l > eps) {
        const double s2 = s2_raw / l;
        if (s2 * fudge_factor > s) {
          s = s2;
          normalC = {n[0] / l, n[1] / l, n[2] / l};
          normal_is_face = false;
          invert_normal = tmp < 0;
          code = 7 + 3 * i + j;
        }
      }
    }
  }
  if (!code) return std::nullopt;
  Vec3 normal = normal_is_face ? normal_face : Rotate(X_W1, normalC);
  if (invert_normal) normal = {-normal[0], -normal[1], -normal[2]};
  const double depth = -s;

  if (code > 6) {   // edge x edge
    Vec3 pa = T1, pb = T2;
    for (int j = 0; j < 3; ++j) {
      pa = AddScaled(pa, Col(X_W1, j), A[j] * (Dot(Col(X_W1, j), normal) > 0 ? 1.0 : -1.0));
      pb = AddScaled(pb, Col(X_W2, j), B[j] * (Dot(Col(X_W2, j), normal) > 0 ? -1.0 : 1.0));
    }
    const Vec3 ua = Col(X_W1, (code - 7) / 3), ub = Col(X_W2, (code - 7) % 3);
    double alpha, beta;
    LineClosestApproach(pa, ua, pb, ub, &alpha, &beta);
    pa = AddScaled(pa, ua, alpha);
    pb = AddScaled(pb, ub, beta);
    return FclContact{normal, {0.5 * (pa[0] + pb[0]), 0.5 * (pa[1] + pb[1]), 0.5 * (pa[2] + pb[2])}, depth};
  }

  // A face of box a (the reference face, perpendicular to the normal) against the closest face of box b (the incident one).
  const bool first = code <= 3;
  const Pose& Xa = first ? X_W1 : X_W2;
  const Pose& Xb = first ? X_W2 : X_W1;
  const Vec3& pa = first ? T1 : T2;
  const Vec3& pb = first ? T2 : T1;
  const double* Sa = first ? A : B;
  const double* Sb = first ? B : A;
  const Vec3 normal2 = first ? normal : Vec3{-normal[0], -normal[1], -normal[2]};
  const Vec3 nr = RotateT(Xb, normal2);
  const double anr[3] = {std::abs(nr[0]), std::abs(nr[1]), std::abs(nr[2])};
  int lanr, a1, a2;   // the incident face's normal axis and its two in-plane axes
  if (anr[1] > anr[0]) {
    if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  // Centre of the incident face relative to the centre of box a.
  const double side = nr[lanr] < 0 ? Sb[lanr] : -Sb[lanr];
  const Vec3 center = AddScaled({pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, Col(Xb, lanr), side);
  const int codeN = first ? code - 1 : code - 4;
  const int code1 = codeN == 0 ? 1 : 0, code2 = codeN == 2 ? 1 : 2;
  // The incident face's corners in the reference face's 2-D coordinates.
  const double c1 = Dot(Col(Xa, code1), center), c2 = Dot(Col(Xa, code2), center);
  double m11 = Dot(Col(Xb, a1), Col(Xa, code1)), m12 = Dot(Col(Xb, a2), Col(Xa, code1));
  double m21 = Dot(Col(Xb, a1), Col(Xa, code2)), m22 = Dot(Col(Xb, a2), Col(Xa, code2));
  {
    const double k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
    const double quad[8] = {c1 - k1 - k3, c2 - k2 - k4, c1 - k1 + k3, c2 - k2 + k4, c1 + k1 + k3, c2 + k2 + k4, c1 + k1 - k3, c2 + k2 - k4};
    const double rect[2] = {Sa[code1], Sa[code2]};
    double ret[16];
    const int n_intersect = IntersectRectQuad(rect, quad, ret);
    if (n_intersect < 1) return std::nullopt;   // "this should never happen"
    const double det1 = 1.0 / (m11 * m22 - m12 * m21);
    m11 *= det1;
    m12 *= det1;
    m21 *= det1;
    m22 *= det1;
    bool found = false;
    Vec3 best_point{0, 0, 0};
    double best_dep = 0;
    for (int j = 0; j < n_intersect; ++j) {
      const double k1b = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
      const double k2b = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
      const Vec3 point = AddScaled(AddScaled(center, Col(Xb, a1), k1b), Col(Xb, a2), k2b);   // relative to box a's centre
      const double dep = Sa[codeN] - Dot(normal2, point);
      if (dep >= 0 && (!found || dep > best_dep)) {
        found = true;
        best_dep = dep;
        best_point = point;
      }
    }
    if (!found) return std::nullopt;
    // The deepest point of the incident face, moved half its depth back along the reference face's normal.
    const Vec3 on_b{best_point[0] + pa[0], best_point[1] + pa[1], best_point[2] + pa[2]};
    return FclContact{normal, AddScaled(on_b, normal2, 0.5 * best_dep), best_dep};
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Every other convex pair: Minkowski Portal Refinement. fcl::collide() with GST_LIBCCD (proximity_engine.cc:698,742)
// computes the contact of two convex shapes without a primitive algorithm through libccd's ccdMPRPenetration (G. Snethen,
// "XenoCollide", Game Programming Gems 7; libccd src/mpr.c) over per-shape support functions, tolerance 1e-6, at most
// 500 refinements. Restated here from the published algorithm:
//   * the Minkowski difference D = A - B contains the origin iff the shapes overlap; v0 = centre(A) - centre(B) is inside D;
//   * discover a portal (a triangle of support points of D that the ray from v0 through the origin passes), refine it
//     outwards until the origin is behind it, then keep expanding towards the surface until the new support point is
//     within the tolerance of the portal's plane;
//   * depth = distance of the origin to the final portal triangle, normal = the direction of its nearest point, position =
//     the barycentric mix of the support points' witnesses on A and on B, averaged.
// What MPR reports is the penetration seen along the ray from v0, not the minimum over all directions: for shallow contacts
// the two coincide, for deep or oblique ones they do not -- the reference's own BoxPenetrationTest accepts 1e-1 ... 1e-4 on
// these pairs for that reason (Drake issue 7656). Pinned by that test's scenes for cylinder, capsule and convex tangent
// shapes at the tolerances it states (tests/test_point_pair_cpu.py); libccd's arithmetic is not reproduced bit for bit,
// and with it the path the refinement takes: at libccd's own tolerance (1e-6) this restatement meets nine of the ten
// scenes (TangentProneCapsule2: 3.3e-4 where the reference states 1e-4), at 1e-7 -- the default here -- all ten.

struct ConvexShape {
  enum Kind { kSphere, kBox, kCylinder, kCapsule, kEllipsoid, kVertices } kind;
  Vec3 dims;                    // radius | full side lengths | radius, length | radius, length | semi-axes
  const double* vertices;       // kVertices: [num_vertices * 3] in the shape's frame
  std::size_t num_vertices;
  Pose X_WS;
};

namespace mpr_detail {
inline Vec3 Sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline Vec3 Cross(const Vec3& a, const Vec3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline Vec3 Scale(const Vec3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }
inline constexpr double kEps = std::numeric_limits<double>::epsilon();
inline bool IsZero(double x) { return std::abs(x) < kEps; }
inline double SignOrZero(double x) { return IsZero(x) ? 0.0 : (x < 0 ? -1.0 : 1.0); }

// The shape's support point in World for a World direction (FCL's supportBox / supportCyl / supportCap / supportSphere /
// supportEllipsoid / supportConvex, gjk_libccd-inl.h): the farthest point of the shape along dir.
inline Vec3 Support(const ConvexShape& sh, const Vec3& dir_W) {
  const Vec3 d = RotateT(sh.X_WS, dir_W);
  Vec3 v{0, 0, 0};
  switch (sh.kind) {
    case ConvexShape::kSphere: {
      const double n = Norm(d);
      v = Scale(d, sh.dims[0] / n);
      break;
    }
    case ConvexShape::kBox:
      v = {SignOrZero(d[0]) * sh.dims[0] * 0.5, SignOrZero(d[1]) * sh.dims[1] * 0.5, SignOrZero(d[2]) * sh.dims[2] * 0.5};
      break;
    case ConvexShape::kCylinder: {
      const double zdist = std::sqrt(d[0] * d[0] + d[1] * d[1]);
      if (IsZero(zdist)) {
        v = {0, 0, SignOrZero(d[2]) * sh.dims[1] * 0.5};
      } else {
        const double rad = sh.dims[0] / zdist;
        v = {rad * d[0], rad * d[1], SignOrZero(d[2]) * sh.dims[1] * 0.5};
      }
      break;
    }
    case ConvexShape::kCapsule: {
      const double n = Norm(d);
      v = Scale(d, sh.dims[0] / n);
      v[2] += d[2] > 0 ? sh.dims[1] * 0.5 : -sh.dims[1] * 0.5;
      break;
    }
    case ConvexShape::kEllipsoid: {
      const Vec3 abc2{sh.dims[0] * sh.dims[0] * d[0], sh.dims[1] * sh.dims[1] * d[1], sh.dims[2] * sh.dims[2] * d[2]};
      v = Scale(abc2, 1.0 / std::sqrt(Dot(d, abc2)));
      break;
    }
    case ConvexShape::kVertices: {
      double best = -std::numeric_limits<double>::max();
      for (std::size_t k = 0; k < sh.num_vertices; ++k) {
        const Vec3 p{sh.vertices[3 * k], sh.vertices[3 * k + 1], sh.vertices[3 * k + 2]};
        const double t = Dot(d, p);
        if (t > best) {
          best = t;
          v = p;
        }
      }
      break;
    }
  }
  return Apply(sh.X_WS, v);
}
// An interior point: the frame origin (a vertex cloud: the mean of its vertices).
inline Vec3 Center(const ConvexShape& sh) {
  if (sh.kind != ConvexShape::kVertices) return {sh.X_WS[9], sh.X_WS[10], sh.X_WS[11]};
  Vec3 c{0, 0, 0};
  for (std::size_t k = 0; k < sh.num_vertices; ++k) c = {c[0] + sh.vertices[3 * k], c[1] + sh.vertices[3 * k + 1], c[2] + sh.vertices[3 * k + 2]};
  return Apply(sh.X_WS, Scale(c, 1.0 / static_cast<double>(sh.num_vertices)));
}

struct SupportPoint {   // a point of A - B with its witnesses
  Vec3 v, v1, v2;
};
inline SupportPoint MinkowskiSupport(const ConvexShape& a, const ConvexShape& b, const Vec3& dir) {
  SupportPoint s;
  s.v1 = Support(a, dir);
  s.v2 = Support(b, Scale(dir, -1.0));
  s.v = Sub(s.v1, s.v2);
  return s;
}
inline bool Normalize(Vec3* v) {
  const double n = Norm(*v);
  if (n == 0) return false;
  *v = Scale(*v, 1.0 / n);
  return true;
}
inline Vec3 PortalDir(const SupportPoint p[4]) {
  Vec3 dir = Cross(Sub(p[2].v, p[1].v), Sub(p[3].v, p[1].v));
  Normalize(&dir);
  return dir;
}
// Squared distance of the origin to the triangle (a, b, c) and the nearest point (libccd's ccdVec3PointTriDist2 reduces to
// this): the projection when it falls inside, else the nearest point of the three edges.
inline double OriginTriangleDist2(const Vec3& a, const Vec3& b, const Vec3& c, Vec3* witness) {
  const Vec3 d1 = Sub(b, a), d2 = Sub(c, a);
  const double v = Dot(d1, d1), w = Dot(d2, d2), p = Dot(a, d1), q = Dot(a, d2), r = Dot(d1, d2);
  const double div = w * v - r * r;
  double s = -1, t = -1;
  if (!IsZero(div)) {
    s = (q * r - w * p) / div;
    t = (-s * r - q) / w;
  }
  if (!IsZero(div) && (IsZero(s) || s > 0) && (s < 1 || IsZero(s - 1)) && (IsZero(t) || t > 0) && (t < 1 || IsZero(t - 1)) &&
      (t + s < 1 || IsZero(t + s - 1))) {
    *witness = {a[0] + s * d1[0] + t * d2[0], a[1] + s * d1[1] + t * d2[1], a[2] + s * d1[2] + t * d2[2]};
    return Dot(*witness, *witness);
  }
  auto segment = [](const Vec3& x0, const Vec3& x1, Vec3* wit) {
    const Vec3 e = Sub(x1, x0);
    double k = -Dot(x0, e) / Dot(e, e);
    k = k < 0 ? 0 : (k > 1 ? 1 : k);
    *wit = {x0[0] + k * e[0], x0[1] + k * e[1], x0[2] + k * e[2]};
    return Dot(*wit, *wit);
  };
  Vec3 w2;
  double dist = segment(a, b, witness);
  double dist2 = segment(a, c, &w2);
  if (dist2 < dist) {
    dist = dist2;
    *witness = w2;
  }
  dist2 = segment(b, c, &w2);
  if (dist2 < dist) {
    dist = dist2;
    *witness = w2;
  }
  return dist;
}
}  // namespace mpr_detail

// ccdMPRPenetration as fcl::collide() reads it (normal from A into B). nullopt: the shapes do not overlap.
inline std::optional<FclContact> MprPenetration(const ConvexShape& a, const ConvexShape& b, double tolerance = 1e-7,
                                                int max_iterations = 500) {
  using namespace mpr_detail;
  SupportPoint p[4];   // p[0]: the interior point v0; p[1..3]: the portal
  // ---- discover a portal ----
  p[0].v1 = Center(a);
  p[0].v2 = Center(b);
  p[0].v = Sub(p[0].v1, p[0].v2);
  if (IsZero(p[0].v[0]) && IsZero(p[0].v[1]) && IsZero(p[0].v[2])) p[0].v = {kEps * 10, 0, 0};   // coincident centres
  Vec3 dir = Scale(p[0].v, -1.0);
  Normalize(&dir);
  p[1] = MinkowskiSupport(a, b, dir);
  double dot = Dot(p[1].v, dir);
  if (IsZero(dot) || dot < 0) return std::nullopt;
  dir = Cross(p[0].v, p[1].v);
  int found = 0;   // 0: a proper portal; 1: the origin is p[1] (touching); 2: the origin lies on the segment v0 - p[1]
  if (IsZero(Norm(dir))) {
    found = (IsZero(p[1].v[0]) && IsZero(p[1].v[1]) && IsZero(p[1].v[2])) ? 1 : 2;
  } else {
    Normalize(&dir);
    p[2] = MinkowskiSupport(a, b, dir);
    dot = Dot(p[2].v, dir);
    if (IsZero(dot) || dot < 0) return std::nullopt;
    dir = Cross(Sub(p[1].v, p[0].v), Sub(p[2].v, p[0].v));
    Normalize(&dir);
    if (Dot(dir, p[0].v) > 0) {   // the portal must face away from v0
      std::swap(p[1], p[2]);
      dir = Scale(dir, -1.0);
    }
    for (int guard = 0;; ++guard) {
      if (guard > max_iterations) return std::nullopt;
      p[3] = MinkowskiSupport(a, b, dir);
      dot = Dot(p[3].v, dir);
      if (IsZero(dot) || dot < 0) return std::nullopt;
      bool cont = false;
      // The origin must lie inside the cone (v0; p1, p2, p3): else replace the offending corner.
      if (Dot(Cross(p[1].v, p[3].v), p[0].v) < -kEps) {
        p[2] = p[3];
        cont = true;
      } else if (Dot(Cross(p[3].v, p[2].v), p[0].v) < -kEps) {
        p[1] = p[3];
        cont = true;
      }
      if (!cont) break;
      dir = Cross(Sub(p[1].v, p[0].v), Sub(p[2].v, p[0].v));
      Normalize(&dir);
    }
  }
  auto midpoint = [](const Vec3& x, const Vec3& y) { return Vec3{0.5 * (x[0] + y[0]), 0.5 * (x[1] + y[1]), 0.5 * (x[2] + y[2])}; };
  if (found == 1) return FclContact{{0, 0, 0}, midpoint(p[1].v1, p[1].v2), 0.0};   // touching: depth 0 (Drake drops it)
  if (found == 2) {   // the origin on the segment v0 - p1: the penetration is p1 itself
    Vec3 n = p[1].v;
    const double depth = Norm(n);
    Normalize(&n);
    return FclContact{n, midpoint(p[1].v1, p[1].v2), depth};
  }
  // ---- refine the portal until the origin is behind it ----
  for (int guard = 0;; ++guard) {
    dir = PortalDir(p);
    dot = Dot(p[1].v, dir);
    if (IsZero(dot) || dot > 0) break;   // the portal encapsulates the origin
    const SupportPoint v4 = MinkowskiSupport(a, b, dir);
    const double d4 = Dot(v4.v, dir);
    if (!(IsZero(d4) || d4 > 0)) return std::nullopt;   // the origin is outside A - B
    const double reach = std::min({d4 - Dot(p[1].v, dir), d4 - Dot(p[2].v, dir), d4 - Dot(p[3].v, dir)});
    if (reach < tolerance || IsZero(reach - tolerance) || guard > max_iterations) return std::nullopt;
    // expand the portal by v4
    const Vec3 v4v0 = Cross(v4.v, p[0].v);
    if (Dot(p[1].v, v4v0) > 0) {
      if (Dot(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4;
    } else {
      if (Dot(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4;
    }
  }
  // ---- push the portal to the surface ----
  for (int iterations = 0;; ++iterations) {
    dir = PortalDir(p);
    const SupportPoint v4 = MinkowskiSupport(a, b, dir);
    const double d4 = Dot(v4.v, dir);
    const double reach = std::min({d4 - Dot(p[1].v, dir), d4 - Dot(p[2].v, dir), d4 - Dot(p[3].v, dir)});
    if (reach < tolerance || IsZero(reach - tolerance) || iterations > max_iterations) {
      Vec3 pdir;
      const double depth = std::sqrt(OriginTriangleDist2(p[1].v, p[2].v, p[3].v, &pdir));
      if (IsZero(pdir[0]) && IsZero(pdir[1]) && IsZero(pdir[2])) pdir = dir;
      Normalize(&pdir);
      // position: barycentric coordinates of the origin's ray in the portal, applied to the witnesses
      double bc[4] = {Dot(Cross(p[1].v, p[2].v), p[3].v), Dot(Cross(p[3].v, p[2].v), p[0].v),
                      Dot(Cross(p[0].v, p[1].v), p[3].v), Dot(Cross(p[2].v, p[1].v), p[0].v)};
      double sum = ((bc[0] + bc[1]) + bc[2]) + bc[3];
      if (IsZero(sum) || sum < 0) {
        bc[0] = 0;
        bc[1] = Dot(Cross(p[2].v, p[3].v), dir);
        bc[2] = Dot(Cross(p[3].v, p[1].v), dir);
        bc[3] = Dot(Cross(p[1].v, p[2].v), dir);
        sum = (bc[1] + bc[2]) + bc[3];
      }
      const double inv = 1.0 / sum;
      Vec3 p1{0, 0, 0}, p2{0, 0, 0};
      for (int k = 0; k < 4; ++k) {
        p1 = AddScaled(p1, p[k].v1, bc[k]);
        p2 = AddScaled(p2, p[k].v2, bc[k]);
      }
      return FclContact{pdir, Scale(midpoint(p1, p2), inv), depth};
    }
    const Vec3 v4v0 = Cross(v4.v, p[0].v);
    if (Dot(p[1].v, v4v0) > 0) {
      if (Dot(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4;
    } else {
      if (Dot(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4;
    }
  }
}

// CalcDistanceFallback<double> (penetration_as_point_pair_callback.cc:164-207) on a contact whose normal points from A
// into B.
inline std::optional<Result> FclContactToPointPair(const std::optional<FclContact>& contact) {
  if (!contact) return std::nullopt;
  if (contact->depth <= std::numeric_limits<double>::epsilon()) return std::nullopt;
  Result r;
  r.depth = contact->depth;
  r.nhat_BA_W = {-contact->normal[0], -contact->normal[1], -contact->normal[2]};
  r.p_WCa = AddScaled(contact->pos, r.nhat_BA_W, -0.5 * r.depth);
  r.p_WCb = AddScaled(contact->pos, r.nhat_BA_W, 0.5 * r.depth);
  return r;
}

}  // namespace point_pair
}  // namespace drake_b200
#endif  // DRAKE_B200_POINT_PAIR_H_
