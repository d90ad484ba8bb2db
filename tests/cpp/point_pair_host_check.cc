// Host check of include/drake_b200_point_pair.h (no GPU, no library): reads pairs from stdin, one per line
//   <kindA> <paramsA...> <poseA x12> <kindB> <paramsB...> <poseB x12>
//   kind: s r | b x y z | h | c r L | k r L | e a b c | v n x0 y0 z0 ... (n vertices)
// and prints `none` or `depth p_WCa(3) p_WCb(3) nhat(3)` with 17 significant digits, the way
// ComputeNarrowPhasePenetration (penetration_as_point_pair_callback.cc:203-300) orders its operands; a pair without a
// sphere must hold exactly one half space (FCL's closed-form half-space primitives behind CalcDistanceFallback) or be
// two boxes (FCL's box-box); every other pair without a half space goes through MprPenetration.
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "drake_b200_point_pair.h"

using namespace drake_b200::point_pair;

struct Geo {
  char kind;
  double r;
  Vec3 size;
  std::vector<double> vertices;
  Pose X;
};

static bool Read(Geo* g) {
  std::string k;
  if (!(std::cin >> k)) return false;
  g->kind = k[0];
  if (g->kind == 's') std::cin >> g->r;
  if (g->kind == 'b' || g->kind == 'e') std::cin >> g->size[0] >> g->size[1] >> g->size[2];
  if (g->kind == 'v') {
    int n = 0;
    std::cin >> n;
    g->vertices.resize(3 * n);
    for (double& v : g->vertices) std::cin >> v;
  }
  if (g->kind == 'c' || g->kind == 'k') std::cin >> g->size[0] >> g->size[1];   // cylinder / capsule: radius, length
  for (double& v : g->X) std::cin >> v;
  return true;
}

int main() {
  Geo a, b;
  while (Read(&a) && Read(&b)) {
    const bool a_is_sphere = a.kind == 's';
    if (a.kind == 'b' && b.kind == 'b') {   // FCL's box-box behind CalcDistanceFallback
      const auto res = FclContactToPointPair(BoxBox(a.size, a.X, b.size, b.X));
      if (!res) {
        std::printf("none\n");
        continue;
      }
      std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", res->depth, res->p_WCa[0], res->p_WCa[1],
                  res->p_WCa[2], res->p_WCb[0], res->p_WCb[1], res->p_WCb[2], res->nhat_BA_W[0], res->nhat_BA_W[1],
                  res->nhat_BA_W[2]);
      continue;
    }
    // Neither Drake's closed forms (a sphere against s b h c k) nor a half space: libccd's MPR.
    const auto drake_has = [](char k) { return k == 's' || k == 'b' || k == 'h' || k == 'c' || k == 'k'; };
    const bool drake_closed_form = (a_is_sphere && drake_has(b.kind)) || (b.kind == 's' && drake_has(a.kind));
    if (a.kind != 'h' && b.kind != 'h' && !drake_closed_form) {
      auto shape = [](const Geo& g) {
        ConvexShape sh{};
        sh.kind = g.kind == 's' ? ConvexShape::kSphere : g.kind == 'b' ? ConvexShape::kBox : g.kind == 'c' ? ConvexShape::kCylinder
                : g.kind == 'k' ? ConvexShape::kCapsule : g.kind == 'e' ? ConvexShape::kEllipsoid : ConvexShape::kVertices;
        sh.dims = g.kind == 's' ? Vec3{g.r, 0, 0} : g.size;
        sh.vertices = g.vertices.data();
        sh.num_vertices = g.vertices.size() / 3;
        sh.X_WS = g.X;
        return sh;
      };
      const auto res = FclContactToPointPair(MprPenetration(shape(a), shape(b)));
      if (!res) {
        std::printf("none\n");
        continue;
      }
      std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", res->depth, res->p_WCa[0], res->p_WCa[1],
                  res->p_WCa[2], res->p_WCb[0], res->p_WCb[1], res->p_WCb[2], res->nhat_BA_W[0], res->nhat_BA_W[1],
                  res->nhat_BA_W[2]);
      continue;
    }
    if (!a_is_sphere && b.kind != 's') {
      const bool a_is_half_space = a.kind == 'h';
      const Geo& h = a_is_half_space ? a : b;
      const Geo& x = a_is_half_space ? b : a;
      std::optional<FclHalfSpaceContact> c;
      if (x.kind == 'b') c = BoxHalfSpace(x.size, x.X, h.X);
      else if (x.kind == 'c') c = CylinderHalfSpace(x.size[0], x.size[1], x.X, h.X);
      else if (x.kind == 'k') c = CapsuleHalfSpace(x.size[0], x.size[1], x.X, h.X);
      else if (x.kind == 'e') c = EllipsoidHalfSpace(x.size, x.X, h.X);
      else c = VerticesHalfSpace(x.vertices.data(), x.vertices.size() / 3, x.X, h.X);
      const auto res = HalfSpaceContactToPointPair(c, h.X, a_is_half_space);
      if (!res) {
        std::printf("none\n");
        continue;
      }
      std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", res->depth, res->p_WCa[0], res->p_WCa[1],
                  res->p_WCa[2], res->p_WCb[0], res->p_WCb[1], res->p_WCb[2], res->nhat_BA_W[0], res->nhat_BA_W[1],
                  res->nhat_BA_W[2]);
      continue;
    }
    const Geo& s = a_is_sphere ? a : b;
    const Geo& o = a_is_sphere ? b : a;
    const Vec3 p_WSo{s.X[9], s.X[10], s.X[11]};
    SignedDistanceToPoint d;
    if (o.kind == 's') d = DistanceToSphere(o.r, o.X, p_WSo);
    else if (o.kind == 'b') d = DistanceToBox(o.size, o.X, p_WSo);
    else if (o.kind == 'c') d = DistanceToCylinder(o.size[0], o.size[1], o.X, p_WSo);
    else if (o.kind == 'k') d = DistanceToCapsule(o.size[0], o.size[1], o.X, p_WSo);
    else d = DistanceToHalfSpace(o.X, p_WSo);
    auto res = SphereShapePenetration(s.r, s.X, o.X, d);
    if (!res) {
      std::printf("none\n");
      continue;
    }
    if (!a_is_sphere) SwapAAndB(&*res);
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", res->depth, res->p_WCa[0], res->p_WCa[1],
                res->p_WCa[2], res->p_WCb[0], res->p_WCb[1], res->p_WCb[2], res->nhat_BA_W[0], res->nhat_BA_W[1],
                res->nhat_BA_W[2]);
  }
  return 0;
}
