// The HalfSpace shape through the C++ mirror of the reference interface, the way the reference's tests
// dispatch it (hydroelastic_calculator_test.cc:373-460 CompliantMeshRigidHalfSpace /
// CompliantHalfSpaceRigidMesh, :576-631 the rejected combinations; proximity_engine.cc:1361-1390).
// Exit code 0 = all checks passed; 3 = no CUDA device (the library has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "drake_b200_proximity_engine.h"

using namespace drake_b200;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d: EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

template <class F>
static bool ThrowsLogicError(F&& f, const char* needle) {
  try {
    f();
  } catch (const std::logic_error& err) {
    return std::string(err.what()).find(needle) != std::string::npos;
  }
  return false;
}

int main() {
  try {
    ProximityEngine probe(0);
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    return 3;
  }
  const double R = 0.25;
  ProximityProperties soft;  // compliant_properties() of the reference test: hint 0.25, E = 1e8, slab 0.25
  soft.compliance_type = HydroelasticType::kCompliant;
  soft.hydroelastic_modulus = 1e8;
  soft.resolution_hint = 0.25;
  soft.slab_thickness = 0.25;
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  rigid.resolution_hint = 0.1;
  const auto kTri = HydroelasticContactRepresentation::kTriangle;
  const auto kPoly = HydroelasticContactRepresentation::kPolygon;

  {  // compliant sphere (id 1) x rigid half space (id 2): colliding, then separated
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, R / 2), 1, soft);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 2, rigid);
    for (auto rep : {kTri, kPoly}) {
      auto s = e.ComputeContactSurfaces(rep, {});
      EXPECT(s.size() == 1 && s[0].num_faces() > 0);
      EXPECT(s[0].id_M() == 1 && s[0].id_N() == 2 && s[0].HasGradE_M() && !s[0].HasGradE_N());
      // The boundary plane z = 0 cuts the ball of radius R centred at R / 2: a disc-like patch whose
      // faces all point up (out of the half space N into the ball M).
      double sum = 0;
      for (int f = 0; f < s[0].num_faces(); ++f) sum += s[0].area(f);
      EXPECT(std::fabs(sum - s[0].total_area()) < 1e-12);
      EXPECT(s[0].total_area() > 0.05 && s[0].total_area() < M_PI * R * R);
      // Welded like the reference's (cut_edges, mesh_plane_intersection.cc:205-231): the polygons share the vertices
      // cut from the same tetrahedron edge, so there are far fewer vertices than face corners.
      const int corners = s[0].is_triangle() ? 3 * s[0].num_faces()
                                             : static_cast<int>(s[0].face_data().size()) - s[0].num_faces();
      EXPECT(s[0].num_vertices() < corners * 2 / 3);
      for (int32_t idx : s[0].face_data()) EXPECT(idx >= 0 && idx < std::max(s[0].num_vertices(), 8));
    }
    EXPECT(e.ComputeContactSurfaces(kPoly, {{1, RigidTransformd::FromTranslation(0, 0, 1.5 * R)}}).empty());
  }
  {  // rigid sphere (id 1) x compliant half space (id 2): M is the rigid mesh after the id swap
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, R / 2), 1, rigid);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 2, soft);
    for (auto rep : {kTri, kPoly}) {
      auto s = e.ComputeContactSurfaces(rep, {});
      EXPECT(s.size() == 1 && s[0].num_faces() > 0);
      EXPECT(s[0].id_M() == 1 && s[0].id_N() == 2 && !s[0].HasGradE_M() && s[0].HasGradE_N());
    }
    EXPECT(e.ComputeContactSurfaces(kPoly, {{1, RigidTransformd::FromTranslation(0, 0, 1.5 * R)}}).empty());
  }
  {  // compliant half space x compliant mesh; half space x half space; rigid half space x rigid mesh
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, R / 2), 1, soft);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 2, soft);
    EXPECT(ThrowsLogicError([&] { e.ComputeContactSurfaces(kPoly, {}); }, "one of which is a half space"));
    std::vector<ContactSurface> surfaces;
    std::vector<PenetrationAsPointPair> pairs;
    e.ComputeContactSurfacesWithFallback(kPoly, {}, &surfaces, &pairs);
    EXPECT(surfaces.empty());
  }
  {
    ProximityEngine e(0);
    e.AddDynamicGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
    e.AddDynamicGeometry(HalfSpace{}, RigidTransformd::Identity(), 2, soft);
    EXPECT(ThrowsLogicError([&] { e.ComputeContactSurfaces(kPoly, {}); }, "two half spaces"));
  }
  {
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, R / 2), 1, rigid);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 2, rigid);
    EXPECT(ThrowsLogicError([&] { e.ComputeContactSurfaces(kPoly, {}); }, "rigid-rigid"));
  }
  std::printf("ok\n");
  return 0;
}
