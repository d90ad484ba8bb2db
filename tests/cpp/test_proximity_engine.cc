// Exercises the C++ mirror of the reference interface (include/drake_b200_proximity_engine.h) the way
// the reference's own tests exercise ProximityEngine / ContactCalculator:
//   hydroelastic_calculator_test.cc:320-355  compliant sphere x rigid sphere  -> 100 triangles / 28 polygons
//   hydroelastic_calculator_test.cc:467-504  compliant x compliant            ->  12 triangles /  4 polygons
//   proximity_engine_test.cc:1956-2249       ordering, rigid-rigid throws, WithFallback appends
// Exit code 0 = all checks passed; 3 = no CUDA device (the library has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "drake_b200_proximity_engine.h"

using namespace drake_b200;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d: EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

static RigidTransformd RotX(double angle, double x, double y, double z) {
  RigidTransformd X;
  const double c = std::cos(angle), s = std::sin(angle);
  // column-major: columns (1,0,0), (0,c,s), (0,-s,c)
  X.m = {1, 0, 0, 0, c, s, 0, -s, c, x, y, z};
  return X;
}

// `margin <faces> <total area> <sum of face pressures> <faces> <area> <pressures>`: a compliant Sphere and a compliant Box
// registered with margin > 0 against a rigid sphere; the expected values come from the oracle's construction on the
// INFLATED shapes (hydroelastic_internal.cc:412-465), computed by tests/test_cpp_adapter.py.
static int MarginCase(char** argv) {
  const double R = 0.25, margin = 0.01;
  ProximityProperties soft;
  soft.compliance_type = HydroelasticType::kCompliant;
  soft.hydroelastic_modulus = 1e8;
  soft.resolution_hint = 0.25;
  soft.margin = margin;
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  rigid.resolution_hint = 0.1;
  const RigidTransformd X_WA = RotX(M_PI, 0, 0, 0);
  const RigidTransformd X_WB = RigidTransformd::FromTranslation(0, 0, -(R + R) + 0.125);
  // Cases 2-4: Capsule and Cylinder get the reference's representations too (MakeCapsuleVolumeMesh /
  // MakeCylinderVolumeMeshWithMa on the inflated shape, hydroelastic_internal.cc:465-508; rigid: the boundary of the
  // same mesh, :344-363).
  for (int k = 0; k < 5; ++k) {
    ProximityEngine e(0);
    if (k == 0 || k == 4) e.AddDynamicGeometry(Sphere{R}, X_WA, 1, soft);
    else if (k == 1) e.AddDynamicGeometry(Box{0.5, 0.4, 0.5}, X_WA, 1, soft);
    else if (k == 2) e.AddDynamicGeometry(Capsule{0.2, 0.3}, X_WA, 1, soft);
    else e.AddDynamicGeometry(Cylinder{0.25, 0.4}, X_WA, 1, soft);
    if (k < 4) e.AddDynamicGeometry(Sphere{R}, X_WB, 2, rigid);
    else e.AddDynamicGeometry(Cylinder{0.2, 0.3}, X_WB, 2, rigid);
    auto poly = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(poly.size() == 1);
    const int faces = std::atoi(argv[2 + 3 * k]);
    const double area = std::atof(argv[3 + 3 * k]), psum = std::atof(argv[4 + 3 * k]);
    double got_p = 0;
    for (int f = 0; f < poly[0].num_faces(); ++f) got_p += poly[0].pressure_at_centroid(f);
    std::printf("case %d: faces %d (want %d) area %.17g (want %.17g) pressure sum %.17g (want %.17g)\n", k,
                poly[0].num_faces(), faces, poly[0].total_area(), area, got_p, psum);
    EXPECT(poly[0].num_faces() == faces);
    EXPECT(std::fabs(poly[0].total_area() - area) <= 1e-10 * area);
    EXPECT(std::fabs(got_p - psum) <= 1e-10 * std::fabs(psum));
  }
  {  // A rigid sphere against a Cylinder and a Capsule without hydroelastic representation: Drake's closed forms
     // (penetration_as_point_pair_callback.cc:92-118, distance_to_point_callback.cc:79-160,194-288) fill point_pairs.
    ProximityProperties none;   // HydroelasticType::kUndefined
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::Identity(), 1, rigid);
    e.AddDynamicGeometry(Cylinder{0.1, 0.4}, RigidTransformd::FromTranslation(0.3, 0, 0), 2, none);
    e.AddDynamicGeometry(Capsule{0.1, 0.4}, RigidTransformd::FromTranslation(0, 0, 0.5), 3, none);
    e.ExcludeCollisionsBetween(2, 3);   // cylinder x capsule would need FCL
    std::vector<ContactSurface> surfaces;
    std::vector<PenetrationAsPointPair> pairs;
    e.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(surfaces.empty() && pairs.size() == 2);
    EXPECT(pairs[0].id_A == 1 && pairs[0].id_B == 2 && pairs[1].id_A == 1 && pairs[1].id_B == 3);
    EXPECT(std::fabs(pairs[0].depth - 0.05) < 1e-15 && std::fabs(pairs[0].nhat_BA_W[0] + 1.0) < 1e-15);
    EXPECT(std::fabs(pairs[0].p_WCa[0] - R) < 1e-15 && std::fabs(pairs[0].p_WCb[0] - 0.2) < 1e-15);
    EXPECT(std::fabs(pairs[1].depth - 0.05) < 1e-15 && std::fabs(pairs[1].nhat_BA_W[2] + 1.0) < 1e-15);
    EXPECT(std::fabs(pairs[1].p_WCa[2] - R) < 1e-15 && std::fabs(pairs[1].p_WCb[2] - 0.2) < 1e-15);
  }
  std::printf("ok\n");
  return 0;
}

// A generated mesh as plain arrays (what a caller who loaded a .vtk / .obj would hold).
static void CopyMesh(dbh_mesh* m, std::vector<double>* vertices, std::vector<int32_t>* elements) {
  int32_t nv = 0, ne = 0, vpe = 0, has_p = 0;
  dbh_mesh_get_sizes(m, &nv, &ne, &vpe, &has_p);
  vertices->resize(3 * static_cast<size_t>(nv));
  elements->resize(static_cast<size_t>(vpe) * ne);
  dbh_mesh_copy(m, vertices->data(), elements->data(), nullptr);
}

// The reference's compliant Convex (MakeConvexVolumeMesh of the hull + MakeVolumeMeshPressureField) and compliant
// Mesh(.vtk) (the tetrahedra as given + MakeVolumeMeshPressureField), hydroelastic_internal.cc:555-640, against a rigid
// sphere; the expected numbers come from the oracle (tests/test_cpp_adapter.py).
static int ConvexCase(char** argv) {
  const double R = 0.25;
  ProximityProperties soft;
  soft.compliance_type = HydroelasticType::kCompliant;
  soft.hydroelastic_modulus = 1e8;
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  rigid.resolution_hint = 0.1;
  const RigidTransformd X_WA = RotX(M_PI, 0, 0, 0);
  const RigidTransformd X_WB = RigidTransformd::FromTranslation(0, 0, -(R + R) + 0.125);
  for (int k = 0; k < 3; ++k) {
    ProximityEngine e(0);
    dbh_mesh *vol = nullptr, *surf = nullptr;
    if (k == 2) {   // the reference's RIGID Mesh(.vtk): the boundary of the given tetrahedra, under a compliant sphere
      VolumeMeshShape tets;
      EXPECT(dbh_mesh_make_box_volume(0.5, 0.4, 0.5, 0.125, &vol) == DBH_OK);
      CopyMesh(vol, &tets.vertices, &tets.tetrahedra);
      soft.resolution_hint = 0.25;
      e.AddDynamicGeometry(Sphere{R}, X_WB, 1, soft);
      e.AddDynamicGeometry(tets, X_WA, 2, rigid);
    } else if (k == 0) {
      SurfaceMeshShape hull;
      EXPECT(dbh_mesh_make_box_volume_with_ma(0.5, 0.4, 0.5, &vol) == DBH_OK);
      EXPECT(dbh_mesh_convert_volume_to_surface(vol, &surf) == DBH_OK);
      CopyMesh(surf, &hull.vertices, &hull.triangles);
      e.AddDynamicGeometry(hull, X_WA, 1, soft);
    } else {
      VolumeMeshShape tets;   // no pressure given
      EXPECT(dbh_mesh_make_box_volume(0.5, 0.4, 0.5, 0.125, &vol) == DBH_OK);
      CopyMesh(vol, &tets.vertices, &tets.tetrahedra);
      e.AddDynamicGeometry(tets, X_WA, 1, soft);
    }
    if (vol) dbh_mesh_destroy(vol);
    if (surf) dbh_mesh_destroy(surf);
    if (k < 2) e.AddDynamicGeometry(Sphere{R}, X_WB, 2, rigid);
    auto poly = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(poly.size() == 1);
    const int faces = std::atoi(argv[2 + 3 * k]);
    const double area = std::atof(argv[3 + 3 * k]), psum = std::atof(argv[4 + 3 * k]);
    double got_p = 0;
    for (int f = 0; f < poly[0].num_faces(); ++f) got_p += poly[0].pressure_at_centroid(f);
    std::printf("convex case %d: faces %d (want %d) area %.17g (want %.17g) pressure sum %.17g (want %.17g)\n", k,
                poly[0].num_faces(), faces, poly[0].total_area(), area, got_p, psum);
    EXPECT(poly[0].num_faces() == faces);
    EXPECT(std::fabs(poly[0].total_area() - area) <= 1e-10 * area);
    EXPECT(std::fabs(got_p - psum) <= 1e-10 * std::fabs(psum));
  }
  bool threw = false;
  try {   // margin > 0 would need MakeInflatedMesh / MakeConvexHull
    ProximityEngine bad(0);
    soft.margin = 0.01;
    bad.AddDynamicGeometry(VolumeMeshShape{{0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 1, 2, 3}, {}}, X_WA, 1, soft);
  } catch (const std::logic_error&) {
    threw = true;
  }
  EXPECT(threw);
  std::printf("ok\n");
  return 0;
}

int main(int argc, char** argv) {
  std::unique_ptr<ProximityEngine> engine;
  try {
    engine = std::make_unique<ProximityEngine>(0);
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    return 3;
  }
  if (argc == 17 && std::string(argv[1]) == "margin") return MarginCase(argv);
  if (argc == 11 && std::string(argv[1]) == "convex") return ConvexCase(argv);
  const double R = 0.25;
  ProximityProperties soft;
  soft.compliance_type = HydroelasticType::kCompliant;
  soft.hydroelastic_modulus = 1e8;
  soft.resolution_hint = 0.25;
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  rigid.resolution_hint = 0.1;
  const RigidTransformd X_WA = RotX(M_PI, 0, 0, 0);
  const RigidTransformd X_WB = RigidTransformd::FromTranslation(0, 0, -(R + R) + 0.125);

  {  // compliant (id 1) x rigid (id 2)
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, X_WA, 1, soft);
    e.AddDynamicGeometry(Sphere{R}, X_WB, 2, rigid);
    auto tri = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kTriangle, {});
    EXPECT(tri.size() == 1 && tri[0].num_faces() == 100 && tri[0].is_triangle());
    auto poly = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(poly.size() == 1 && poly[0].num_faces() == 28);
    EXPECT(poly[0].id_M() == 1 && poly[0].id_N() == 2);
    EXPECT(poly[0].HasGradE_M() && !poly[0].HasGradE_N());
    EXPECT(std::fabs(poly[0].total_area() - tri[0].total_area()) < 1e-12 * poly[0].total_area());
    double sum = 0;
    for (int f = 0; f < poly[0].num_faces(); ++f) sum += poly[0].area(f);
    EXPECT(std::fabs(sum - poly[0].total_area()) < 1e-12);
    bool threw = false;
    try {
      poly[0].EvaluateGradE_N_W(0);
    } catch (const std::runtime_error&) {
      threw = true;
    }
    EXPECT(threw);
    // Pose updates through the map, as QueryObject::FullPoseUpdate would deliver them.
    auto none = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon,
                                         {{2, RigidTransformd::FromTranslation(0, 0, -5.0)}});
    EXPECT(none.empty());
  }
  {  // compliant x compliant
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, X_WA, 1, soft);
    e.AddDynamicGeometry(Sphere{R}, X_WB, 2, soft);
    EXPECT(e.ComputeContactSurfaces(HydroelasticContactRepresentation::kTriangle, {})[0].num_faces() == 12);
    auto poly = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(poly[0].num_faces() == 4 && poly[0].HasGradE_M() && poly[0].HasGradE_N());
  }
  {  // compliant x compliant through the surface-seeded intersector (use_surfaces = true, field_intersection.h:299-320):
     // the same 12 triangles / 4 polygons, the same total area as the volume intersector.
    ProximityEngine by_volume(0), by_surface(0, /*use_surfaces=*/true);
    for (ProximityEngine* e : {&by_volume, &by_surface}) {
      e->AddDynamicGeometry(Sphere{R}, X_WA, 1, soft);
      e->AddDynamicGeometry(Sphere{R}, X_WB, 2, soft);
      e->AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0.3, -0.2), 3, soft);
    }
    EXPECT(by_surface.ComputeContactSurfaces(HydroelasticContactRepresentation::kTriangle, {})[0].num_faces() == 12);
    auto a = by_volume.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    auto b = by_surface.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(a.size() == b.size() && a.size() >= 2 && b[0].num_faces() == 4);
    for (size_t i = 0; i < a.size(); ++i) {
      EXPECT(a[i].id_M() == b[i].id_M() && a[i].id_N() == b[i].id_N() && a[i].num_faces() == b[i].num_faces());
      EXPECT(std::fabs(a[i].total_area() - b[i].total_area()) <= 1e-13 * a[i].total_area());
    }
  }
  {  // ordering, rigid-rigid, fallback, filter, batching
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0.2, 0.2), 30, soft);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::Identity(), 10, rigid);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, 0.375), 20, soft);
    auto s = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    EXPECT(s.size() == 3);
    EXPECT(s[0].id_M() == 10 && s[0].id_N() == 20 && s[1].id_M() == 10 && s[1].id_N() == 30 && s[2].id_M() == 20 &&
           s[2].id_N() == 30);
    EXPECT(!s[0].HasGradE_M() && s[0].HasGradE_N());  // rigid id < compliant id: M and N swapped
    e.ExcludeCollisionsBetween(10, 30);
    EXPECT(e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {}).size() == 2);
    // Batched: 4 scenes, the ball pair separated in scenes 1 and 3.
    e.set_num_scenes(4);
    std::vector<double> poses;
    for (int sc = 0; sc < 4; ++sc) {
      const RigidTransformd X[3] = {RigidTransformd::FromTranslation(0, 0.2, 0.2), RigidTransformd::Identity(),
                                    RigidTransformd::FromTranslation(0, 0, (sc % 2) ? 3.0 : 0.375)};
      for (const auto& x : X) poses.insert(poses.end(), x.m.begin(), x.m.end());
    }
    auto batched = e.ComputeContactSurfacesBatched(HydroelasticContactRepresentation::kPolygon, poses);
    EXPECT(batched.size() == 4 && batched[0].size() == 2 && batched[1].size() == 0 && batched[2].size() == 2);
    EXPECT(batched[0][0].num_faces() == batched[2][0].num_faces());
  }
  {
    ProximityEngine e(0);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::Identity(), 1, rigid);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, 0.3), 2, rigid);
    bool threw = false;
    try {
      e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
    } catch (const std::logic_error& err) {
      threw = std::string(err.what()).find("rigid-rigid") != std::string::npos;
    }
    EXPECT(threw);
    std::vector<ContactSurface> surfaces(1);  // WithFallback appends, never clears
    std::vector<PenetrationAsPointPair> pairs;
    e.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(surfaces.size() == 1);
    // The rigid-rigid pair comes back as a point pair (proximity_engine.cc:958-970): two spheres of radius R whose
    // centres are 0.3 apart along z penetrate by 2 R - 0.3, normal from B (the upper one) into A.
    EXPECT(pairs.size() == 1 && pairs[0].id_A == 1 && pairs[0].id_B == 2);
    EXPECT(std::fabs(pairs[0].depth - (2 * R - 0.3)) < 1e-15);
    EXPECT(std::fabs(pairs[0].nhat_BA_W[2] + 1.0) < 1e-15 && pairs[0].nhat_BA_W[0] == 0 && pairs[0].nhat_BA_W[1] == 0);
    EXPECT(std::fabs(pairs[0].p_WCa[2] - R) < 1e-15 && std::fabs(pairs[0].p_WCb[2] - (0.3 - R)) < 1e-15);
    // Far apart: neither an exception from the strict query nor a point pair (the geometry-level broad phase).
    auto none = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon,
                                         {{2, RigidTransformd::FromTranslation(0, 0, 3.0)}});
    EXPECT(none.empty());
    pairs.clear();
    e.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(pairs.empty());
  }
  {  // A compliant sphere on a rigid half space AND a rigid sphere on the same half space: one surface, one point pair.
    ProximityEngine e(0);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(0, 0, 0.2), 2, soft);
    e.AddDynamicGeometry(Sphere{R}, RigidTransformd::FromTranslation(2.0, 0, 0.15), 3, rigid);
    std::vector<ContactSurface> surfaces;
    std::vector<PenetrationAsPointPair> pairs;
    e.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(surfaces.size() == 1 && surfaces[0].id_M() == 1 && surfaces[0].id_N() == 2);
    EXPECT(pairs.size() == 1 && pairs[0].id_A == 1 && pairs[0].id_B == 3);
    EXPECT(std::fabs(pairs[0].depth - (R - 0.15)) < 1e-15);
    // A is the half space, B the sphere: the normal points from the sphere into the half space, i.e. down.
    EXPECT(std::fabs(pairs[0].nhat_BA_W[2] + 1.0) < 1e-15);
    EXPECT(std::fabs(pairs[0].p_WCa[2]) < 1e-15 && std::fabs(pairs[0].p_WCb[2] - (0.15 - R)) < 1e-15);
    // Two rigid boxes, the upper one rotated about z and 0.01 deep in the lower one's top face: FCL's box-box.
    ProximityEngine bb(0);
    bb.AddDynamicGeometry(Box{0.4, 0.4, 0.2}, RigidTransformd::Identity(), 1, rigid);
    RigidTransformd X_top;
    X_top.m = {std::cos(0.3), std::sin(0.3), 0, -std::sin(0.3), std::cos(0.3), 0, 0, 0, 1, 0.02, -0.01, 0.1 + 0.05 - 0.01};
    bb.AddDynamicGeometry(Box{0.1, 0.1, 0.1}, X_top, 2, rigid);
    pairs.clear();
    bb.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(pairs.size() == 1 && pairs[0].id_A == 1 && pairs[0].id_B == 2);
    if (pairs.size() == 1) {
      EXPECT(std::fabs(pairs[0].depth - 0.01) < 1e-15 && std::fabs(pairs[0].nhat_BA_W[2] + 1.0) < 1e-15);
      EXPECT(std::fabs(pairs[0].p_WCa[2] - 0.1) < 1e-15 && std::fabs(pairs[0].p_WCb[2] - 0.09) < 1e-15);
    }
    // A rigid cylinder standing 0.01 deep in a rigid box's top face (the reference's TangentStandingCylinder situation
    // turned over): libccd's MPR behind fcl::collide.
    ProximityEngine bc(0);
    bc.AddDynamicGeometry(Box{0.4, 0.4, 0.2}, RigidTransformd::Identity(), 1, rigid);
    bc.AddDynamicGeometry(Cylinder{0.1, 0.2}, RigidTransformd::FromTranslation(0.02, -0.01, 0.1 + 0.1 - 0.01), 2, rigid);
    pairs.clear();
    bc.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(pairs.size() == 1 && pairs[0].id_A == 1 && pairs[0].id_B == 2);
    if (pairs.size() == 1) {
      EXPECT(std::fabs(pairs[0].depth - 0.01) < 1e-6 && std::fabs(pairs[0].nhat_BA_W[2] + 1.0) < 1e-6);
      EXPECT(std::fabs(pairs[0].p_WCa[2] - 0.1) < 1e-3 && std::fabs(pairs[0].p_WCb[2] - 0.09) < 1e-3);
    }
    // The batched fallback query: 3 scenes of {rigid HalfSpace ground, rigid box, compliant sphere}; the box sits 0.02,
    // 0.0 (apart by 0.5) and 0.05 deep in the ground, the sphere always rests 0.03 deep: one surface per scene, and a
    // point pair (FCL's box x half space) in scenes 0 and 2 with that scene's depth.
    ProximityEngine sc(0);
    sc.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
    sc.AddDynamicGeometry(Box{0.2, 0.2, 0.2}, RigidTransformd::Identity(), 2, rigid);
    sc.AddDynamicGeometry(Sphere{R}, RigidTransformd::Identity(), 3, soft);
    sc.ExcludeCollisionsBetween(2, 3);
    sc.set_num_scenes(3);
    const double box_z[3] = {0.1 - 0.02, 0.1 + 0.5, 0.1 - 0.05};
    std::vector<double> scene_poses;
    for (int sidx = 0; sidx < 3; ++sidx) {
      const RigidTransformd Xs[3] = {RigidTransformd::Identity(), RigidTransformd::FromTranslation(1.0, 0, box_z[sidx]),
                                     RigidTransformd::FromTranslation(-1.0, 0, R - 0.03)};
      for (const auto& X : Xs) scene_poses.insert(scene_poses.end(), X.m.begin(), X.m.end());
    }
    std::vector<std::vector<ContactSurface>> batch_surfaces;
    std::vector<std::vector<PenetrationAsPointPair>> batch_pairs;
    sc.ComputeContactSurfacesWithFallbackBatched(HydroelasticContactRepresentation::kPolygon, scene_poses, &batch_surfaces,
                                                 &batch_pairs);
    EXPECT(batch_surfaces.size() == 3 && batch_pairs.size() == 3);
    if (batch_pairs.size() == 3) {
      for (int sidx = 0; sidx < 3; ++sidx) {
        EXPECT(batch_surfaces[sidx].size() == 1 && batch_surfaces[sidx][0].id_M() == 1 && batch_surfaces[sidx][0].id_N() == 3);
        EXPECT(batch_pairs[sidx].size() == (sidx == 1 ? 0u : 1u));
        if (batch_pairs[sidx].size() == 1) {
          const PenetrationAsPointPair& q = batch_pairs[sidx][0];
          EXPECT(q.id_A == 1 && q.id_B == 2 && std::fabs(q.depth - (0.1 - box_z[sidx])) < 1e-15);
          EXPECT(std::fabs(q.nhat_BA_W[2] + 1) < 1e-15 && std::fabs(q.p_WCa[2]) < 1e-15);
        }
      }
    }
    {  // ProximityEngineTests.ComputePointPairPenetration (geometry/test/proximity_engine_test.cc:1569-1640): spheres of
       // radius 0.5 with centres 0.9 apart in four clusters -- dynamic x dynamic IS reported, anchored x anchored is not,
       // dynamic x anchored IS, a filtered dynamic pair is not; results ordered and stable.
      ProximityEngine pe(0);
      EXPECT(pe.ComputePointPairPenetration({}).empty());
      ProximityProperties none;
      const Sphere sphere{0.5};
      const double d = 0.9;
      pe.AddDynamicGeometry(sphere, RigidTransformd::FromTranslation(0, 0, 0), 1, none);      // A
      pe.AddDynamicGeometry(sphere, RigidTransformd::FromTranslation(d, 0, 0), 2, none);      // B
      pe.AddAnchoredGeometry(sphere, RigidTransformd::FromTranslation(0, 100, 0), 3, none);
      pe.AddAnchoredGeometry(sphere, RigidTransformd::FromTranslation(d, 100, 0), 4, none);
      pe.AddDynamicGeometry(sphere, RigidTransformd::FromTranslation(0, 200, 0), 5, none);    // C
      pe.AddAnchoredGeometry(sphere, RigidTransformd::FromTranslation(d, 200, 0), 6, none);   // D
      pe.AddDynamicGeometry(sphere, RigidTransformd::FromTranslation(0, 300, 0), 7, none);    // G
      pe.AddDynamicGeometry(sphere, RigidTransformd::FromTranslation(d, 300, 0), 8, none);    // H
      pe.ExcludeCollisionsBetween(7, 8);
      const auto results = pe.ComputePointPairPenetration({});
      EXPECT(results.size() == 2);
      if (results.size() == 2) {
        EXPECT(results[0].id_A == 1 && results[0].id_B == 2 && results[1].id_A == 5 && results[1].id_B == 6);
        EXPECT(std::fabs(results[0].depth - 0.1) < 1e-15 && std::fabs(results[1].depth - 0.1) < 1e-15);
      }
      const auto again = pe.ComputePointPairPenetration({});
      EXPECT(again.size() == 2 && again[0].id_A == 1 && again[1].id_A == 5);
      // A mixed scene: a rigid-hydroelastic box and a cylinder without hydroelastic properties on a half-space ground.
      ProximityEngine mixed(0);
      mixed.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
      mixed.AddDynamicGeometry(Box{0.2, 0.2, 0.2}, RigidTransformd::FromTranslation(0, 0, 0.09), 2, rigid);
      mixed.AddDynamicGeometry(Cylinder{0.1, 0.2}, RigidTransformd::FromTranslation(0.15, 0, 0.08), 3, none);
      const auto mp = mixed.ComputePointPairPenetration({});
      EXPECT(mp.size() == 3);   // ground x box (0.01), ground x cylinder (0.02), box x cylinder (0.05 of lateral overlap)
      if (mp.size() == 3) {
        EXPECT(mp[0].id_A == 1 && mp[0].id_B == 2 && std::fabs(mp[0].depth - 0.01) < 1e-15);
        EXPECT(mp[1].id_A == 1 && mp[1].id_B == 3 && std::fabs(mp[1].depth - 0.02) < 1e-15);
        EXPECT(mp[2].id_A == 2 && mp[2].id_B == 3 && std::fabs(mp[2].depth - 0.05) < 1e-4 && std::fabs(mp[2].nhat_BA_W[0] + 1) < 1e-3);
      }
    }
    // Two half spaces: the reference rejects the pair itself -- loud unless the caller opts out.
    ProximityEngine b(0);
    b.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
    b.AddDynamicGeometry(HalfSpace{}, RigidTransformd::FromTranslation(0, 0, 0.1), 2, rigid);
    bool threw = false;
    try {
      b.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    } catch (const std::logic_error& err) {
      threw = std::string(err.what()).find("HalfSpace") != std::string::npos;
    }
    EXPECT(threw);
    b.set_skip_unsupported_point_pairs(true);
    pairs.clear();
    b.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(pairs.empty());
  }
  {  // Rigid shapes on a rigid HalfSpace ground: rigid-rigid, so the fallback hands them to FCL's closed-form half-space
     // primitives (drake_b200_point_pair.h). The reference's BoxPenetrationTest TangentPlane2 scene
     // (penetration_as_point_pair_callback_test.cc:806-820,906-912): a unit box standing on its corner (-.5,-.5,-.5),
     // 1e-3 under the plane; expected p_WCa = 0, p_WCb = (0, 0, -1e-3), nhat_BA_W = (0, 0, -1), depth 1e-3 to 1e-12.
    const double kDepth = 1e-3;
    const double angle = std::atan(std::sqrt(2.0)), ax = std::sqrt(0.5), ay = -std::sqrt(0.5);
    const double c = std::cos(angle), sn = std::sin(angle), t = 1 - c;
    // Rodrigues for the axis (ax, ay, 0), column-major.
    RigidTransformd X_WB;
    X_WB.m = {t * ax * ax + c, t * ax * ay, -sn * ay, t * ax * ay, t * ay * ay + c, sn * ax, sn * ay, -sn * ax, c, 0, 0, 0};
    for (int r = 0; r < 3; ++r)   // p_WB = p_WC - R_WB p_BoC_B
      X_WB.m[9 + r] = (r == 2 ? -kDepth : 0.0) + 0.5 * (X_WB.m[r] + X_WB.m[3 + r] + X_WB.m[6 + r]);
    ProximityEngine e(0);
    e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, rigid);
    e.AddDynamicGeometry(Box{1, 1, 1}, X_WB, 2, rigid);
    // A capsule standing 0.02 deep, a cylinder lying on its side 0.03 deep, an ellipsoid 0.01 deep, none with a
    // hydroelastic representation (kUndefined): kUnsupported pairs, same fallback.
    ProximityProperties none;
    e.AddDynamicGeometry(Capsule{0.1, 0.4}, RigidTransformd::FromTranslation(3, 0, 0.3 - 0.02), 3, none);
    RigidTransformd X_prone;   // +90 degrees about x: the cylinder's axis along -y... (z -> y)
    X_prone.m = {1, 0, 0, 0, 0, 1, 0, -1, 0, 6, 0, 0.2 - 0.03};
    e.AddDynamicGeometry(Cylinder{0.2, 0.5}, X_prone, 4, none);
    e.AddDynamicGeometry(Ellipsoid{0.1, 0.2, 0.3}, RigidTransformd::FromTranslation(9, 0, 0.3 - 0.01), 5, none);
    std::vector<ContactSurface> surfaces;
    std::vector<PenetrationAsPointPair> pairs;
    e.ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation::kPolygon, {}, &surfaces, &pairs);
    EXPECT(surfaces.empty());
    EXPECT(pairs.size() == 4);
    if (pairs.size() == 4) {
      const PenetrationAsPointPair& b = pairs[0];
      EXPECT(b.id_A == 1 && b.id_B == 2 && std::fabs(b.depth - kDepth) < 1e-12);
      EXPECT(std::fabs(b.nhat_BA_W[0]) < 1e-12 && std::fabs(b.nhat_BA_W[1]) < 1e-12 && std::fabs(b.nhat_BA_W[2] + 1) < 1e-12);
      for (int r = 0; r < 3; ++r) {
        EXPECT(std::fabs(b.p_WCa[r]) < 1e-12);
        EXPECT(std::fabs(b.p_WCb[r] - (r == 2 ? -kDepth : 0.0)) < 1e-12);
      }
      const double depths[4] = {kDepth, 0.02, 0.03, 0.01};
      const double xs[4] = {0, 3, 6, 9};
      for (int k = 1; k < 4; ++k) {
        const PenetrationAsPointPair& q = pairs[k];
        EXPECT(q.id_A == 1 && q.id_B == 3 + (k - 1));
        EXPECT(std::fabs(q.depth - depths[k]) < 1e-15 && std::fabs(q.nhat_BA_W[2] + 1) < 1e-15);
        EXPECT(std::fabs(q.p_WCa[0] - xs[k]) < 1e-15 && std::fabs(q.p_WCa[2]) < 1e-15);
        EXPECT(std::fabs(q.p_WCb[0] - xs[k]) < 1e-15 && std::fabs(q.p_WCb[2] + depths[k]) < 1e-15);
      }
    }
  }
  std::printf("ok\n");
  return 0;
}
