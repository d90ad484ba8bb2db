// ComputeHydroelasticContactForces through the C++ mirror, against the reference's own fixture
// (multibody/plant/test/hydroelastic_traction_calculator_test.cc:536-650): a rigid unit box pressed 0.5 into a
// compliant half space of pressure scale 1 (e = 0.5 on its bottom face, area 1). Compiled by the CPU suite
// (tests/test_cpp_adapter.py), executed on the GPU box by tests/test_zz_gpu_wrench.py.
// Exit code 0 = all checks passed; 3 = no CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "drake_b200_proximity_engine.h"

using namespace drake_b200;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d: EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

int main() {
  try {
    ProximityEngine probe(0);
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    return 3;
  }
  ProximityProperties pad;  // pressure scale = modulus / slab thickness = 1
  pad.compliance_type = HydroelasticType::kCompliant;
  pad.hydroelastic_modulus = 1.0;
  pad.slab_thickness = 1.0;
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  ProximityEngine e(0);
  e.AddAnchoredGeometry(HalfSpace{}, RigidTransformd::Identity(), 1, pad);
  e.AddDynamicGeometry(Box{1.0, 1.0, 1.0}, RigidTransformd::Identity(), 2, rigid);
  const auto surfaces = e.ComputeContactSurfaces(HydroelasticContactRepresentation::kPolygon, {});
  EXPECT(surfaces.size() == 1 && surfaces[0].id_M() == 1 && surfaces[0].id_N() == 2);
  std::unordered_map<GeometryId, ContactMaterial> materials;
  materials[1] = ContactMaterial{1.0, 0.0, 0.0};
  materials[2] = ContactMaterial{std::numeric_limits<double>::infinity(), 0.0, 0.0};
  {  // VanillaTractionOverPatch: f_Bo = (0, 0, 0.5), equal and opposite on A.
    const auto F = e.ComputeHydroelasticContactForces(materials, {});
    EXPECT(F.size() == 1 && F[0].id_M == 1 && F[0].id_N == 2);
    EXPECT(std::fabs(F[0].F_Bo_W.translational[2] - 0.5) < 1e-12);
    // No check of the horizontal components: the box's side faces have vertices exactly on the boundary, where the
    // reference's `s[i0] >= 0` vertex choice (mesh_half_space_intersection.cc:188-189) builds lopsided polygons;
    // that behaviour is reproduced, not corrected.
    for (int k = 0; k < 3; ++k) EXPECT(std::fabs(F[0].F_Ao_W.translational[k] + F[0].F_Bo_W.translational[k]) < 1e-12);
  }
  {  // TractionOverPatchWithDissipation: the box moves down at 1 m/s, dissipation 1 s/m on the compliant side:
     // the bottom face's normal traction doubles (e + e d v); the side faces see no normal velocity.
    materials[1].hunt_crossley_dissipation = 1.0;
    std::unordered_map<GeometryId, BodyKinematics> bodies;
    bodies[2].V_WB.translational = {0.0, 0.0, -1.0};
    const auto F = e.ComputeHydroelasticContactForces(materials, bodies);
    EXPECT(std::fabs(F[0].F_Bo_W.translational[2] - 1.0) < 1e-12);
  }
  std::printf("ok\n");
  return 0;
}
