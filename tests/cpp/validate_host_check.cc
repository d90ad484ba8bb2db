// The property checks of the C++ mirror's geometry registration (ProximityEngine::ValidateProperties, static: no device)
// against the reference's HydroelasticRigidGeometryErrorTests / HydroelasticCompliantGeometryErrorTests
// (geometry/proximity/test/hydroelastic_internal_test.cc: TestPropertyErrors :905-940, the suites :974-996,1406-1466): every bad property throws std::logic_error
// with the reference's message "Cannot create <compliance> <Shape>; the ('hydroelastic', '<property>') property must be
// positive | non-negative" (hydroelastic_internal.cc:280-311). Test infrastructure.
#include <cstdio>
#include <regex>
#include <string>

#include "drake_b200_proximity_engine.h"

using namespace drake_b200;

static int failures = 0;

static void ExpectThrows(const Shape& shape, const ProximityProperties& p, const std::string& pattern, const char* what) {
  try {
    ProximityEngine::ValidateProperties(shape, p);
  } catch (const std::logic_error& e) {
    if (!std::regex_search(std::string(e.what()), std::regex(pattern))) {
      std::printf("FAIL %s: message '%s' does not match '%s'\n", what, e.what(), pattern.c_str());
      ++failures;
    }
    return;
  }
  std::printf("FAIL %s: did not throw\n", what);
  ++failures;
}

static void ExpectOk(const Shape& shape, const ProximityProperties& p, const char* what) {
  try {
    ProximityEngine::ValidateProperties(shape, p);
  } catch (const std::exception& e) {
    std::printf("FAIL %s: threw '%s'\n", what, e.what());
    ++failures;
  }
}

int main() {
  ProximityProperties rigid;
  rigid.compliance_type = HydroelasticType::kRigid;
  ProximityProperties soft;
  soft.compliance_type = HydroelasticType::kCompliant;
  soft.hydroelastic_modulus = 1e8;
  soft.resolution_hint = 0.1;
  const Shape curved[] = {Sphere{0.5}, Capsule{0.5, 1.0}, Cylinder{0.5, 1.0}, Ellipsoid{0.5, 0.8, 0.3}};
  const char* names[] = {"Sphere", "Capsule", "Cylinder", "Ellipsoid"};

  for (int i = 0; i < 4; ++i) {
    // BadResolutionHint, rigid (:978-988) and compliant (:1410-1422): negative and zero.
    for (double hint : {-0.2, 0.0}) {
      ProximityProperties r = rigid, s = soft;
      r.resolution_hint = s.resolution_hint = hint;
      ExpectThrows(curved[i], r, std::string("Cannot create rigid ") + names[i] + ".+'resolution_hint'.+ positive", "rigid hint");
      ExpectThrows(curved[i], s, std::string("Cannot create compliant ") + names[i] + ".+'resolution_hint'.+ positive", "soft hint");
    }
    ExpectOk(curved[i], rigid, "rigid defaults");
    ExpectOk(curved[i], soft, "soft defaults");
  }
  // A Box needs no resolution hint, rigid or compliant (hydroelastic_internal.cc:338-343,440-463).
  {
    ProximityProperties r = rigid, s = soft;
    r.resolution_hint = s.resolution_hint = -1;
    ExpectOk(Box{0.5, 0.5, 0.5}, r, "rigid box without hint");
    ExpectOk(Box{0.5, 0.5, 0.5}, s, "soft box without hint");
    ExpectOk(HalfSpace{}, r, "rigid half space");
  }
  // BadElasticModulus (:1424-1439): every compliant shape, negative and zero.
  const Shape all[] = {Sphere{0.5}, Capsule{0.5, 1.0}, Cylinder{0.5, 1.0}, Ellipsoid{0.5, 0.8, 0.3}, Box{0.5, 0.5, 0.5},
                       HalfSpace{}};
  const char* all_names[] = {"Sphere", "Capsule", "Cylinder", "Ellipsoid", "Box", "HalfSpace"};
  for (int i = 0; i < 6; ++i) {
    for (double modulus : {-0.2, 0.0}) {
      ProximityProperties s = soft;
      s.slab_thickness = 1.0;
      s.hydroelastic_modulus = modulus;
      ExpectThrows(all[i], s, std::string("Cannot create compliant ") + all_names[i] + ".+'hydroelastic_modulus'.+ positive",
                   "modulus");
    }
    ProximityProperties s = soft;
    s.slab_thickness = 1.0;
    s.margin = -1e-3;  // NonNegativeDouble (:297-311)
    ExpectThrows(all[i], s, std::string("Cannot create compliant ") + all_names[i] + ".+'margin'.+ non-negative", "margin");
    s.margin = 0.0;
    ExpectOk(all[i], s, "margin 0");
    // The rigid representations never read the margin or the modulus.
    ProximityProperties r = rigid;
    r.margin = -1;
    r.hydroelastic_modulus = -1;
    ExpectOk(all[i], r, "rigid ignores margin / modulus");
  }
  // BadSlabThickness (:1441-1458): a compliant half space only.
  for (double thickness : {-0.2, 0.0}) {
    ProximityProperties s = soft;
    s.slab_thickness = thickness;
    ExpectThrows(HalfSpace{}, s, "Cannot create compliant HalfSpace.+'slab_thickness'.+ positive", "slab thickness");
    ExpectOk(Sphere{0.5}, s, "slab thickness ignored by a sphere");
  }
  // Shape constructors (shape_specification.cc): positive, finite dimensions.
  ExpectThrows(Box{0.5, -0.5, 0.5}, soft, "Box", "box size");
  ExpectThrows(Ellipsoid{0.5, 0.0, 0.3}, rigid, "Ellipsoid", "ellipsoid axes");
  ExpectThrows(Cylinder{-1.0, 1.0}, soft, "Cylinder", "cylinder size");
  ExpectThrows(Capsule{1.0, 0.0}, soft, "Capsule", "capsule size");
  ExpectThrows(Sphere{-1.0}, soft, "Sphere", "sphere radius");
  // kUndefined: nothing is made, nothing is checked (hydroelastic_internal.cc:106-136).
  ProximityProperties none;
  none.resolution_hint = -1;
  ExpectOk(Sphere{0.5}, none, "undefined");
  std::printf(failures ? "%d failures\n" : "ok\n", failures);
  return failures ? 1 : 0;
}
