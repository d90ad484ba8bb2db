// drake_b200_proximity_engine.h — host-side C++ mirror of the reference interface for the
// hydroelastic contact-surface path, over the C-ABI in drake_b200_hydro.h.
//
// Header-only C++17, no Eigen. Mirrors, name for name and with the same argument meaning and
// error behaviour, the part of geometry::internal::ProximityEngine<double> that sits on the path
// (reference: geometry/proximity_engine.h:60-395):
//   AddDynamicGeometry / AddAnchoredGeometry      proximity_engine.h:122-133 (shapes incl. HalfSpace)
//   RemoveGeometry                                proximity_engine.h:186
//   num_geometries / num_dynamic / num_anchored   proximity_engine.h:196-202
//   UpdateWorldPoses                              proximity_engine.h:226-227
//   ComputeContactSurfaces                        proximity_engine.h:309-315
//   ComputeContactSurfacesWithFallback            proximity_engine.h:320-326
//   collision_filter().  (as ExcludeCollisionsBetween)   proximity_engine.h:87
// and geometry::ContactSurface<double> (query_results/contact_surface.h:160-415).
// Results are returned by value; WithFallback appends to caller vectors and never clears them
// (query_object.h:439-447); failures are std::logic_error / std::runtime_error as in
// proximity_engine.cc:1347-1397.
//
// One engine = one scene template. `set_num_scenes(S)` instances the registered scene S times so a
// single ComputeContactSurfacesBatched() call evaluates S pose sets on the GPU (the batched,
// scene-sharded use the north star asks for); the unbatched calls are the S = 1 case.
#ifndef DRAKE_B200_PROXIMITY_ENGINE_H_
#define DRAKE_B200_PROXIMITY_ENGINE_H_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <map>
#include <optional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <variant>
#include <vector>

#include "drake_b200_hydro.h"
#include "drake_b200_point_pair.h"

namespace drake_b200 {

using GeometryId = int64_t;

// math::RigidTransform<double>: rotation column-major then translation (math/rigid_transform.h:1004-1007).
struct RigidTransformd {
  std::array<double, 12> m{1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  static RigidTransformd Identity() { return {}; }
  static RigidTransformd FromTranslation(double x, double y, double z) {
    RigidTransformd X;
    X.m[9] = x;
    X.m[10] = y;
    X.m[11] = z;
    return X;
  }
};

// geometry/shape_specification.h: every primitive the reference gives a hydroelastic representation, plus explicit
// meshes (its Mesh / Convex arrive here as the vertex and element arrays the caller loaded).
struct Sphere {
  double radius;
};
struct Box {
  double width, depth, height;
};
struct Ellipsoid {
  double a, b, c;
};
struct HalfSpace {};  // z <= 0 of the geometry frame (shape_specification.h)
// Cylinder and Capsule (axis z, shape_specification.h): meshed by MakeCylinderVolumeMeshWithMa / MakeCapsuleVolumeMesh
// like the reference (hydroelastic_internal.cc:344-363,465-508); registered with HydroelasticType::kUndefined they take
// part in ComputeContactSurfacesWithFallback's point pairs against a Sphere (Drake's closed forms) or a HalfSpace
// (FCL's), drake_b200_point_pair.h.
struct Cylinder {
  double radius, length;
};
struct Capsule {
  double radius, length;
};
struct VolumeMeshShape {  // VolumeMesh<double> + per-vertex pressure
  std::vector<double> vertices;  // [nv*3]
  std::vector<int32_t> tetrahedra;  // [nt*4]
  std::vector<double> pressure;  // [nv]; empty: MakeVolumeMeshPressureField computes it (the reference's Mesh(.vtk))
};
// TriangleSurfaceMesh<double>. Rigid: the reference's rigid Mesh / Convex. Compliant: taken to be the convex hull of the
// reference's compliant Convex (closed, wound outwards): MakeConvexVolumeMesh + MakeVolumeMeshPressureField.
struct SurfaceMeshShape {
  std::vector<double> vertices;
  std::vector<int32_t> triangles;
};
using Shape = std::variant<Sphere, Box, Ellipsoid, VolumeMeshShape, SurfaceMeshShape, HalfSpace, Cylinder, Capsule>;

enum class HydroelasticType { kUndefined = DBH_UNDEFINED, kRigid = DBH_RIGID, kCompliant = DBH_COMPLIANT };
enum class HydroelasticContactRepresentation { kTriangle = DBH_TRIANGLE, kPolygon = DBH_POLYGON };
enum class TessellationStrategy { kSingleInteriorVertex = 0, kDenseInteriorVertices = 1 };

// The "hydroelastic" property group (geometry/proximity_properties.cc:16-21).
struct ProximityProperties {
  HydroelasticType compliance_type{HydroelasticType::kUndefined};
  double hydroelastic_modulus{1e7};
  double resolution_hint{0.02};
  double margin{0.0};
  double slab_thickness{0.0};  // kSlabThickness: compliant half spaces only (hydroelastic_internal.cc:539-553)
  TessellationStrategy tessellation_strategy{TessellationStrategy::kSingleInteriorVertex};
};

// geometry::PenetrationAsPointPair<double>, the fallback result (query_results/penetration_as_point_pair.h):
// Drake's own closed forms for a Sphere against Sphere / Box / HalfSpace / Cylinder / Capsule, FCL's closed-form half-space
// primitives, its box-box, and libccd's MPR for every other convex pair (drake_b200_point_pair.h).
struct PenetrationAsPointPair {
  GeometryId id_A, id_B;
  std::array<double, 3> p_WCa, p_WCb, nhat_BA_W;
  double depth;
};

// multibody::SpatialForce<double> / SpatialVelocity<double>: rotational part first, then translational.
struct SpatialVector {
  std::array<double, 3> rotational{}, translational{};
};
// What MultibodyPlant::CalcHydroelasticContactForcesContinuous (multibody_plant.cc:2559-2670) keeps of one contact
// surface: the spatial force about the surface centroid and the same force shifted to the two body origins
// (A = body of geometry M, B = body of geometry N).
struct HydroelasticContactForce {
  GeometryId id_M, id_N;
  SpatialVector F_Ac_W, F_Ao_W, F_Bo_W;
};
// Per geometry: where its body is and how it moves (EvalPoseInWorld / EvalSpatialVelocityInWorld of the body the
// geometry is attached to), and the material properties the continuous model reads (hydroelastic modulus, +infinity
// for a rigid geometry; Hunt-Crossley dissipation; dynamic Coulomb friction).
struct BodyKinematics {
  std::array<double, 3> p_WBo{};
  SpatialVector V_WB;
};
struct ContactMaterial {
  double hydroelastic_modulus{std::numeric_limits<double>::infinity()};
  double hunt_crossley_dissipation{0.0};
  double dynamic_friction{0.0};
};

// geometry::ContactSurface<double>. Faces and vertices are owned by value; all quantities in World.
class ContactSurface {
 public:
  GeometryId id_M() const { return id_M_; }
  GeometryId id_N() const { return id_N_; }
  HydroelasticContactRepresentation representation() const { return rep_; }
  bool is_triangle() const { return rep_ == HydroelasticContactRepresentation::kTriangle; }
  int num_faces() const { return static_cast<int>(areas_.size()); }
  int num_vertices() const { return static_cast<int>(pressures_.size()); }
  double area(int f) const { return areas_[f]; }
  double total_area() const { return total_area_; }
  const double* face_normal(int f) const { return &normals_[3 * f]; }   // out of N into M
  const double* centroid(int f) const { return &centroids_[3 * f]; }
  const double* centroid() const { return centroid_.data(); }
  const double* vertex(int v) const { return &vertices_[3 * v]; }
  double vertex_pressure(int v) const { return pressures_[v]; }          // e_MN at vertex v
  // poly_e_MN().EvaluateCartesian(f, centroid(f)) / tri_e_MN().Evaluate(f, (1/3,1/3,1/3)).
  double pressure_at_centroid(int f) const { return e_centroid_[f]; }
  bool HasGradE_M() const { return has_grad_M_; }
  bool HasGradE_N() const { return has_grad_N_; }
  const double* EvaluateGradE_M_W(int f) const {
    if (!has_grad_M_) throw std::runtime_error("ContactSurface::EvaluateGradE_M_W() invalid; no gradient values stored");
    return &grad_eM_[3 * f];
  }
  const double* EvaluateGradE_N_W(int f) const {
    if (!has_grad_N_) throw std::runtime_error("ContactSurface::EvaluateGradE_N_W() invalid; no gradient values stored");
    return &grad_eN_[3 * f];
  }
  // PolygonSurfaceMesh::face_data_ encoding |n|v0..vn-1| (polygon_surface_mesh.h:141-185) or 3
  // vertex indices per triangle.
  const std::vector<int32_t>& face_data() const { return face_data_; }
  const std::vector<double>& vertices() const { return vertices_; }
  const std::vector<double>& pressures() const { return pressures_; }

 private:
  friend class ProximityEngine;
  GeometryId id_M_{}, id_N_{};
  HydroelasticContactRepresentation rep_{HydroelasticContactRepresentation::kPolygon};
  std::vector<double> vertices_, pressures_, areas_, normals_, centroids_, e_centroid_, grad_eM_, grad_eN_;
  std::vector<int32_t> face_data_;
  double total_area_{0};
  std::array<double, 3> centroid_{};
  bool has_grad_M_{false}, has_grad_N_{false};
};

class ProximityEngine {
 public:
  // use_surfaces: HydroelasticVolumeIntersector::IntersectCompliantVolumes' switch (field_intersection.h:299-320).
  // false (the default, and what the reference's ComputeContactSurfaceFromCompliantVolumes passes,
  // field_intersection.cc:511-517): compliant x compliant pairs through the two volume hierarchies; true: through the
  // surface-seeded breadth-first intersector (dbh_context_set_soft_soft_intersector, DBH_INTERSECTOR_SURFACE_BFS).
  explicit ProximityEngine(int device_ordinal = 0, bool use_surfaces = false) {
    Check(dbh_context_create(device_ordinal, &ctx_), "ProximityEngine");
    if (use_surfaces)
      Check(dbh_context_set_soft_soft_intersector(ctx_, DBH_INTERSECTOR_SURFACE_BFS), "ProximityEngine");
  }
  ~ProximityEngine() {
    if (batch_) dbh_batch_destroy(batch_);
    if (ctx_) dbh_context_destroy(ctx_);
  }
  ProximityEngine(const ProximityEngine&) = delete;
  ProximityEngine& operator=(const ProximityEngine&) = delete;

  void AddDynamicGeometry(const Shape& shape, const RigidTransformd& X_WG, GeometryId id,
                          const ProximityProperties& props = {}) {
    AddGeometry(shape, X_WG, id, props, /*anchored=*/false);
  }
  void AddAnchoredGeometry(const Shape& shape, const RigidTransformd& X_WG, GeometryId id,
                           const ProximityProperties& props = {}) {
    AddGeometry(shape, X_WG, id, props, /*anchored=*/true);
  }
  int num_geometries() const { return static_cast<int>(ids_.size()); }
  // proximity_engine.h:196-202.
  int num_anchored() const {
    int n = 0;
    for (uint8_t a : anchored_) n += a;
    return n;
  }
  int num_dynamic() const { return num_geometries() - num_anchored(); }

  // RemoveGeometry (proximity_engine.h:186, proximity_engine.cc:483-520): the geometry leaves its tree, the
  // hydroelastic representations and the collision filter. `is_dynamic` must match how it was added
  // (the reference looks it up in the tree that flag names and .at() throws otherwise).
  void RemoveGeometry(GeometryId id, bool is_dynamic) {
    const int g = IndexOf(id);
    if ((anchored_[g] == 0) != is_dynamic)
      throw std::out_of_range("RemoveGeometry: GeometryId " + std::to_string(id) + " is not " +
                              (is_dynamic ? "dynamic" : "anchored"));
    ids_.erase(ids_.begin() + g);
    shapes_.erase(shapes_.begin() + g);
    elements_.erase(elements_.begin() + g);
    vertices_per_element_.erase(vertices_per_element_.begin() + g);
    assets_.erase(assets_.begin() + g);
    vanished_.erase(vanished_.begin() + g);
    anchored_.erase(anchored_.begin() + g);
    poses_.erase(poses_.begin() + g);
    index_.clear();
    for (size_t k = 0; k < ids_.size(); ++k) index_[ids_[k]] = static_cast<int>(k);
    std::vector<std::pair<int, int>> kept;  // CollisionFilter::RemoveGeometry: declarations naming it go away
    for (const auto& [a, b] : excluded_) {
      if (a == g || b == g) continue;
      kept.emplace_back(a > g ? a - 1 : a, b > g ? b - 1 : b);
    }
    excluded_ = std::move(kept);
    Invalidate();  // the mesh asset stays registered in the context (assets are shared by handle)
  }

  // Geometries::is_vanished (hydroelastic_internal.h): a compliant PRIMITIVE so small that its pressure field has no
  // gradient (MeshFieldLinear::is_gradient_field_degenerate) is accepted but takes no part in hydroelastic contact:
  // every pair naming it is "calculated, no surface" -- no surface, no failure, no point pair
  // (hydroelastic_internal.cc:194-204, hydroelastic_calculator.cc:81-85). A degenerate Mesh / Convex throws instead.
  bool is_vanished(GeometryId id) const { return vanished_[IndexOf(id)] != 0; }

  // CollisionFilter declaration ExcludeBetween for one pair (collision_filter.h:158 CanCollideWith).
  void ExcludeCollisionsBetween(GeometryId a, GeometryId b) {
    excluded_.emplace_back(IndexOf(a), IndexOf(b));
    Invalidate();
  }

  // Number of scene instances evaluated per batched call (default 1).
  void set_num_scenes(int num_scenes) {
    if (num_scenes < 1) throw std::logic_error("set_num_scenes: need at least one scene");
    num_scenes_ = num_scenes;
    Invalidate();
  }
  int num_scenes() const { return num_scenes_; }

  // UpdateWorldPoses: only ids present in the map are updated (proximity_engine.cc:551-576).
  void UpdateWorldPoses(const std::unordered_map<GeometryId, RigidTransformd>& X_WGs) {
    for (const auto& [id, X] : X_WGs) {
      auto it = index_.find(id);
      if (it != index_.end()) poses_[it->second] = X;
    }
  }

  std::vector<ContactSurface> ComputeContactSurfaces(
      HydroelasticContactRepresentation representation,
      const std::unordered_map<GeometryId, RigidTransformd>& X_WGs) {
    UpdateWorldPoses(X_WGs);
    auto all = Run(representation, DBH_QUERY_STRICT, PosesOfOneScene());
    return std::move(all[0]);
  }

  void ComputeContactSurfacesWithFallback(HydroelasticContactRepresentation representation,
                                          const std::unordered_map<GeometryId, RigidTransformd>& X_WGs,
                                          std::vector<ContactSurface>* surfaces,
                                          std::vector<PenetrationAsPointPair>* point_pairs) {
    if (surfaces == nullptr || point_pairs == nullptr)  // DRAKE_DEMAND (query_object.cc:174-175)
      throw std::logic_error("ComputeContactSurfacesWithFallback: null output");
    UpdateWorldPoses(X_WGs);
    auto all = Run(representation, DBH_QUERY_WITH_FALLBACK, PosesOfOneScene());
    for (auto& s : all[0]) surfaces->push_back(std::move(s));
    // The pairs whose surface could not be made and whose world boxes overlap go to the point-pair narrow phase
    // (proximity_engine.cc:958-970), in (id_A, id_B) order.
    int64_t n = 0;
    Check(dbh_batch_copy_fallback_pairs(batch_, nullptr, &n), "copy_fallback_pairs");
    std::vector<int32_t> quads(static_cast<size_t>(4 * n));
    if (n) Check(dbh_batch_copy_fallback_pairs(batch_, quads.data(), &n), "copy_fallback_pairs");
    for (int64_t k = 0; k < n; ++k) {
      auto pp = MaybeMakePointPair(quads[4 * k + 1], quads[4 * k + 2]);
      if (pp) point_pairs->push_back(*pp);
    }
  }

  // What to do with the one pair the reference itself rejects, HalfSpace x HalfSpace (it throws): false (the default)
  // throws like the reference, true skips the pair.
  void set_skip_unsupported_point_pairs(bool skip) { skip_unsupported_point_pairs_ = skip; }

  // Batched form: X_WGs is [num_scenes * num_geometries * 12] doubles, scene-major, geometries in
  // registration order. Returns the surfaces of every scene, each list sorted by (id_M, id_N).
  std::vector<std::vector<ContactSurface>> ComputeContactSurfacesBatched(
      HydroelasticContactRepresentation representation, const std::vector<double>& X_WGs) {
    if (X_WGs.size() != static_cast<size_t>(num_scenes_) * ids_.size() * 12)
      throw std::logic_error("ComputeContactSurfacesBatched: expected num_scenes * num_geometries poses");
    return Run(representation, DBH_QUERY_STRICT, X_WGs);
  }

  // ProximityEngine::ComputePointPairPenetration (proximity_engine.cc:740-760: FclCollide over dynamic x dynamic and
  // dynamic x anchored with penetration_as_point_pair::Callback, then std::sort): one PenetrationAsPointPair per
  // unfiltered pair of penetrating geometries, whatever their hydroelastic types, ordered by (id_A, id_B). Host-side
  // throughout -- the same narrow phase ComputeContactSurfacesWithFallback uses for its no-surface pairs
  // (drake_b200_point_pair.h) behind a geometry-level broad phase over FCL's world boxes (computeAABB: the local box
  // translated under an identity rotation, else the cube around its bounding sphere; a half space bounded only under the
  // identity). It is the point-contact model's query, next to the path rather than on it: nothing here runs on the GPU.
  std::vector<PenetrationAsPointPair> ComputePointPairPenetration(const std::unordered_map<GeometryId, RigidTransformd>& X_WGs) {
    UpdateWorldPoses(X_WGs);
    const int G = static_cast<int>(ids_.size());
    std::vector<int> order(G);
    for (int g = 0; g < G; ++g) order[g] = g;
    std::sort(order.begin(), order.end(), [this](int x, int y) { return ids_[x] < ids_[y]; });
    struct WorldBox {
      double mn[3], mx[3];
    };
    std::vector<WorldBox> boxes(G);
    for (int g = 0; g < G; ++g) {
      WorldBox& wb = boxes[g];
      for (int k = 0; k < 3; ++k) {
        wb.mn[k] = -std::numeric_limits<double>::infinity();
        wb.mx[k] = std::numeric_limits<double>::infinity();
      }
      const auto& X = poses_[g].m;
      bool identity = true;
      for (int k = 0; k < 9; ++k) identity = identity && std::abs(X[k] - (k % 4 == 0 ? 1.0 : 0.0)) <= 1e-12;
      std::array<double, 6> local;
      if (std::holds_alternative<HalfSpace>(shapes_[g])) {
        if (identity) wb.mx[2] = X[11];
      } else if (LocalBox(shapes_[g], &local)) {
        if (identity) {
          for (int k = 0; k < 3; ++k) {
            wb.mn[k] = X[9 + k] + (local[k] - local[3 + k]);
            wb.mx[k] = X[9 + k] + (local[k] + local[3 + k]);
          }
        } else {
          const point_pair::Vec3 c = point_pair::Apply(X, {local[0], local[1], local[2]});
          const double r = point_pair::Norm({local[3], local[4], local[5]});
          for (int k = 0; k < 3; ++k) {
            wb.mn[k] = c[k] - r;
            wb.mx[k] = c[k] + r;
          }
        }
      }
    }
    std::vector<PenetrationAsPointPair> out;
    for (int i = 0; i < G; ++i) {
      for (int j = i + 1; j < G; ++j) {
        const int ga = order[i], gb = order[j];   // id(ga) < id(gb)
        if (anchored_[ga] && anchored_[gb]) continue;
        bool filtered = false;
        for (const auto& [a, b] : excluded_) filtered = filtered || (a == ga && b == gb) || (a == gb && b == ga);
        if (filtered) continue;
        bool overlap = true;
        for (int k = 0; k < 3; ++k) overlap = overlap && boxes[ga].mn[k] <= boxes[gb].mx[k] && boxes[gb].mn[k] <= boxes[ga].mx[k];
        if (!overlap) continue;
        auto pp = MaybeMakePointPair(ga, gb);
        if (pp) out.push_back(*pp);
      }
    }
    return out;
  }

  // ComputeContactSurfacesWithFallback for every scene of the batch: the surfaces per scene and, per scene, the point
  // pairs of the pairs whose surface cannot be made (proximity_engine.cc:936-977), each list in (id_A, id_B) order.
  void ComputeContactSurfacesWithFallbackBatched(HydroelasticContactRepresentation representation,
                                                 const std::vector<double>& X_WGs,
                                                 std::vector<std::vector<ContactSurface>>* surfaces,
                                                 std::vector<std::vector<PenetrationAsPointPair>>* point_pairs) {
    if (surfaces == nullptr || point_pairs == nullptr) throw std::logic_error("ComputeContactSurfacesWithFallbackBatched: null output");
    const size_t G = ids_.size();
    if (X_WGs.size() != static_cast<size_t>(num_scenes_) * G * 12)
      throw std::logic_error("ComputeContactSurfacesWithFallbackBatched: expected num_scenes * num_geometries poses");
    *surfaces = Run(representation, DBH_QUERY_WITH_FALLBACK, X_WGs);
    point_pairs->assign(static_cast<size_t>(num_scenes_), {});
    if (G == 0) return;
    int64_t n = 0;
    Check(dbh_batch_copy_fallback_pairs(batch_, nullptr, &n), "copy_fallback_pairs");
    std::vector<int32_t> quads(static_cast<size_t>(4 * n));
    if (n) Check(dbh_batch_copy_fallback_pairs(batch_, quads.data(), &n), "copy_fallback_pairs");
    struct Restore {   // MaybeMakePointPair may throw
      const double*& p;
      ~Restore() { p = nullptr; }
    } restore{scene_poses_};
    for (int64_t k = 0; k < n; ++k) {
      const int scene = quads[4 * k];
      scene_poses_ = X_WGs.data() + static_cast<size_t>(scene) * G * 12;
      auto pp = MaybeMakePointPair(quads[4 * k + 1] - scene * static_cast<int>(G), quads[4 * k + 2] - scene * static_cast<int>(G));
      if (pp) (*point_pairs)[scene].push_back(*pp);
    }
  }

  dbh_stats stats() const {
    dbh_stats s{};
    if (batch_) dbh_batch_stats(batch_, &s);
    return s;
  }

  // The consumer of the surfaces in the continuous hydroelastic model, on the device:
  // HydroelasticTractionCalculator::ComputeSpatialForcesAtCentroidFromHydroelasticModel +
  // ShiftSpatialForcesAtCentroidToBodyOrigins (multibody/plant/hydroelastic_traction_calculator.cc:24-83) for every
  // surface of the last ComputeContactSurfaces* call of a one-scene engine, in result order. vslip_regularizer is the
  // calculator's constructor argument (the plant passes its stiction tolerance). Geometries without an entry in
  // `bodies` sit at rest with their body origin at the geometry frame's origin as last posed; every geometry needs a
  // ContactMaterial.
  std::vector<HydroelasticContactForce> ComputeHydroelasticContactForces(
      const std::unordered_map<GeometryId, ContactMaterial>& materials,
      const std::unordered_map<GeometryId, BodyKinematics>& bodies, double vslip_regularizer = 1e-6) {
    if (!batch_ || num_scenes_ != 1)
      throw std::logic_error("ComputeHydroelasticContactForces: run ComputeContactSurfaces on a one-scene engine first");
    const size_t G = ids_.size();
    std::vector<double> E(G), d(G), mu(G), origin(3 * G), vel(6 * G, 0.0);
    for (size_t g = 0; g < G; ++g) {
      const auto m = materials.find(ids_[g]);
      if (m == materials.end())
        throw std::logic_error("ComputeHydroelasticContactForces: no ContactMaterial for GeometryId " + std::to_string(ids_[g]));
      E[g] = m->second.hydroelastic_modulus;
      d[g] = m->second.hunt_crossley_dissipation;
      mu[g] = m->second.dynamic_friction;
      const auto b = bodies.find(ids_[g]);
      for (int k = 0; k < 3; ++k) {
        origin[3 * g + k] = b != bodies.end() ? b->second.p_WBo[k] : poses_[g].m[9 + k];
        if (b != bodies.end()) {
          vel[6 * g + k] = b->second.V_WB.rotational[k];
          vel[6 * g + 3 + k] = b->second.V_WB.translational[k];
        }
      }
    }
    Check(dbh_batch_set_contact_properties(batch_, E.data(), d.data(), mu.data()), "set_contact_properties");
    Check(dbh_batch_set_body_kinematics(batch_, origin.data(), vel.data()), "set_body_kinematics");
    Check(dbh_batch_compute_spatial_forces(batch_, vslip_regularizer), "compute_spatial_forces");
    Check(dbh_batch_sync(batch_), "ComputeHydroelasticContactForces");
    dbh_result_sizes sz{};
    Check(dbh_batch_result_sizes(batch_, &sz), "result_sizes");
    const size_t ns = sz.num_surfaces;
    std::vector<double> F_Ac(6 * ns), F_Ao(6 * ns), F_Bo(6 * ns);
    std::vector<int64_t> id_M(ns), id_N(ns);
    Check(dbh_batch_copy_spatial_forces(batch_, F_Ac.data(), F_Ao.data(), F_Bo.data()), "copy_spatial_forces");
    dbh_result_buffers buf{};
    buf.surf_id_M = id_M.data();
    buf.surf_id_N = id_N.data();
    Check(dbh_batch_copy_results(batch_, &buf), "copy_results");
    std::vector<HydroelasticContactForce> out(ns);
    auto fill = [](const std::vector<double>& src, size_t s, SpatialVector* v) {
      for (int k = 0; k < 3; ++k) {
        v->rotational[k] = src[6 * s + k];
        v->translational[k] = src[6 * s + 3 + k];
      }
    };
    for (size_t s = 0; s < ns; ++s) {
      out[s].id_M = id_M[s];
      out[s].id_N = id_N[s];
      fill(F_Ac, s, &out[s].F_Ac_W);
      fill(F_Ao, s, &out[s].F_Ao_W);
      fill(F_Bo, s, &out[s].F_Bo_W);
    }
    return out;
  }

 private:
  static void Check(dbh_status s, const char* where) {
    if (s == DBH_OK) return;
    const std::string msg = std::string(where) + ": " + dbh_last_error();
    // The reference throws std::logic_error for unsupported / rigid-rigid pairs
    // (proximity_engine.cc:1361-1390); everything else is a runtime failure.
    if (s == DBH_ERR_RIGID_RIGID || s == DBH_ERR_UNSUPPORTED || s == DBH_ERR_INVALID_ARGUMENT ||
        s == DBH_ERR_HALF_SPACE_HALF_SPACE || s == DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH)
      throw std::logic_error(msg);
    throw std::runtime_error(msg);
  }
  static void CheckMesh(dbh_status s, const char* where) {
    if (s != DBH_OK) throw std::logic_error(std::string(where) + ": " + dbh_mesh_last_error());
  }
  int IndexOf(GeometryId id) const {
    auto it = index_.find(id);
    if (it == index_.end()) throw std::logic_error("unknown GeometryId " + std::to_string(id));
    return it->second;
  }
  void Invalidate() {
    if (batch_) dbh_batch_destroy(batch_);
    batch_ = nullptr;
  }

  // hydroelastic::Geometries::MaybeAddGeometry (hydroelastic_internal.cc:145-153, 328-343, 412-536).
  // penetration_as_point_pair::ComputeNarrowPhasePenetration (penetration_as_point_pair_callback.cc:203-300) for the
  // geometries at indices ga, gb (id(ga) < id(gb)) at their current poses.
  // The pose the point-pair arithmetic sees for geometry g: the engine's own (one scene) or, inside the batched
  // fallback query, that scene's slice of the caller's pose array.
  point_pair::Pose PoseOf(int g) const {
    if (scene_poses_ == nullptr) return poses_[g].m;
    point_pair::Pose X;
    std::copy(scene_poses_ + 12 * g, scene_poses_ + 12 * g + 12, X.begin());
    return X;
  }

  std::optional<PenetrationAsPointPair> MaybeMakePointPair(int ga, int gb) const {
    namespace pp = point_pair;
    const Shape& a = shapes_[ga];
    const Shape& b = shapes_[gb];
    const bool a_is_sphere = std::holds_alternative<Sphere>(a), b_is_sphere = std::holds_alternative<Sphere>(b);
    if (!a_is_sphere && !b_is_sphere) {
      const bool a_is_half_space = std::holds_alternative<HalfSpace>(a), b_is_half_space = std::holds_alternative<HalfSpace>(b);
      if (a_is_half_space != b_is_half_space) {
        // FCL's closed-form half-space primitives behind CalcDistanceFallback (:145-207), drake_b200_point_pair.h.
        const int gh = a_is_half_space ? ga : gb, gx = a_is_half_space ? gb : ga;   // half space, the other shape
        const pp::Pose X_WH = PoseOf(gh);
        const pp::Pose X_WX = PoseOf(gx);
        std::optional<pp::FclHalfSpaceContact> contact;
        if (const auto* bx = std::get_if<Box>(&shapes_[gx])) {
          contact = pp::BoxHalfSpace({bx->width, bx->depth, bx->height}, X_WX, X_WH);
        } else if (const auto* cy = std::get_if<Cylinder>(&shapes_[gx])) {
          contact = pp::CylinderHalfSpace(cy->radius, cy->length, X_WX, X_WH);
        } else if (const auto* ca = std::get_if<Capsule>(&shapes_[gx])) {
          contact = pp::CapsuleHalfSpace(ca->radius, ca->length, X_WX, X_WH);
        } else if (const auto* el = std::get_if<Ellipsoid>(&shapes_[gx])) {
          contact = pp::EllipsoidHalfSpace({el->a, el->b, el->c}, X_WX, X_WH);
        } else if (const auto* sm = std::get_if<SurfaceMeshShape>(&shapes_[gx])) {
          contact = pp::VerticesHalfSpace(sm->vertices.data(), sm->vertices.size() / 3, X_WX, X_WH);
        } else if (const auto* vm = std::get_if<VolumeMeshShape>(&shapes_[gx])) {
          contact = pp::VerticesHalfSpace(vm->vertices.data(), vm->vertices.size() / 3, X_WX, X_WH);
        }
        auto res = pp::HalfSpaceContactToPointPair(contact, X_WH, a_is_half_space);
        if (!res) return std::nullopt;
        return PenetrationAsPointPair{ids_[ga], ids_[gb], res->p_WCa, res->p_WCb, res->nhat_BA_W, res->depth};
      }
      if (const auto *ba = std::get_if<Box>(&a), *bb = std::get_if<Box>(&b); ba && bb) {
        // FCL's box-box (boxBox2) behind CalcDistanceFallback, drake_b200_point_pair.h.
        auto res = pp::FclContactToPointPair(pp::BoxBox({ba->width, ba->depth, ba->height}, PoseOf(ga),
                                                        {bb->width, bb->depth, bb->height}, PoseOf(gb)));
        if (!res) return std::nullopt;
        return PenetrationAsPointPair{ids_[ga], ids_[gb], res->p_WCa, res->p_WCb, res->nhat_BA_W, res->depth};
      }
      if (a_is_half_space) {
        // Two half spaces: the reference rejects the pair itself (ScalarSupport<double>, penetration_as_point_pair_callback.cc:
        // 329-338,441-449; :525-555 of its test).
        if (skip_unsupported_point_pairs_) return std::nullopt;
        throw std::logic_error("Penetration queries between shapes 'HalfSpace' and 'HalfSpace' are not supported for scalar "
                               "type double (geometries " + std::to_string(ids_[ga]) + " and " + std::to_string(ids_[gb]) + ")");
      }
      return MprPointPair(ga, gb);   // every other convex pair: libccd's MPR behind fcl::collide
    }
    const int gs = a_is_sphere ? ga : gb, go = a_is_sphere ? gb : ga;   // sphere, other
    const pp::Pose X_WS = PoseOf(gs);
    const pp::Pose X_WO = PoseOf(go);
    const pp::Vec3 p_WSo{X_WS[9], X_WS[10], X_WS[11]};
    const double r = std::get<Sphere>(shapes_[gs]).radius;
    pp::SignedDistanceToPoint d;
    if (const auto* s = std::get_if<Sphere>(&shapes_[go])) {
      d = pp::DistanceToSphere(s->radius, X_WO, p_WSo);
    } else if (const auto* bx = std::get_if<Box>(&shapes_[go])) {
      d = pp::DistanceToBox({bx->width, bx->depth, bx->height}, X_WO, p_WSo);
    } else if (std::holds_alternative<HalfSpace>(shapes_[go])) {
      d = pp::DistanceToHalfSpace(X_WO, p_WSo);
    } else if (const auto* cy = std::get_if<Cylinder>(&shapes_[go])) {
      d = pp::DistanceToCylinder(cy->radius, cy->length, X_WO, p_WSo);
    } else if (const auto* ca = std::get_if<Capsule>(&shapes_[go])) {
      d = pp::DistanceToCapsule(ca->radius, ca->length, X_WO, p_WSo);
    } else {
      return MprPointPair(ga, gb);   // sphere x ellipsoid / mesh: the reference calls FCL too (:273-278)
    }
    auto res = pp::SphereShapePenetration(r, X_WS, X_WO, d);
    if (!res) return std::nullopt;
    if (!a_is_sphere) pp::SwapAAndB(&*res);
    return PenetrationAsPointPair{ids_[ga], ids_[gb], res->p_WCa, res->p_WCb, res->nhat_BA_W, res->depth};
  }

  // The shape as FCL's convex object: primitives as they are, a mesh as its vertex cloud (FCL is given the convex hull of
  // a Mesh or Convex; the support function of the hull is that of the vertices).
  point_pair::ConvexShape ToConvex(int g) const {
    namespace pp = point_pair;
    pp::ConvexShape sh{};
    sh.X_WS = PoseOf(g);
    if (const auto* s = std::get_if<Sphere>(&shapes_[g])) {
      sh.kind = pp::ConvexShape::kSphere;
      sh.dims = {s->radius, 0, 0};
    } else if (const auto* bx = std::get_if<Box>(&shapes_[g])) {
      sh.kind = pp::ConvexShape::kBox;
      sh.dims = {bx->width, bx->depth, bx->height};
    } else if (const auto* cy = std::get_if<Cylinder>(&shapes_[g])) {
      sh.kind = pp::ConvexShape::kCylinder;
      sh.dims = {cy->radius, cy->length, 0};
    } else if (const auto* ca = std::get_if<Capsule>(&shapes_[g])) {
      sh.kind = pp::ConvexShape::kCapsule;
      sh.dims = {ca->radius, ca->length, 0};
    } else if (const auto* el = std::get_if<Ellipsoid>(&shapes_[g])) {
      sh.kind = pp::ConvexShape::kEllipsoid;
      sh.dims = {el->a, el->b, el->c};
    } else {
      const auto* sm = std::get_if<SurfaceMeshShape>(&shapes_[g]);
      const auto* vm = std::get_if<VolumeMeshShape>(&shapes_[g]);
      const std::vector<double>* v = sm ? &sm->vertices : (vm ? &vm->vertices : nullptr);
      if (v == nullptr || v->size() < 3) throw std::logic_error("point-pair fallback: geometry " + std::to_string(ids_[g]) + " has no convex representation");
      sh.kind = pp::ConvexShape::kVertices;
      sh.vertices = v->data();
      sh.num_vertices = v->size() / 3;
    }
    return sh;
  }
  std::optional<PenetrationAsPointPair> MprPointPair(int ga, int gb) const {
    auto res = point_pair::FclContactToPointPair(point_pair::MprPenetration(ToConvex(ga), ToConvex(gb)));
    if (!res) return std::nullopt;
    return PenetrationAsPointPair{ids_[ga], ids_[gb], res->p_WCa, res->p_WCb, res->nhat_BA_W, res->depth};
  }

  // The property checks of MakeRigidRepresentation / MakeCompliantRepresentation (hydroelastic_internal.cc:236-311, the
  // PositiveDouble / NonNegativeDouble validators, and the per-shape Extract calls :326-640), with the reference's
  // messages: a rigid Sphere / Cylinder / Capsule / Ellipsoid needs a positive resolution_hint; a compliant shape a
  // non-negative margin and a positive hydroelastic_modulus, the curved primitives also a positive resolution_hint,
  // a HalfSpace a positive slab_thickness. Shape dimensions must be positive (the shape constructors,
  // shape_specification.cc). Throws std::logic_error like the reference.
 public:
  static void ValidateProperties(const Shape& shape, const ProximityProperties& p) {
    if (p.compliance_type == HydroelasticType::kUndefined) return;
    const bool soft = p.compliance_type == HydroelasticType::kCompliant;
    const char* compliance = soft ? "compliant" : "rigid";
    const char* name = "Shape";
    bool needs_hint = false;
    auto positive_size = [](const char* what, std::initializer_list<double> values) {
      for (double v : values)
        if (!(v > 0) || !std::isfinite(v)) throw std::logic_error(std::string(what) + " dimensions must all be positive and finite");
    };
    if (const auto* s = std::get_if<Sphere>(&shape)) {
      name = "Sphere", needs_hint = true;
      if (!(s->radius >= 0) || !std::isfinite(s->radius)) throw std::logic_error("Sphere radius should be >= 0");
    } else if (const auto* b = std::get_if<Box>(&shape)) {
      name = "Box";
      positive_size("Box", {b->width, b->depth, b->height});
    } else if (const auto* e = std::get_if<Ellipsoid>(&shape)) {
      name = "Ellipsoid", needs_hint = true;
      positive_size("Ellipsoid", {e->a, e->b, e->c});
    } else if (const auto* cy = std::get_if<Cylinder>(&shape)) {
      name = "Cylinder", needs_hint = true;
      positive_size("Cylinder", {cy->radius, cy->length});
    } else if (const auto* ca = std::get_if<Capsule>(&shape)) {
      name = "Capsule", needs_hint = true;
      positive_size("Capsule", {ca->radius, ca->length});
    } else if (std::get_if<HalfSpace>(&shape)) {
      name = "HalfSpace";
    } else if (std::get_if<VolumeMeshShape>(&shape)) {
      name = "Mesh";
    } else if (std::get_if<SurfaceMeshShape>(&shape)) {
      name = soft ? "Convex" : "Mesh";
    }
    auto fail = [&](const char* property, const char* must_be) {
      throw std::logic_error(std::string("Cannot create ") + compliance + " " + name + "; the ('hydroelastic', '" + property +
                             "') property must be " + must_be);
    };
    if (needs_hint && !(p.resolution_hint > 0)) fail("resolution_hint", "positive");
    if (!soft) return;
    if (!(p.margin >= 0)) fail("margin", "non-negative");
    if (std::get_if<HalfSpace>(&shape) && !(p.slab_thickness > 0)) fail("slab_thickness", "positive");
    if (!(p.hydroelastic_modulus > 0)) fail("hydroelastic_modulus", "positive");
    const int strategy = static_cast<int>(p.tessellation_strategy);
    if (strategy != 0 && strategy != 1) throw std::logic_error("tessellation_strategy must be kSingleInteriorVertex or kDenseInteriorVertices");
  }

 private:
  int32_t MakeAsset(const Shape& shape, const ProximityProperties& p) {
    last_elements_.clear();
    last_vertices_per_element_ = 0;
    last_vanished_ = false;
    if (p.compliance_type == HydroelasticType::kUndefined) return -1;
    ValidateProperties(shape, p);
    const bool soft = p.compliance_type == HydroelasticType::kCompliant;
    int32_t asset = -1;
    struct MeshHandle {
      dbh_mesh* h{nullptr};
      ~MeshHandle() {
        if (h) dbh_mesh_destroy(h);
      }
    } vol, surf;
    if (std::get_if<HalfSpace>(&shape)) {
      // Rigid: hydroelastic_internal.cc:323-326; compliant: pressure scale = modulus / slab thickness (:539-553).
      if (soft && !(p.slab_thickness > 0)) throw std::logic_error("a compliant HalfSpace needs a positive slab_thickness");
      Check(dbh_register_half_space(ctx_, soft ? DBH_COMPLIANT : DBH_RIGID,
                                    soft ? p.hydroelastic_modulus / p.slab_thickness : 0.0, soft ? p.margin : 0.0, &asset),
            "register_half_space");
      return asset;
    }
    // A compliant primitive is meshed, and its pressure field built, on the shape INFLATED by the margin (every
    // face moved `margin` along its outward normal): Sphere(radius + margin), Box(size + 2 margin),
    // Ellipsoid(a + margin, ...); Make*PressureField is then given the inflated shape and the margin, so the
    // zero-pressure level set stays on the nominal surface (hydroelastic_internal.cc:412-465,514-536). Rigid
    // representations ignore the margin.
    const double m = soft ? p.margin : 0.0;
    if (const auto* s = std::get_if<Sphere>(&shape)) {
      const double r = s->radius + m;
      CheckMesh(dbh_mesh_make_sphere_volume(r, p.resolution_hint, soft ? static_cast<int>(p.tessellation_strategy) : 0,
                                            &vol.h),
                "MakeSphereVolumeMesh");
      if (soft) CheckMesh(dbh_mesh_set_sphere_pressure(vol.h, r, p.hydroelastic_modulus, m), "MakeSpherePressureField");
    } else if (const auto* b = std::get_if<Box>(&shape)) {
      const double w = b->width + 2.0 * m, d = b->depth + 2.0 * m, h = b->height + 2.0 * m;
      CheckMesh(dbh_mesh_make_box_volume_with_ma(w, d, h, &vol.h), "MakeBoxVolumeMeshWithMa");
      if (soft) CheckMesh(dbh_mesh_set_box_pressure(vol.h, w, d, h, p.hydroelastic_modulus, m), "MakeBoxPressureField");
    } else if (const auto* e = std::get_if<Ellipsoid>(&shape)) {
      const double ea = e->a + m, eb = e->b + m, ec = e->c + m;
      CheckMesh(dbh_mesh_make_ellipsoid_volume(ea, eb, ec, p.resolution_hint,
                                               soft ? static_cast<int>(p.tessellation_strategy) : 0, &vol.h),
                "MakeEllipsoidVolumeMesh");
      if (soft)
        CheckMesh(dbh_mesh_set_ellipsoid_pressure(vol.h, ea, eb, ec, p.hydroelastic_modulus, m), "MakeEllipsoidPressureField");
    } else if (const auto* cy = std::get_if<Cylinder>(&shape)) {
      // MakeCylinderVolumeMeshWithMa on Cylinder(radius + margin, length + 2 margin) (hydroelastic_internal.cc:465-486);
      // rigid: MakeCylinderSurfaceMesh = its boundary (:344-353).
      const double r = cy->radius + m, l = cy->length + 2.0 * m;
      CheckMesh(dbh_mesh_make_cylinder_volume_with_ma(r, l, p.resolution_hint, &vol.h), "MakeCylinderVolumeMeshWithMa");
      if (soft) CheckMesh(dbh_mesh_set_cylinder_pressure(vol.h, r, l, p.hydroelastic_modulus, m), "MakeCylinderPressureField");
    } else if (const auto* ca = std::get_if<Capsule>(&shape)) {
      // MakeCapsuleVolumeMesh on Capsule(radius + margin, length) (hydroelastic_internal.cc:488-508); rigid:
      // MakeCapsuleSurfaceMesh (:355-363).
      const double r = ca->radius + m;
      CheckMesh(dbh_mesh_make_capsule_volume(r, ca->length, p.resolution_hint, &vol.h), "MakeCapsuleVolumeMesh");
      if (soft) CheckMesh(dbh_mesh_set_capsule_pressure(vol.h, r, ca->length, p.hydroelastic_modulus, m), "MakeCapsulePressureField");
    } else if (const auto* v = std::get_if<VolumeMeshShape>(&shape); v && soft && v->pressure.empty()) {
      // The reference's compliant Mesh(.vtk): the tetrahedra as given, MakeVolumeMeshPressureField on them
      // (hydroelastic_internal.cc:579-640). With margin > 0 the reference first inflates the mesh (MakeInflatedMesh, a
      // QP per vertex): not built here.
      if (p.margin > 0) throw std::logic_error("a compliant volume mesh with margin > 0 needs MakeInflatedMesh, which this back end does not provide");
      CheckMesh(dbh_mesh_from_arrays(static_cast<int32_t>(v->vertices.size() / 3), v->vertices.data(),
                                     static_cast<int32_t>(v->tetrahedra.size() / 4), v->tetrahedra.data(), 4, &vol.h),
                "mesh_from_arrays");
      CheckMesh(dbh_mesh_set_volume_mesh_pressure(vol.h, p.hydroelastic_modulus, 0.0), "MakeVolumeMeshPressureField");
    } else if (const auto* cs = std::get_if<SurfaceMeshShape>(&shape); cs && soft) {
      // The reference's compliant Convex (and non-.vtk Mesh): MakeConvexVolumeMesh of the convex hull -- here the closed,
      // outward-wound surface the caller gives IS taken to be that hull (the reference computes it with qhull) -- and
      // MakeVolumeMeshPressureField (hydroelastic_internal.cc:555-577,598-601). margin > 0 needs the hull re-made with
      // the margin (MakeConvexHull): not built here.
      if (p.margin > 0) throw std::logic_error("a compliant convex surface with margin > 0 needs MakeConvexHull, which this back end does not provide");
      MeshHandle hull;
      CheckMesh(dbh_mesh_from_arrays(static_cast<int32_t>(cs->vertices.size() / 3), cs->vertices.data(),
                                     static_cast<int32_t>(cs->triangles.size() / 3), cs->triangles.data(), 3, &hull.h),
                "mesh_from_arrays");
      CheckMesh(dbh_mesh_make_convex_volume(hull.h, &vol.h), "MakeConvexVolumeMesh");
      CheckMesh(dbh_mesh_set_volume_mesh_pressure(vol.h, p.hydroelastic_modulus, 0.0), "MakeVolumeMeshPressureField");
    } else if (const auto* rv = std::get_if<VolumeMeshShape>(&shape); rv && !soft) {
      // The reference's rigid Mesh(.vtk): the boundary surface of the volume mesh (hydroelastic_internal.cc:375-393).
      CheckMesh(dbh_mesh_from_arrays(static_cast<int32_t>(rv->vertices.size() / 3), rv->vertices.data(),
                                     static_cast<int32_t>(rv->tetrahedra.size() / 4), rv->tetrahedra.data(), 4, &vol.h),
                "mesh_from_arrays");
    } else if (const auto* v = std::get_if<VolumeMeshShape>(&shape)) {
      Check(dbh_register_compliant_mesh(ctx_, static_cast<int32_t>(v->vertices.size() / 3), v->vertices.data(),
                                        static_cast<int32_t>(v->tetrahedra.size() / 4), v->tetrahedra.data(),
                                        v->pressure.data(), &asset),
            "register_compliant_mesh");
      last_elements_ = v->tetrahedra;
      last_vertices_per_element_ = 4;
      return asset;
    } else if (const auto* t = std::get_if<SurfaceMeshShape>(&shape)) {
      Check(dbh_register_rigid_mesh(ctx_, static_cast<int32_t>(t->vertices.size() / 3), t->vertices.data(),
                                    static_cast<int32_t>(t->triangles.size() / 3), t->triangles.data(), &asset),
            "register_rigid_mesh");
      last_elements_ = t->triangles;
      last_vertices_per_element_ = 3;
      return asset;
    }
    if (soft) {
      const dbh_status registered = dbh_register_mesh(ctx_, vol.h, &asset);
      const bool primitive = std::get_if<Sphere>(&shape) || std::get_if<Box>(&shape) || std::get_if<Ellipsoid>(&shape) ||
                             std::get_if<Cylinder>(&shape) || std::get_if<Capsule>(&shape);
      if (registered == DBH_ERR_DEGENERATE_FIELD && primitive) {  // vanished (is_vanished above)
        last_vanished_ = true;
        return -1;
      }
      Check(registered, "register_mesh");
    } else {
      CheckMesh(dbh_mesh_convert_volume_to_surface(vol.h, &surf.h), "ConvertVolumeToSurfaceMesh");
      Check(dbh_register_mesh(ctx_, surf.h, &asset), "register_mesh");
    }
    // The connectivity of the registered mesh: welding the vertices of half-space surfaces needs it (Run()).
    {
      const dbh_mesh* m = soft ? vol.h : surf.h;
      int32_t nv = 0, ne = 0, vpe = 0, has_p = 0;
      CheckMesh(dbh_mesh_get_sizes(m, &nv, &ne, &vpe, &has_p), "mesh_get_sizes");
      last_elements_.resize(static_cast<size_t>(ne) * vpe);
      last_vertices_per_element_ = vpe;
      CheckMesh(dbh_mesh_copy(m, nullptr, last_elements_.data(), nullptr), "mesh_copy");
    }
    return asset;
  }

  void AddGeometry(const Shape& shape, const RigidTransformd& X_WG, GeometryId id, const ProximityProperties& props,
                   bool anchored) {
    if (index_.count(id)) throw std::logic_error("GeometryId " + std::to_string(id) + " already registered");
    const int32_t asset = MakeAsset(shape, props);
    index_[id] = static_cast<int>(ids_.size());
    ids_.push_back(id);
    shapes_.push_back(shape);
    elements_.push_back(std::move(last_elements_));
    vertices_per_element_.push_back(last_vertices_per_element_);
    assets_.push_back(asset);
    vanished_.push_back(last_vanished_ ? 1 : 0);
    anchored_.push_back(anchored ? 1 : 0);
    poses_.push_back(X_WG);
    Invalidate();
  }

  std::vector<double> PosesOfOneScene() const {
    if (num_scenes_ != 1) throw std::logic_error("use ComputeContactSurfacesBatched when num_scenes > 1");
    std::vector<double> out(ids_.size() * 12);
    for (size_t g = 0; g < ids_.size(); ++g) std::copy(poses_[g].m.begin(), poses_[g].m.end(), out.begin() + 12 * g);
    return out;
  }

  void EnsureBatch() {
    if (batch_) return;
    const int G = static_cast<int>(ids_.size());
    std::vector<int32_t> offsets(num_scenes_ + 1), assets;
    std::vector<int64_t> ids;
    std::vector<uint8_t> anchored;
    std::vector<int32_t> excluded;
    for (int s = 0; s <= num_scenes_; ++s) offsets[s] = s * G;
    for (int s = 0; s < num_scenes_; ++s) {
      assets.insert(assets.end(), assets_.begin(), assets_.end());
      ids.insert(ids.end(), ids_.begin(), ids_.end());
      anchored.insert(anchored.end(), anchored_.begin(), anchored_.end());
      for (const auto& [a, b] : excluded_) {
        excluded.push_back(s * G + a);
        excluded.push_back(s * G + b);
      }
      for (int g = 0; g < G; ++g) {  // a vanished geometry meets nothing in the hydroelastic queries
        if (!vanished_[g]) continue;
        for (int other = 0; other < G; ++other) {
          if (other == g || (vanished_[other] && other < g)) continue;
          excluded.push_back(s * G + g);
          excluded.push_back(s * G + other);
        }
      }
    }
    Check(dbh_batch_create(ctx_, num_scenes_, offsets.data(), assets.data(), ids.data(), anchored.data(),
                           static_cast<int32_t>(excluded.size() / 2), excluded.data(), &batch_),
          "batch_create");
    // Geometries without a hydroelastic representation (asset -1) still take part in the geometry-level broad phase
    // (FindCollisionCandidates, proximity_engine.cc:867-883): give the library the box FCL's computeLocalAABB gives
    // each shape, so that only pairs whose world boxes overlap reach the point-pair fallback. A HalfSpace stays
    // unbounded (non-finite entry: its pairs always count).
    bool any_unrepresented = false;
    std::vector<double> boxes(static_cast<size_t>(num_scenes_) * G * 6, std::numeric_limits<double>::quiet_NaN());
    for (int g = 0; g < G; ++g) {
      if (assets_[g] >= 0) continue;
      any_unrepresented = true;
      std::array<double, 6> box;
      if (!LocalBox(shapes_[g], &box)) continue;
      for (int s = 0; s < num_scenes_; ++s) std::copy(box.begin(), box.end(), boxes.begin() + (static_cast<size_t>(s) * G + g) * 6);
    }
    if (any_unrepresented) Check(dbh_batch_set_geometry_bounds(batch_, boxes.data()), "set_geometry_bounds");
  }

  // The shape's axis-aligned box in its own frame as {centre, half widths}; false for a HalfSpace.
  static bool LocalBox(const Shape& shape, std::array<double, 6>* box) {
    auto from_vertices = [box](const std::vector<double>& v) {
      if (v.size() < 3) return false;
      for (int k = 0; k < 3; ++k) {
        double lo = v[k], hi = v[k];
        for (size_t i = k; i < v.size(); i += 3) {
          lo = std::min(lo, v[i]);
          hi = std::max(hi, v[i]);
        }
        (*box)[k] = 0.5 * (lo + hi);
        (*box)[3 + k] = 0.5 * (hi - lo);
      }
      return true;
    };
    if (const auto* sp = std::get_if<Sphere>(&shape)) {
      *box = {0, 0, 0, sp->radius, sp->radius, sp->radius};
    } else if (const auto* bx = std::get_if<Box>(&shape)) {
      *box = {0, 0, 0, bx->width / 2, bx->depth / 2, bx->height / 2};
    } else if (const auto* el = std::get_if<Ellipsoid>(&shape)) {
      *box = {0, 0, 0, el->a, el->b, el->c};
    } else if (const auto* cy = std::get_if<Cylinder>(&shape)) {
      *box = {0, 0, 0, cy->radius, cy->radius, cy->length / 2};
    } else if (const auto* ca = std::get_if<Capsule>(&shape)) {
      *box = {0, 0, 0, ca->radius, ca->radius, ca->length / 2 + ca->radius};
    } else if (const auto* vm = std::get_if<VolumeMeshShape>(&shape)) {
      return from_vertices(vm->vertices);
    } else if (const auto* sm = std::get_if<SurfaceMeshShape>(&shape)) {
      return from_vertices(sm->vertices);
    } else {
      return false;
    }
    return true;
  }

  std::vector<std::vector<ContactSurface>> Run(HydroelasticContactRepresentation rep, int query_kind,
                                               const std::vector<double>& X_WGs) {
    std::vector<std::vector<ContactSurface>> out(num_scenes_);
    if (ids_.empty()) return out;
    EnsureBatch();
    for (int attempt = 0;; ++attempt) {
      Check(dbh_batch_set_poses_host(batch_, X_WGs.data()), "set_poses");
      Check(dbh_batch_query(batch_, static_cast<int32_t>(rep), query_kind, nullptr), "query");
      const dbh_status s = dbh_batch_sync(batch_);
      if (s == DBH_ERR_CAPACITY && attempt < 64) {  // first queries size the device work lists
        Check(dbh_batch_grow_capacity(batch_), "grow_capacity");
        continue;
      }
      Check(s, "ComputeContactSurfaces");
      break;
    }
    dbh_result_sizes sz{};
    Check(dbh_batch_result_sizes(batch_, &sz), "result_sizes");
    const size_t ns = sz.num_surfaces, nf = sz.num_faces, nv = sz.num_vertices, nfd = sz.face_data_ints;
    std::vector<int32_t> scene(ns), face_data(nfd);
    std::vector<int64_t> id_M(ns), id_N(ns), foff(ns + 1), voff(ns + 1);
    std::vector<double> tarea(ns), scent(3 * ns), verts(3 * nv), pres(nv), areas(nf), normals(3 * nf), cents(3 * nf),
        ecent(nf), gM(3 * nf), gN(3 * nf);
    std::vector<uint8_t> flags(ns);
    bool any_half_space = false;
    for (const Shape& sh : shapes_) any_half_space = any_half_space || std::holds_alternative<HalfSpace>(sh);
    std::vector<int32_t> origin(any_half_space ? nv : 0), elem_M(any_half_space ? nf : 0), elem_N(any_half_space ? nf : 0);
    dbh_result_buffers buf{};
    if (any_half_space) {
      buf.vertex_origin = origin.data();
      buf.elem_M = elem_M.data();
      buf.elem_N = elem_N.data();
    }
    buf.surf_scene = scene.data();
    buf.surf_id_M = id_M.data();
    buf.surf_id_N = id_N.data();
    buf.surf_face_offset = foff.data();
    buf.surf_vertex_offset = voff.data();
    buf.surf_total_area = tarea.data();
    buf.surf_centroid = scent.data();
    buf.surf_flags = flags.data();
    buf.vertices = verts.data();
    buf.pressures = pres.data();
    buf.face_data = face_data.data();
    buf.areas = areas.data();
    buf.normals = normals.data();
    buf.centroids = cents.data();
    buf.pressure_at_centroid = ecent.data();
    buf.grad_eM = gM.data();
    buf.grad_eN = gN.data();
    Check(dbh_batch_copy_results(batch_, &buf), "copy_results");
    const bool poly = rep == HydroelasticContactRepresentation::kPolygon;
    for (size_t k = 0; k < ns; ++k) {
      ContactSurface cs;
      cs.id_M_ = id_M[k];
      cs.id_N_ = id_N[k];
      cs.rep_ = rep;
      const size_t f0 = foff[k], f1 = foff[k + 1], v0 = voff[k], v1 = voff[k + 1];
      // The concatenated face_data of a surface starts at f0 + v0 (polygon: one count per face plus
      // one index per vertex) or 3 * f0 (triangles).
      const size_t d0 = poly ? f0 + v0 : 3 * f0, d1 = poly ? f1 + v1 : 3 * f1;
      cs.vertices_.assign(verts.begin() + 3 * v0, verts.begin() + 3 * v1);
      cs.pressures_.assign(pres.begin() + v0, pres.begin() + v1);
      cs.face_data_.assign(face_data.begin() + d0, face_data.begin() + d1);
      cs.areas_.assign(areas.begin() + f0, areas.begin() + f1);
      cs.normals_.assign(normals.begin() + 3 * f0, normals.begin() + 3 * f1);
      cs.centroids_.assign(cents.begin() + 3 * f0, cents.begin() + 3 * f1);
      cs.e_centroid_.assign(ecent.begin() + f0, ecent.begin() + f1);
      cs.has_grad_M_ = flags[k] & 1;
      cs.has_grad_N_ = flags[k] & 2;
      if (cs.has_grad_M_) cs.grad_eM_.assign(gM.begin() + 3 * f0, gM.begin() + 3 * f1);
      if (cs.has_grad_N_) cs.grad_eN_.assign(gN.begin() + 3 * f0, gN.begin() + 3 * f1);
      cs.total_area_ = tarea[k];
      cs.centroid_ = {scent[3 * k], scent[3 * k + 1], scent[3 * k + 2]};
      if (any_half_space) {
        const int gM = index_.at(cs.id_M_), gN = index_.at(cs.id_N_);
        const bool hM = std::holds_alternative<HalfSpace>(shapes_[gM]), hN = std::holds_alternative<HalfSpace>(shapes_[gN]);
        if (hM || hN) {
          const int gmesh = hM ? gN : gM;
          WeldHalfSpaceSurface(&cs, origin.data() + v0, elem_M.data() + f0, elem_N.data() + f0, elements_[gmesh],
                               vertices_per_element_[gmesh]);
        }
      }
      out[scene[k]].push_back(std::move(cs));
    }
    return out;
  }

  // The reference shares the vertices of a half-space surface between its polygons: every vertex cut from the same
  // mesh edge (cut_edges, mesh_plane_intersection.cc:205-231; edges_to_newly_created_vertices, mesh_half_space_
  // intersection.cc:52-101) or copied from the same mesh vertex (vertices_to_newly_created_vertices) exists once. The
  // kernels give every polygon its own vertices plus, per vertex, where it comes from (dbh_result_buffers::
  // vertex_origin); this merges them, first occurrence in face order keeping its position and pressure (as the
  // reference's maps do in ITS visiting order: positions can differ from the reference's in the last bits, counts
  // and connectivity cannot). Face quantities are untouched.
  static void WeldHalfSpaceSurface(ContactSurface* cs, const int32_t* origin, const int32_t* elem_M, const int32_t* elem_N,
                                   const std::vector<int32_t>& elements, int vertices_per_element) {
    const size_t nv = cs->pressures_.size();
    const bool tri = cs->rep_ == HydroelasticContactRepresentation::kTriangle;
    std::map<std::array<int64_t, 3>, int32_t> key_to_new;
    std::vector<int32_t> new_index(nv, -1);
    std::vector<double> verts, pres;
    size_t i = 0;
    const size_t nf = cs->areas_.size();
    for (size_t f = 0; f < nf; ++f) {
      const int n = tri ? 3 : cs->face_data_[i];
      int32_t* idx = cs->face_data_.data() + (tri ? 3 * f : i + 1);
      const int32_t elem = elem_M[f] >= 0 ? elem_M[f] : elem_N[f];
      const int32_t* ev = elements.data() + static_cast<size_t>(elem) * vertices_per_element;
      for (int c = 0; c < n; ++c) {
        const int32_t v = idx[c];
        if (new_index[v] < 0) {
          const int32_t o = origin[v];
          std::array<int64_t, 3> key;
          if (o < 0) {
            key = {2, v, 0};   // the centroid vertex of a triangle fan: never shared
          } else if (o < 16) {
            key = {0, ev[o], 0};
          } else {
            const int64_t a = ev[(o - 16) / 4], b = ev[(o - 16) % 4];
            key = {1, std::min(a, b), std::max(a, b)};
          }
          auto it = key_to_new.find(key);
          if (it == key_to_new.end()) {
            it = key_to_new.emplace(key, static_cast<int32_t>(pres.size())).first;
            verts.insert(verts.end(), cs->vertices_.begin() + 3 * v, cs->vertices_.begin() + 3 * v + 3);
            pres.push_back(cs->pressures_[v]);
          }
          new_index[v] = it->second;
        }
        idx[c] = new_index[v];
      }
      if (!tri) i += n + 1;
    }
    cs->vertices_ = std::move(verts);
    cs->pressures_ = std::move(pres);
  }

  dbh_context* ctx_{nullptr};
  dbh_batch* batch_{nullptr};
  int num_scenes_{1};
  std::vector<GeometryId> ids_;
  std::vector<Shape> shapes_;
  std::vector<std::vector<int32_t>> elements_;   // per geometry: connectivity of its registered mesh (empty: none)
  std::vector<int> vertices_per_element_;
  std::vector<int32_t> last_elements_;
  int last_vertices_per_element_{0};
  bool skip_unsupported_point_pairs_{false};
  const double* scene_poses_{nullptr};   // set only inside ComputeContactSurfacesWithFallbackBatched
  std::vector<int32_t> assets_;
  std::vector<uint8_t> vanished_;   // per geometry: is_vanished()
  bool last_vanished_{false};
  std::vector<uint8_t> anchored_;
  std::vector<RigidTransformd> poses_;
  std::vector<std::pair<int, int>> excluded_;
  std::unordered_map<GeometryId, int> index_;
};

}  // namespace drake_b200

#endif  // DRAKE_B200_PROXIMITY_ENGINE_H_
