// drake_b200_deformable_contact.h — host-side C++ mirror of the reference's deformable-contact interface, over the
// C-ABI in drake_b200_hydro.h. Header-only C++17, no Eigen. Mirrors, name for name:
//   geometry::internal::deformable::Geometries            geometry/proximity/deformable_contact_internal.h
//     AddRigidGeometry (here: the rigid geometry's volume mesh + pressure field, i.e. the result of
//       MakeMeshRepresentation, deformable_contact_geometries.h), AddDeformableGeometry, UpdateRigidWorldPose,
//       UpdateDeformableVertexPositions, ComputeDeformableContact(collision filter)
//   geometry::internal::DeformableContact<double>, DeformableContactSurface<double>, ContactParticipation
//                                                          geometry/query_results/deformable_contact.{h,cc}
// What runs where: the two intersections (deforming surface x rigid pressure volume; deforming volume x deforming
// volume) and the per-contact-point data run on the GPU (dbh_context_update_vertices, dbh_batch_query,
// dbh_batch_compute_deformable_contact); the index bookkeeping that needs the deformable meshes' connectivity --
// contact vertex indexes, participating vertices, is_element_inverted (deformable_mesh_intersection.cc:127-170,
// deformable_field_intersection.cc:27-42) -- is a gather here. Failures are std::logic_error / std::runtime_error.
#ifndef DRAKE_B200_DEFORMABLE_CONTACT_H_
#define DRAKE_B200_DEFORMABLE_CONTACT_H_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#include "drake_b200_hydro.h"

namespace drake_b200 {
namespace deformable {

using GeometryId = int64_t;
using Vector3d = std::array<double, 3>;
using Vector4d = std::array<double, 4>;
using Vector3i = std::array<int, 3>;
using Vector4i = std::array<int, 4>;
using Matrix3d = std::array<double, 9>;   // column-major, as Eigen stores it

struct RigidTransformd {   // rotation column-major then translation (math/rigid_transform.h:1004-1007)
  std::array<double, 12> m{1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  static RigidTransformd FromTranslation(double x, double y, double z) {
    RigidTransformd X;
    X.m[9] = x, X.m[10] = y, X.m[11] = z;
    return X;
  }
};

struct VolumeMeshd {            // VolumeMesh<double>
  std::vector<double> vertices;  // [V*3]
  std::vector<int> tetrahedra;   // [T*4]
  int num_vertices() const { return static_cast<int>(vertices.size() / 3); }
  int num_elements() const { return static_cast<int>(tetrahedra.size() / 4); }
};
struct TriangleSurfaceMeshd {   // TriangleSurfaceMesh<double>
  std::vector<double> vertices;  // [V*3]
  std::vector<int> triangles;    // [T*3]
  int num_vertices() const { return static_cast<int>(vertices.size() / 3); }
};
struct PolygonSurfaceMeshd {    // PolygonSurfaceMesh<double>: what a DeformableContactSurface keeps of it
  std::vector<double> vertices;   // [V*3], World
  std::vector<int> face_data;     // |n|v0..vn-1| per face
  std::vector<double> areas;      // [F]
  std::vector<Vector3d> face_normals, element_centroids;
  int num_faces() const { return static_cast<int>(areas.size()); }
};

// ContactParticipation (deformable_contact.h:22-63).
class ContactParticipation {
 public:
  explicit ContactParticipation(int num_vertices) : participation_(num_vertices, false) {}
  void Participate(const std::unordered_set<int>& vertices) {
    for (int v : vertices) {
      if (v < 0 || v >= static_cast<int>(participation_.size())) throw std::logic_error("ContactParticipation: bad vertex index");
      if (!participation_[v]) {
        ++num_vertices_in_contact_;
        participation_[v] = true;
      }
    }
  }
  int num_vertices_in_contact() const { return num_vertices_in_contact_; }
  int num_vertices() const { return static_cast<int>(participation_.size()); }
  bool participates(int v) const { return participation_.at(v); }

 private:
  std::vector<bool> participation_;
  int num_vertices_in_contact_{0};
};

// DeformableContactSurface<double> (deformable_contact.h:85-330). Geometry A is always deformable.
class DeformableContactSurface {
 public:
  GeometryId id_A() const { return id_A_; }
  GeometryId id_B() const { return id_B_; }
  int num_contact_points() const { return contact_mesh_W_.num_faces(); }
  const PolygonSurfaceMeshd& contact_mesh_W() const { return contact_mesh_W_; }
  const std::vector<Vector3d>& contact_points_W() const { return contact_points_W_; }
  const std::vector<Vector3d>& nhats_W() const { return nhats_W_; }
  const std::vector<Matrix3d>& R_WCs() const { return R_WCs_; }
  bool is_B_deformable() const { return is_B_deformable_; }
  // Deformable x rigid.
  const std::vector<double>& pressures() const { return Need(!is_B_deformable_, pressures_); }
  const std::vector<Vector3d>& pressure_gradients_W() const { return Need(!is_B_deformable_, pressure_gradients_W_); }
  const std::vector<bool>& is_element_inverted() const { return Need(!is_B_deformable_, is_element_inverted_); }
  const std::vector<Vector3i>& tri_contact_vertex_indexes_A() const { return Need(!is_B_deformable_, tri_indexes_A_); }
  const std::vector<Vector3d>& tri_barycentric_coordinates_A() const { return Need(!is_B_deformable_, tri_bary_A_); }
  // Deformable x deformable.
  const std::vector<double>& signed_distances() const { return Need(is_B_deformable_, signed_distances_); }
  const std::vector<Vector4i>& tet_contact_vertex_indexes_A() const { return Need(is_B_deformable_, tet_indexes_A_); }
  const std::vector<Vector4d>& tet_barycentric_coordinates_A() const { return Need(is_B_deformable_, tet_bary_A_); }
  const std::vector<Vector4i>& contact_vertex_indexes_B() const { return Need(is_B_deformable_, tet_indexes_B_); }
  const std::vector<Vector4d>& barycentric_coordinates_B() const { return Need(is_B_deformable_, tet_bary_B_); }

 private:
  friend class Geometries;
  template <class T>
  static const T& Need(bool ok, const T& v) {
    if (!ok) throw std::logic_error("DeformableContactSurface: this accessor belongs to the other kind of surface");
    return v;
  }
  GeometryId id_A_{0}, id_B_{0};
  bool is_B_deformable_{false};
  PolygonSurfaceMeshd contact_mesh_W_;
  std::vector<Vector3d> contact_points_W_, nhats_W_;
  std::vector<Matrix3d> R_WCs_;
  std::vector<double> pressures_, signed_distances_;
  std::vector<Vector3d> pressure_gradients_W_;
  std::vector<bool> is_element_inverted_;
  std::vector<Vector3i> tri_indexes_A_;
  std::vector<Vector3d> tri_bary_A_;
  std::vector<Vector4i> tet_indexes_A_, tet_indexes_B_;
  std::vector<Vector4d> tet_bary_A_, tet_bary_B_;
};

// DeformableContact<double> (deformable_contact.h:335-510).
class DeformableContact {
 public:
  void RegisterDeformableGeometry(GeometryId id, int num_vertices) { participations_.emplace(id, ContactParticipation(num_vertices)); }
  bool IsRegistered(GeometryId id) const { return participations_.count(id) > 0; }
  const ContactParticipation& contact_participation(GeometryId id) const { return participations_.at(id); }
  const std::vector<DeformableContactSurface>& contact_surfaces() const { return contact_surfaces_; }

 private:
  friend class Geometries;
  std::map<GeometryId, ContactParticipation> participations_;
  std::vector<DeformableContactSurface> contact_surfaces_;
};

// geometry::internal::deformable::Geometries (deformable_contact_internal.h:40-190).
class Geometries {
 public:
  explicit Geometries(int device = 0) {
    if (dbh_context_create(device, &ctx_) != DBH_OK) throw std::runtime_error(std::string("drake_b200: ") + dbh_last_error());
  }
  ~Geometries() {
    if (batch_) dbh_batch_destroy(batch_);
    if (ctx_) dbh_context_destroy(ctx_);
  }
  Geometries(const Geometries&) = delete;
  Geometries& operator=(const Geometries&) = delete;

  bool is_rigid(GeometryId id) const { return rigid_.count(id) > 0; }
  bool is_deformable(GeometryId id) const { return deformable_.count(id) > 0; }

  // The rigid geometry as deformable contact sees it (RigidGeometry, deformable_contact_geometries.h:150-190): a volume
  // mesh with a pressure field, posed in World.
  void AddRigidGeometry(GeometryId id, const VolumeMeshd& mesh_R, const std::vector<double>& pressure, const RigidTransformd& X_WG) {
    CheckNew(id);
    Rigid r;
    Check(dbh_register_compliant_mesh(ctx_, mesh_R.num_vertices(), mesh_R.vertices.data(), mesh_R.num_elements(),
                                      mesh_R.tetrahedra.data(), pressure.data(), &r.asset));
    r.X_WG = X_WG;
    rigid_.emplace(id, r);
    dirty_ = true;
  }

  // AddDeformableGeometry (deformable_contact_internal.cc:57-68). The signed distance field is the reference's
  // ApproximateSignedDistanceField (deformable_contact_geometries.cc:21-47): 0 on boundary vertices, minus the distance
  // to the surface mesh inside.
  void AddDeformableGeometry(GeometryId id, VolumeMeshd volume_mesh, TriangleSurfaceMeshd surface_mesh,
                             std::vector<int> surface_index_to_volume_index, std::vector<int> surface_tri_to_volume_tet) {
    CheckNew(id);
    Deformable d;
    d.signed_distance = ApproximateSignedDistanceField(volume_mesh, surface_mesh, surface_index_to_volume_index);
    Check(dbh_register_compliant_mesh(ctx_, volume_mesh.num_vertices(), volume_mesh.vertices.data(), volume_mesh.num_elements(),
                                      volume_mesh.tetrahedra.data(), d.signed_distance.data(), &d.volume_asset));
    Check(dbh_register_rigid_mesh(ctx_, surface_mesh.num_vertices(), surface_mesh.vertices.data(),
                                  static_cast<int>(surface_mesh.triangles.size() / 3), surface_mesh.triangles.data(), &d.surface_asset));
    d.volume = std::move(volume_mesh);
    d.surface = std::move(surface_mesh);
    d.surface_index_to_volume_index = std::move(surface_index_to_volume_index);
    d.surface_tri_to_volume_tet = std::move(surface_tri_to_volume_tet);
    deformable_.emplace(id, std::move(d));
    dirty_ = true;
  }

  void UpdateRigidWorldPose(GeometryId id, const RigidTransformd& X_WG) {
    auto it = rigid_.find(id);
    if (it != rigid_.end()) it->second.X_WG = X_WG;
  }

  // UpdateDeformableVertexPositions(id, q_WV, q_WS) (deformable_contact_internal.cc:70-76): both meshes re-posed on the
  // device in one call.
  void UpdateDeformableVertexPositions(GeometryId id, const std::vector<double>& q_WV, const std::vector<double>& q_WS) {
    auto it = deformable_.find(id);
    if (it == deformable_.end()) return;
    Deformable& d = it->second;
    if (q_WV.size() != d.volume.vertices.size() || q_WS.size() != d.surface.vertices.size())
      throw std::logic_error("UpdateDeformableVertexPositions: wrong number of positions");
    d.volume.vertices = q_WV;
    d.surface.vertices = q_WS;
    const int32_t assets[2] = {d.volume_asset, d.surface_asset};
    std::vector<double> both(q_WV);
    both.insert(both.end(), q_WS.begin(), q_WS.end());
    Check(dbh_context_update_vertices(ctx_, 2, assets, both.data()));
  }

  // ComputeDeformableContact(collision_filter) (deformable_contact_internal.cc:78-165).
  DeformableContact ComputeDeformableContact(const std::function<bool(GeometryId, GeometryId)>& can_collide_with = {}) {
    DeformableContact result;
    for (const auto& [id, d] : deformable_) result.RegisterDeformableGeometry(id, d.volume.num_vertices());
    if (deformable_.empty()) return result;
    if (dirty_ || filter_dirty_ || can_collide_with) Rebuild(can_collide_with);
    // Poses: rigid geometries where they are; deformable meshes live in World.
    std::vector<double> X(12 * geoms_.size());
    for (size_t g = 0; g < geoms_.size(); ++g) {
      const RigidTransformd I;
      const RigidTransformd& x = geoms_[g].role == kRigid ? rigid_.at(geoms_[g].id).X_WG : I;
      std::copy(x.m.begin(), x.m.end(), X.begin() + 12 * g);
    }
    Check(dbh_batch_set_poses_host(batch_, X.data()));
    for (int attempt = 0;; ++attempt) {
      Check(dbh_batch_query(batch_, DBH_POLYGON, DBH_QUERY_STRICT, nullptr));
      const dbh_status s = dbh_batch_sync(batch_);
      if (s == DBH_OK) break;
      if (s != DBH_ERR_CAPACITY || attempt > 48) Check(s);
      Check(dbh_batch_grow_capacity(batch_));
    }
    Check(dbh_batch_compute_deformable_contact(batch_));
    Check(dbh_batch_sync(batch_));
    dbh_result_sizes sz;
    Check(dbh_batch_result_sizes(batch_, &sz));
    const size_t S = sz.num_surfaces, F = sz.num_faces, V = sz.num_vertices;
    std::vector<int64_t> id_M(S), id_N(S), f_off(S + 1), v_off(S + 1);
    std::vector<double> verts(3 * V), areas(F), normals(3 * F), centroids(3 * F), rows(24 * F);
    std::vector<int32_t> face_data(sz.face_data_ints), ea(F), eb(F);
    dbh_result_buffers out{};
    out.surf_id_M = id_M.data(), out.surf_id_N = id_N.data();
    out.surf_face_offset = f_off.data(), out.surf_vertex_offset = v_off.data();
    out.vertices = verts.data(), out.face_data = face_data.data();
    out.areas = areas.data(), out.normals = normals.data(), out.centroids = centroids.data();
    Check(dbh_batch_copy_results(batch_, &out));
    int64_t nf = 0;
    Check(dbh_batch_copy_deformable_contact_async(batch_, rows.data(), ea.data(), eb.data(), &nf));
    Check(dbh_batch_sync(batch_));
    for (size_t s = 0; s < S; ++s) {
      const Geom& gm = geoms_[batch_id_to_geom_.at(id_M[s])];
      const Geom& gn = geoms_[batch_id_to_geom_.at(id_N[s])];
      DeformableContactSurface cs;
      const size_t f0 = f_off[s], f1 = f_off[s + 1], v0 = v_off[s], v1 = v_off[s + 1];
      cs.contact_mesh_W_.vertices.assign(verts.begin() + 3 * v0, verts.begin() + 3 * v1);
      cs.contact_mesh_W_.face_data.assign(face_data.begin() + f0 + v0, face_data.begin() + f1 + v1);
      cs.contact_mesh_W_.areas.assign(areas.begin() + f0, areas.begin() + f1);
      for (size_t f = f0; f < f1; ++f) {
        const double* r = rows.data() + 24 * f;
        cs.contact_mesh_W_.face_normals.push_back({normals[3 * f], normals[3 * f + 1], normals[3 * f + 2]});
        cs.contact_mesh_W_.element_centroids.push_back({centroids[3 * f], centroids[3 * f + 1], centroids[3 * f + 2]});
        cs.contact_points_W_.push_back({r[0], r[1], r[2]});
        cs.nhats_W_.push_back({r[3], r[4], r[5]});
        Matrix3d R;
        std::copy(r + 6, r + 15, R.begin());
        cs.R_WCs_.push_back(R);
      }
      const bool surface_pair = gm.role == kDeformableSurface || gn.role == kDeformableSurface;
      if (surface_pair) {   // AddDeformableRigidContactSurface (deformable_mesh_intersection.cc:85-183)
        const Geom& gd = gm.role == kDeformableSurface ? gm : gn;
        const Geom& gr = gm.role == kDeformableSurface ? gn : gm;
        const Deformable& d = deformable_.at(gd.id);
        cs.id_A_ = gd.id, cs.id_B_ = gr.id;
        // The stored ContactSurface has the smaller batch id as M; the mesh the reference keeps has its normals out of
        // the deformable surface, i.e. -nhat (deformable_contact.cc:58).
        std::unordered_set<int> participating;
        for (size_t f = f0; f < f1; ++f) {
          const double* r = rows.data() + 24 * f;
          cs.contact_mesh_W_.face_normals[f - f0] = {-r[3], -r[4], -r[5]};
          cs.pressures_.push_back(r[15]);
          cs.pressure_gradients_W_.push_back({r[16], r[17], r[18]});
          cs.tri_bary_A_.push_back({r[19], r[20], r[21]});
          const int tri = eb[f];
          Vector3i idx;
          for (int k = 0; k < 3; ++k) {
            idx[k] = d.surface_index_to_volume_index.at(d.surface.triangles.at(3 * tri + k));
            participating.insert(idx[k]);
          }
          cs.tri_indexes_A_.push_back(idx);
          cs.is_element_inverted_.push_back(TetrahedronVolume(d.volume, d.surface_tri_to_volume_tet.at(tri)) < 0);
        }
        result.participations_.at(gd.id).Participate(participating);
      } else {   // AddDeformableDeformableContactSurface (deformable_field_intersection.cc:46-124): A = N's, B = M's
        const Deformable& dA = deformable_.at(gn.id);
        const Deformable& dB = deformable_.at(gm.id);
        cs.id_A_ = gn.id, cs.id_B_ = gm.id;
        cs.is_B_deformable_ = true;
        std::unordered_set<int> pA, pB;
        for (size_t f = f0; f < f1; ++f) {
          const double* r = rows.data() + 24 * f;
          cs.signed_distances_.push_back(r[15]);
          cs.tet_bary_B_.push_back({r[16], r[17], r[18], r[19]});
          cs.tet_bary_A_.push_back({r[20], r[21], r[22], r[23]});
          Vector4i iA, iB;
          for (int k = 0; k < 4; ++k) {
            iB[k] = dB.volume.tetrahedra.at(4 * ea[f] + k);
            iA[k] = dA.volume.tetrahedra.at(4 * eb[f] + k);
            pA.insert(iA[k]), pB.insert(iB[k]);
          }
          cs.tet_indexes_A_.push_back(iA);
          cs.tet_indexes_B_.push_back(iB);
        }
        result.participations_.at(gn.id).Participate(pA);
        result.participations_.at(gm.id).Participate(pB);
      }
      result.contact_surfaces_.push_back(std::move(cs));
    }
    return result;
  }

 private:
  enum Role { kRigid, kDeformableSurface, kDeformableVolume };
  struct Rigid {
    int32_t asset{-1};
    RigidTransformd X_WG;
  };
  struct Deformable {
    int32_t volume_asset{-1}, surface_asset{-1};
    VolumeMeshd volume;
    TriangleSurfaceMeshd surface;
    std::vector<int> surface_index_to_volume_index, surface_tri_to_volume_tet;
    std::vector<double> signed_distance;
  };
  struct Geom {
    Role role;
    GeometryId id;
    int64_t batch_id;
  };

  static void Check(dbh_status s) {
    if (s == DBH_OK) return;
    const std::string msg = std::string("drake_b200: ") + dbh_last_error();
    if (s == DBH_ERR_INVALID_ARGUMENT || s == DBH_ERR_DEGENERATE_FIELD) throw std::logic_error(msg);
    throw std::runtime_error(msg);
  }
  void CheckNew(GeometryId id) const {
    if (rigid_.count(id) || deformable_.count(id)) throw std::logic_error("Geometries: GeometryId " + std::to_string(id) + " is already registered");
  }
  static Vector3d Vtx(const std::vector<double>& v, int i) { return {v[3 * i], v[3 * i + 1], v[3 * i + 2]}; }
  static Vector3d Sub(Vector3d a, Vector3d b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
  static double Dot(Vector3d a, Vector3d b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  static Vector3d Cross(Vector3d a, Vector3d b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
  // VolumeMesh::CalcTetrahedronVolume (volume_mesh.h:190-205).
  static double TetrahedronVolume(const VolumeMeshd& m, int e) {
    const Vector3d a = Vtx(m.vertices, m.tetrahedra[4 * e]), b = Vtx(m.vertices, m.tetrahedra[4 * e + 1]);
    const Vector3d c = Vtx(m.vertices, m.tetrahedra[4 * e + 2]), d = Vtx(m.vertices, m.tetrahedra[4 * e + 3]);
    return Dot(Cross(Sub(b, a), Sub(c, a)), Sub(d, a)) / 6.0;
  }
  // Distance from a point to a triangle (closest-point regions; what CalcDistanceToSurfaceMesh reduces to per triangle).
  static double DistanceToTriangle(Vector3d p, Vector3d a, Vector3d b, Vector3d c) {
    const Vector3d ab = Sub(b, a), ac = Sub(c, a), ap = Sub(p, a);
    const double d1 = Dot(ab, ap), d2 = Dot(ac, ap);
    auto len = [](Vector3d v) { return std::sqrt(Dot(v, v)); };
    if (d1 <= 0 && d2 <= 0) return len(ap);
    const Vector3d bp = Sub(p, b);
    const double d3 = Dot(ab, bp), d4 = Dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return len(bp);
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
      const double v = d1 / (d1 - d3);
      return len(Sub(ap, {v * ab[0], v * ab[1], v * ab[2]}));
    }
    const Vector3d cp = Sub(p, c);
    const double d5 = Dot(ab, cp), d6 = Dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return len(cp);
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
      const double w = d2 / (d2 - d6);
      return len(Sub(ap, {w * ac[0], w * ac[1], w * ac[2]}));
    }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
      const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
      const Vector3d bc = Sub(c, b);
      return len(Sub(bp, {w * bc[0], w * bc[1], w * bc[2]}));
    }
    const Vector3d n = Cross(ab, ac);
    return std::fabs(Dot(n, ap)) / len(n);
  }
  static std::vector<double> ApproximateSignedDistanceField(const VolumeMeshd& mesh, const TriangleSurfaceMeshd& surface,
                                                            const std::vector<int>& surface_index_to_volume_index) {
    std::vector<bool> boundary(mesh.num_vertices(), false);
    for (int v : surface_index_to_volume_index) boundary.at(v) = true;
    std::vector<double> sd(mesh.num_vertices(), 0.0);
    const int nt = static_cast<int>(surface.triangles.size() / 3);
    for (int v = 0; v < mesh.num_vertices(); ++v) {
      if (boundary[v]) continue;
      double best = std::numeric_limits<double>::infinity();
      for (int t = 0; t < nt; ++t)
        best = std::min(best, DistanceToTriangle(Vtx(mesh.vertices, v), Vtx(surface.vertices, surface.triangles[3 * t]),
                                                 Vtx(surface.vertices, surface.triangles[3 * t + 1]),
                                                 Vtx(surface.vertices, surface.triangles[3 * t + 2])));
      sd[v] = -best;
    }
    return sd;
  }

  // One scene: every rigid volume, every deformable surface and every deformable volume; every pair excluded except
  // (deformable surface, rigid volume) and (deformable volume, deformable volume) of different bodies that the filter
  // lets collide. Deformable volumes get batch ids in REVERSED order of their GeometryIds, so that the geometry with the
  // larger id is the intersector's first field, as in the reference (deformable_field_intersection.cc:53).
  void Rebuild(const std::function<bool(GeometryId, GeometryId)>& can_collide_with) {
    if (batch_) dbh_batch_destroy(batch_);
    batch_ = nullptr;
    geoms_.clear();
    batch_id_to_geom_.clear();
    std::vector<int32_t> assets;
    int64_t next = 1;
    for (const auto& [id, r] : rigid_) {
      geoms_.push_back({kRigid, id, next++});
      assets.push_back(r.asset);
    }
    for (const auto& [id, d] : deformable_) {
      geoms_.push_back({kDeformableSurface, id, next++});
      assets.push_back(d.surface_asset);
    }
    const int64_t top = next + static_cast<int64_t>(deformable_.size());
    int64_t k = 0;
    for (const auto& [id, d] : deformable_) {   // std::map: ascending ids -> descending batch ids
      geoms_.push_back({kDeformableVolume, id, top - (k++)});
      assets.push_back(d.volume_asset);
    }
    std::vector<int64_t> ids;
    for (size_t g = 0; g < geoms_.size(); ++g) {
      ids.push_back(geoms_[g].batch_id);
      batch_id_to_geom_[geoms_[g].batch_id] = g;
    }
    std::vector<int32_t> excluded;
    for (size_t i = 0; i < geoms_.size(); ++i)
      for (size_t j = i + 1; j < geoms_.size(); ++j) {
        const Geom &a = geoms_[i], &b = geoms_[j];
        const bool surface_rigid = (a.role == kRigid && b.role == kDeformableSurface) || (a.role == kDeformableSurface && b.role == kRigid);
        const bool volume_volume = a.role == kDeformableVolume && b.role == kDeformableVolume && a.id != b.id;
        bool wanted = surface_rigid || volume_volume;
        if (wanted && can_collide_with) wanted = can_collide_with(a.id, b.id);
        if (!wanted) {
          excluded.push_back(static_cast<int32_t>(i));
          excluded.push_back(static_cast<int32_t>(j));
        }
      }
    const int32_t offsets[2] = {0, static_cast<int32_t>(geoms_.size())};
    std::vector<uint8_t> anchored(geoms_.size(), 0);
    Check(dbh_batch_create(ctx_, 1, offsets, assets.data(), ids.data(), anchored.data(), static_cast<int32_t>(excluded.size() / 2),
                           excluded.empty() ? nullptr : excluded.data(), &batch_));
    dirty_ = false;
    filter_dirty_ = static_cast<bool>(can_collide_with);   // a caller-supplied filter may change between calls
  }

  dbh_context* ctx_{nullptr};
  dbh_batch* batch_{nullptr};
  std::map<GeometryId, Rigid> rigid_;
  std::map<GeometryId, Deformable> deformable_;
  std::vector<Geom> geoms_;
  std::map<int64_t, size_t> batch_id_to_geom_;
  bool dirty_{true}, filter_dirty_{false};
};

}  // namespace deformable
}  // namespace drake_b200

#endif  // DRAKE_B200_DEFORMABLE_CONTACT_H_
