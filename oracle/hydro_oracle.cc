// ORACLE — TEST INFRASTRUCTURE ONLY (see ho_math.h).
// C entry points (for ctypes) over the CPU restatement, plus the scene-level
// loop restating ProximityEngine::Impl::ComputeContactSurfaces
// (geometry/proximity_engine.cc:905-934) and the pair dispatcher
// ContactCalculator::MaybeMakeContactSurface (hydroelastic_calculator.cc:77-156).
//
// Parity status: pinned by the reference's own known-answer tests (see
// tests/test_oracle_kat.py): heptagon / octagon literal polygons, half-space
// clip cases, equilibrium plane, 100/28 and 12/4 face counts, generator tet
// counts, benchmark doc-comment triangle counts + areas.
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "ho_bvh.h"
#include "ho_obb.h"
#include "ho_contact.h"
#include "ho_halfspace.h"
#include "ho_traction.h"
#include "ho_cylinder_capsule.h"

using namespace ho;

namespace {

struct VolAsset {
  VolumeMesh mesh;
  std::unique_ptr<std::vector<Aabb>> boxes;
  std::unique_ptr<Bvh> bvh_ref;    // reference-shaped (1 tet / leaf)
  std::unique_ptr<ObbBvh> bvh_obb; // Bvh<Obb, VolumeMesh> (hydroelastic_internal.h:71-100)
  // CompliantMesh's surface side (hydroelastic_internal.h:71-100): surface mesh, tri_to_tet, its Bvh<Obb>, topology.
  std::unique_ptr<TriMesh> surface;
  std::vector<int> tri_to_tet;
  std::unique_ptr<ObbBvh> surface_bvh;
  std::unique_ptr<std::vector<std::array<int, 4>>> neighbors;
};
struct TriAsset {
  TriMesh mesh;
  std::unique_ptr<std::vector<Aabb>> boxes;
  std::unique_ptr<Bvh> bvh_ref;    // reference-shaped (<= 3 tris / leaf)
  std::unique_ptr<Bvh> bvh_leaf1;  // one triangle per leaf
  std::unique_ptr<ObbBvh> bvh_obb; // Bvh<Obb, TriangleSurfaceMesh> (hydroelastic_internal.h:268)
};

std::vector<std::unique_ptr<VolAsset>> g_vols;
std::vector<std::unique_ptr<TriAsset>> g_tris;
std::vector<std::unique_ptr<VolumeField>> g_fields;
std::vector<std::unique_ptr<SurfaceResult>> g_results;
std::vector<std::unique_ptr<std::vector<std::pair<int, int>>>> g_pairlists;
std::mutex g_mu;
thread_local std::string g_err;

int AddVol(VolumeMesh m) {
  std::lock_guard<std::mutex> l(g_mu);
  auto a = std::make_unique<VolAsset>();
  a->mesh = std::move(m);
  g_vols.push_back(std::move(a));
  return static_cast<int>(g_vols.size()) - 1;
}
int AddTri(TriMesh m) {
  std::lock_guard<std::mutex> l(g_mu);
  auto a = std::make_unique<TriAsset>();
  a->mesh = std::move(m);
  g_tris.push_back(std::move(a));
  return static_cast<int>(g_tris.size()) - 1;
}
int AddField(VolumeField f) {
  std::lock_guard<std::mutex> l(g_mu);
  g_fields.push_back(std::make_unique<VolumeField>(std::move(f)));
  return static_cast<int>(g_fields.size()) - 1;
}
int AddResult(SurfaceResult r) {
  std::lock_guard<std::mutex> l(g_mu);
  for (size_t i = 0; i < g_results.size(); ++i)
    if (!g_results[i]) {
      g_results[i] = std::make_unique<SurfaceResult>(std::move(r));
      return static_cast<int>(i);
    }
  g_results.push_back(std::make_unique<SurfaceResult>(std::move(r)));
  return static_cast<int>(g_results.size()) - 1;
}

const std::vector<Aabb>& Boxes(VolAsset& a) {
  if (!a.boxes) a.boxes = std::make_unique<std::vector<Aabb>>(ElementBoxes(a.mesh));
  return *a.boxes;
}
const std::vector<Aabb>& Boxes(TriAsset& a) {
  if (!a.boxes) a.boxes = std::make_unique<std::vector<Aabb>>(ElementBoxes(a.mesh));
  return *a.boxes;
}
const Bvh& RefBvh(VolAsset& a) {
  if (!a.bvh_ref) a.bvh_ref = std::make_unique<Bvh>(MakeBvh(a.mesh));
  return *a.bvh_ref;
}
const Bvh& RefBvh(TriAsset& a) {
  if (!a.bvh_ref) a.bvh_ref = std::make_unique<Bvh>(MakeBvh(a.mesh, 3));
  return *a.bvh_ref;
}
template <class Asset>
const ObbBvh& RefObbBvh(Asset& a) {
  if (!a.bvh_obb) a.bvh_obb = std::make_unique<ObbBvh>(MakeObbBvh(a.mesh));
  return *a.bvh_obb;
}
void EnsureSurfaceSide(VolAsset& a) {
  if (a.surface) return;
  a.surface = std::make_unique<TriMesh>(ConvertVolumeToSurfaceMesh(a.mesh, &a.tri_to_tet));
  a.surface_bvh = std::make_unique<ObbBvh>(MakeObbBvh(*a.surface));
  a.neighbors = std::make_unique<std::vector<std::array<int, 4>>>(VolumeMeshNeighbors(a.mesh));
}
const Bvh& Leaf1Bvh(VolAsset& a) { return RefBvh(a); }
const Bvh& Leaf1Bvh(TriAsset& a) {
  if (!a.bvh_leaf1) a.bvh_leaf1 = std::make_unique<Bvh>(MakeBvh(a.mesh, 1));
  return *a.bvh_leaf1;
}

// mode 0: leaf predicate by brute force; 1: leaf predicate through trees;
// 2: reference-shaped Bvh<Aabb>::Collide (bvh.h:220-262); 3: the reference's own Bvh<Obb>::Collide (ho_obb.h).
template <class A, class B>
std::vector<std::pair<int, int>> Candidates(A& a, B& b, const Xform& X_AB, int mode, int64_t* bv_tests = nullptr) {
  if (mode == 0) return LeafCandidatesBruteForce(Boxes(a), Boxes(b), X_AB);
  if (mode == 1) return LeafCandidatesTree(Leaf1Bvh(a), Leaf1Bvh(b), X_AB);
  std::vector<std::pair<int, int>> out;
  if (mode == 3) CollideObb(RefObbBvh(a), RefObbBvh(b), X_AB, &out, bv_tests);
  else Collide(RefBvh(a), RefBvh(b), X_AB, &out, bv_tests);
  return out;
}

struct SceneGeom {
  int64_t id;
  int type;  // 0 undefined, 1 rigid, 2 compliant (proximity_properties.h:84-88)
  int vol, field, tri;
  bool anchored;
  // Half space (shape_specification.h HalfSpace): no mesh; a compliant one has
  // pressure_scale = hydroelastic_modulus / slab_thickness and a margin
  // (hydroelastic_internal.h SoftHalfSpace).
  bool half_space{false};
  double pressure_scale{0}, margin{0};
  // Geometry with no hydroelastic representation: the box of its shape in its own frame (centre, half widths) when
  // the caller knows one; otherwise it is everywhere as far as the geometry-level broad phase is concerned.
  bool has_box{false};
  Aabb box{};
};
struct Scene {
  // SampleVolumeFieldOnSurface's filter_face_normal_along_field_gradient (mesh_intersection.h): true in
  // ComputeContactSurfaces, false in the reference's mesh_intersection_benchmark.
  bool normal_filter{true};
  std::vector<SceneGeom> geoms;  // kept sorted by id
  std::set<std::pair<int64_t, int64_t>> filtered;
};
std::vector<std::unique_ptr<Scene>> g_scenes;

struct SceneStats {
  int64_t geometry_pairs{0}, bv_tests{0}, element_pairs{0}, faces{0}, surfaces{0};
  double total_area{0};
};

// Error codes mirror hydroelastic_calculator.h:39-57 failures.
enum { kOk = 0, kErrUnsupported = -2, kErrRigidRigid = -3, kErrHalfSpaceHalfSpace = -4,
       kErrCompliantHalfSpaceCompliantMesh = -5 };

// The geometry-level broad phase (FindCollisionCandidates, proximity_engine.cc:867-883) as far as it matters for the
// pairs that produce no surface: the two world boxes overlap, each built the way FCL's CollisionObject::computeAABB
// builds them (FCL is an external dependency of the reference, pinned in tools/workspace/fcl_internal; restated from
// its published behaviour): the local box translated when the rotation is the identity, otherwise the cube about the
// transformed box centre that holds the box's bounding sphere. A half space (z <= 0) has an unbounded local box; only
// under the identity rotation does the translation leave it bounded (in z, from above).
struct WorldBox {
  double mn[3], mx[3];
};
WorldBox WorldBoxOf(const SceneGeom& g, const Xform& X) {
  const double inf = std::numeric_limits<double>::infinity();
  WorldBox w{{-inf, -inf, -inf}, {inf, inf, inf}};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double m[12];
  X.ToArray(m);
  bool identity = true;
  for (int k = 0; k < 9; ++k) identity = identity && std::fabs(m[k] - I[k]) <= 1e-12;
  if (g.half_space) {
    if (identity) w.mx[2] = m[11];
    return w;
  }
  Aabb box;
  if (g.type == 2) box = RefBvh(*g_vols[g.vol]).nodes[0].bv;
  else if (g.type == 1) box = RefBvh(*g_tris[g.tri]).nodes[0].bv;
  else if (g.has_box) box = g.box;
  else return w;
  const double c[3] = {box.center.x, box.center.y, box.center.z};
  const double h[3] = {box.half_width.x, box.half_width.y, box.half_width.z};
  if (identity) {
    for (int k = 0; k < 3; ++k) {
      w.mn[k] = m[9 + k] + (c[k] - h[k]);
      w.mx[k] = m[9 + k] + (c[k] + h[k]);
    }
  } else {
    const Vec3 cw = X * box.center;
    const double r = std::sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    const double cc[3] = {cw.x, cw.y, cw.z};
    for (int k = 0; k < 3; ++k) {
      w.mn[k] = cc[k] - r;
      w.mx[k] = cc[k] + r;
    }
  }
  return w;
}
bool BroadPhaseReports(const SceneGeom& A, const Xform& X_WA, const SceneGeom& B, const Xform& X_WB) {
  const WorldBox a = WorldBoxOf(A, X_WA), b = WorldBoxOf(B, X_WB);
  for (int k = 0; k < 3; ++k)
    if (!(a.mn[k] <= b.mx[k] && b.mn[k] <= a.mx[k])) return false;
  return true;
}

// Restates the engine loop for one scene. `poses` is indexed like scene.geoms. fallback_pairs (optional): the id pairs
// the WithFallback query hands to MaybeMakePointPair (proximity_engine.cc:958-970).
int ComputeScene(const Scene& scene, const double* poses, Representation rep, int mode, bool with_fallback,
                 std::vector<SurfaceResult>* out, SceneStats* stats,
                 std::vector<std::pair<int64_t, int64_t>>* fallback_pairs = nullptr) {
  const int n = static_cast<int>(scene.geoms.size());
  for (int i = 0; i < n; ++i) {
    for (int j = i + 1; j < n; ++j) {
      const SceneGeom& A = scene.geoms[i];  // id_A < id_B (sorted)
      const SceneGeom& B = scene.geoms[j];
      if (A.anchored && B.anchored) continue;  // proximity_engine.cc:873-878
      if (scene.filtered.count({A.id, B.id})) continue;
      const Xform X_WA = Xform::FromArray(poses + 12 * i);
      const Xform X_WB = Xform::FromArray(poses + 12 * j);
      // The combinations MaybeMakeContactSurface rejects (hydroelastic_calculator.cc:77-156) only matter for the pairs
      // the geometry-level broad phase reports: the strict query throws (proximity_engine.cc:1361-1390), the
      // WithFallback query hands the pair to the point-pair narrow phase (:958-970).
      int failure = 0;
      if (A.type == 0 || B.type == 0) failure = kErrUnsupported;
      else if (A.type == 1 && B.type == 1) failure = kErrRigidRigid;
      else if (A.half_space && B.half_space) failure = kErrHalfSpaceHalfSpace;
      else if ((A.half_space || B.half_space) && A.type == 2 && B.type == 2) failure = kErrCompliantHalfSpaceCompliantMesh;
      if (failure) {
        if (!BroadPhaseReports(A, X_WA, B, X_WB)) continue;
        if (!with_fallback) return failure;
        if (fallback_pairs) fallback_pairs->emplace_back(A.id, B.id);
        continue;
      }
      if (stats) ++stats->geometry_pairs;
      SurfaceResult res;
      bool non_null = false;
      if (A.half_space || B.half_space) {
        // hydroelastic_calculator.cc:25-41 (the two supported cases).
        const bool A_is_rigid = A.type == 1;
        const SceneGeom& S = A_is_rigid ? B : A;
        const SceneGeom& R = A_is_rigid ? A : B;
        const Xform& X_WS = A_is_rigid ? X_WB : X_WA;
        const Xform& X_WR = A_is_rigid ? X_WA : X_WB;
        if (S.half_space) {
          TriAsset& tr = *g_tris[R.tri];
          if (stats) stats->element_pairs += tr.mesh.num_elements();
          non_null = ContactSurfaceFromCompliantHalfSpaceRigidMesh(S.id, X_WS, S.pressure_scale, R.id, tr.mesh, X_WR, rep,
                                                                   S.margin, &res);
        } else {
          VolAsset& vs = *g_vols[S.vol];
          if (stats) stats->element_pairs += vs.mesh.num_elements();
          non_null = ContactSurfaceFromCompliantVolumeRigidHalfSpace(S.id, vs.mesh, *g_fields[S.field], X_WS, R.id, X_WR,
                                                                     rep, &res);
        }
      } else if (A.type == 2 && B.type == 2) {
        VolAsset& va = *g_vols[A.vol];
        VolAsset& vb = *g_vols[B.vol];
        const Xform X_MN = InvertAndCompose(X_WA, X_WB);
        const auto cand = Candidates(va, vb, X_MN, mode, stats ? &stats->bv_tests : nullptr);
        if (stats) stats->element_pairs += cand.size();
        non_null = ContactSurfaceFromCompliantVolumes(A.id, va.mesh, *g_fields[A.field], X_WA, B.id, vb.mesh,
                                                      *g_fields[B.field], X_WB, cand, rep, &res);
      } else {
        const bool A_is_rigid = A.type == 1;
        const SceneGeom& S = A_is_rigid ? B : A;
        const SceneGeom& R = A_is_rigid ? A : B;
        const Xform& X_WS = A_is_rigid ? X_WB : X_WA;
        const Xform& X_WR = A_is_rigid ? X_WA : X_WB;
        VolAsset& vs = *g_vols[S.vol];
        TriAsset& tr = *g_tris[R.tri];
        const Xform X_SR = InvertAndCompose(X_WS, X_WR);
        const auto cand = Candidates(vs, tr, X_SR, mode, stats ? &stats->bv_tests : nullptr);
        if (stats) stats->element_pairs += cand.size();
        non_null = ContactSurfaceFromCompliantVolumeRigidSurface(S.id, vs.mesh, *g_fields[S.field], X_WS, R.id,
                                                                 tr.mesh, X_WR, cand, rep, &res, scene.normal_filter);
      }
      if (non_null) {
        if (stats) {
          ++stats->surfaces;
          stats->faces += res.num_faces();
          stats->total_area += res.total_area;
        }
        if (out) out->push_back(std::move(res));
      }
    }
  }
  // Pairs were visited in (id_A, id_B) order == Order() (proximity_engine.cc:157-162).
  return kOk;
}

void CopyVec3(const std::vector<Vec3>& v, double* out) {
  if (!out) return;
  for (size_t i = 0; i < v.size(); ++i) {
    out[3 * i] = v[i].x;
    out[3 * i + 1] = v[i].y;
    out[3 * i + 2] = v[i].z;
  }
}
std::vector<Vec3> ToVec3(int n, const double* p) {
  std::vector<Vec3> v(n);
  for (int i = 0; i < n; ++i) v[i] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
  return v;
}

}  // namespace

#define HO_TRY try {
#define HO_CATCH(ret)                 \
  }                                   \
  catch (const std::exception& e) {   \
    g_err = e.what();                 \
    return ret;                       \
  }

extern "C" {

const char* ho_last_error() { return g_err.c_str(); }

// ---- generators ------------------------------------------------------------
int ho_make_sphere_volume(double r, double hint, int dense) {
  HO_TRY return AddVol(MakeSphereVolumeMesh(r, hint, dense ? Tessellation::kDenseInteriorVertices
                                                          : Tessellation::kSingleInteriorVertex));
  HO_CATCH(-1)
}
int ho_make_ellipsoid_volume(double a, double b, double c, double hint, int dense) {
  HO_TRY return AddVol(MakeEllipsoidVolumeMesh(a, b, c, hint, dense ? Tessellation::kDenseInteriorVertices
                                                                   : Tessellation::kSingleInteriorVertex));
  HO_CATCH(-1)
}
int ho_make_box_volume_ma(double x, double y, double z) {
  HO_TRY return AddVol(MakeBoxVolumeMeshWithMaAndSymmetricTriangles({x, y, z}));
  HO_CATCH(-1)
}
int ho_make_box_volume_grid(double x, double y, double z, double hint) {
  HO_TRY return AddVol(MakeBoxVolumeMesh({x, y, z}, hint));
  HO_CATCH(-1)
}
int ho_make_cylinder_volume_ma(double radius, double length, double hint) {
  HO_TRY return AddVol(MakeCylinderVolumeMeshWithMa(radius, length, hint));
  HO_CATCH(-1)
}
int ho_make_capsule_volume(double radius, double length, double hint) {
  HO_TRY return AddVol(MakeCapsuleVolumeMesh(radius, length, hint));
  HO_CATCH(-1)
}
int ho_make_volume(int nv, const double* verts, int nt, const int* tets) {
  HO_TRY
  std::vector<Tet> t(nt);
  for (int i = 0; i < nt; ++i) t[i] = {tets[4 * i], tets[4 * i + 1], tets[4 * i + 2], tets[4 * i + 3]};
  return AddVol(VolumeMesh(std::move(t), ToVec3(nv, verts)));
  HO_CATCH(-1)
}
int ho_volume_counts(int h, int* nv, int* nt) {
  *nv = g_vols[h]->mesh.num_vertices();
  *nt = g_vols[h]->mesh.num_elements();
  return 0;
}
int ho_volume_get(int h, double* verts, int* tets, double* inward_normals) {
  const VolumeMesh& m = g_vols[h]->mesh;
  CopyVec3(m.verts, verts);
  if (tets)
    for (size_t i = 0; i < m.tets.size(); ++i)
      for (int k = 0; k < 4; ++k) tets[4 * i + k] = m.tets[i][k];
  if (inward_normals)
    for (size_t i = 0; i < m.tets.size(); ++i)
      for (int k = 0; k < 4; ++k) {
        inward_normals[12 * i + 3 * k] = m.inward_normals[i][k].x;
        inward_normals[12 * i + 3 * k + 1] = m.inward_normals[i][k].y;
        inward_normals[12 * i + 3 * k + 2] = m.inward_normals[i][k].z;
      }
  return 0;
}
double ho_volume_total_volume(int h) {
  const VolumeMesh& m = g_vols[h]->mesh;
  double v = 0;
  for (int e = 0; e < m.num_elements(); ++e) v += m.CalcTetVolume(e);
  return v;
}
int ho_volume_to_surface(int h, int* tri_to_tet /* may be null; sized by a prior call */) {
  HO_TRY
  std::vector<int> map;
  TriMesh s = ConvertVolumeToSurfaceMesh(g_vols[h]->mesh, &map);
  if (tri_to_tet) std::memcpy(tri_to_tet, map.data(), map.size() * sizeof(int));
  return AddTri(std::move(s));
  HO_CATCH(-1)
}
int ho_make_tri(int nv, const double* verts, int nt, const int* tris) {
  HO_TRY
  std::vector<Tri> t(nt);
  for (int i = 0; i < nt; ++i) t[i] = {tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]};
  return AddTri(TriMesh(std::move(t), ToVec3(nv, verts)));
  HO_CATCH(-1)
}
int ho_tri_counts(int h, int* nv, int* nt) {
  *nv = static_cast<int>(g_tris[h]->mesh.verts.size());
  *nt = g_tris[h]->mesh.num_elements();
  return 0;
}
int ho_tri_get(int h, double* verts, int* tris, double* normals, double* areas, double* total_area_centroid4) {
  const TriMesh& m = g_tris[h]->mesh;
  CopyVec3(m.verts, verts);
  if (tris)
    for (size_t i = 0; i < m.tris.size(); ++i)
      for (int k = 0; k < 3; ++k) tris[3 * i + k] = m.tris[i][k];
  CopyVec3(m.normals, normals);
  if (areas) std::memcpy(areas, m.areas.data(), m.areas.size() * sizeof(double));
  if (total_area_centroid4) {
    total_area_centroid4[0] = m.total_area;
    total_area_centroid4[1] = m.centroid.x;
    total_area_centroid4[2] = m.centroid.y;
    total_area_centroid4[3] = m.centroid.z;
  }
  return 0;
}

// ---- fields ----------------------------------------------------------------
int ho_sphere_field(int vol, double r, double modulus, double margin) {
  HO_TRY return AddField(MakeSpherePressureField(r, g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_ellipsoid_field(int vol, double a, double b, double c, double modulus, double margin) {
  HO_TRY return AddField(MakeEllipsoidPressureField(a, b, c, g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_box_field(int vol, double x, double y, double z, double modulus, double margin) {
  HO_TRY return AddField(MakeBoxPressureField({x, y, z}, g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_cylinder_field(int vol, double radius, double length, double modulus, double margin) {
  HO_TRY return AddField(MakeCylinderPressureField(radius, length, g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_capsule_field(int vol, double radius, double length, double modulus, double margin) {
  HO_TRY return AddField(MakeCapsulePressureField(radius, length, g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_mesh_field(int vol, double modulus, double margin) {
  HO_TRY return AddField(MakeVolumeMeshPressureField(g_vols[vol]->mesh, modulus, margin));
  HO_CATCH(-1)
}
int ho_make_convex_volume(int tri) {
  HO_TRY return AddVol(MakeConvexVolumeMesh(g_tris[tri]->mesh));
  HO_CATCH(-1)
}
double ho_distance_to_surface(int tri, const double* p) {
  return CalcDistanceToSurfaceMesh(Vec3(p[0], p[1], p[2]), g_tris[tri]->mesh);
}
int ho_make_field(int vol, const double* values) {
  HO_TRY
  const VolumeMesh& m = g_vols[vol]->mesh;
  return AddField(VolumeField(std::vector<double>(values, values + m.num_vertices()), m));
  HO_CATCH(-1)
}
int ho_field_get(int f, double* values, double* grads, double* at_Mo) {
  const VolumeField& fd = *g_fields[f];
  if (values) std::memcpy(values, fd.values.data(), fd.values.size() * sizeof(double));
  if (fd.degenerate) return 1;
  CopyVec3(fd.gradients, grads);
  if (at_Mo) std::memcpy(at_Mo, fd.values_at_Mo.data(), fd.values_at_Mo.size() * sizeof(double));
  return 0;
}

// ---- low-level functions for known-answer tests ----------------------------
void ho_invert_and_compose(const double* X_BA, const double* X_BC, double* X_AC) {
  InvertAndCompose(Xform::FromArray(X_BA), Xform::FromArray(X_BC)).ToArray(X_AC);
}
void ho_compose(const double* X_AB, const double* X_BC, double* X_AC) {
  (Xform::FromArray(X_AB) * Xform::FromArray(X_BC)).ToArray(X_AC);
}
void ho_calc_intersection(const double* A, const double* B, const double* normal, const double* point, double* out) {
  const PosedHalfSpace H(Vec3{normal[0], normal[1], normal[2]}, Vec3{point[0], point[1], point[2]});
  const Vec3 r = CalcIntersection({A[0], A[1], A[2]}, {B[0], B[1], B[2]}, H);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int ho_clip_polygon_by_half_space(int n, const double* verts, const double* normal, const double* point, double* out) {
  const PosedHalfSpace H(Vec3{normal[0], normal[1], normal[2]}, Vec3{point[0], point[1], point[2]});
  std::vector<Vec3> o;
  ClipPolygonByHalfSpace(ToVec3(n, verts), H, &o);
  CopyVec3(o, out);
  return static_cast<int>(o.size());
}
int ho_remove_nearly_duplicate_vertices(int n, const double* verts, double* out) {
  std::vector<Vec3> p = ToVec3(n, verts);
  RemoveNearlyDuplicateVertices(&p);
  CopyVec3(p, out);
  return static_cast<int>(p.size());
}
int ho_clip_triangle_by_tetrahedron(int vol, int tet, int tri_mesh, int tri, const double* X_MN, double* out) {
  HO_TRY
  const auto p = ClipTriangleByTetrahedron(tet, g_vols[vol]->mesh, tri, g_tris[tri_mesh]->mesh, Xform::FromArray(X_MN));
  CopyVec3(p, out);
  return static_cast<int>(p.size());
  HO_CATCH(-1)
}
int ho_is_face_normal_along_pressure_gradient(int field, int tri_mesh, const double* X_MN, int tet, int tri) {
  return IsFaceNormalAlongPressureGradient(*g_fields[field], g_tris[tri_mesh]->mesh, Xform::FromArray(X_MN), tet, tri);
}
int ho_calc_equilibrium_plane(int field0, int e0, int field1, int e1, const double* X_MN, double* nhat3_disp1) {
  Plane pl;
  if (!CalcEquilibriumPlane(e0, *g_fields[field0], e1, *g_fields[field1], Xform::FromArray(X_MN), &pl)) return 0;
  nhat3_disp1[0] = pl.nhat.x; nhat3_disp1[1] = pl.nhat.y; nhat3_disp1[2] = pl.nhat.z; nhat3_disp1[3] = pl.displacement;
  return 1;
}
int ho_intersect_tetrahedra(int vol0, int e0, int vol1, int e1, const double* X_MN, const double* nhat3_disp1,
                            double* out_verts, int* out_faces) {
  HO_TRY
  Plane pl;
  pl.nhat = {nhat3_disp1[0], nhat3_disp1[1], nhat3_disp1[2]};
  pl.displacement = nhat3_disp1[3];
  const auto [poly, faces] = IntersectTetrahedra(e0, g_vols[vol0]->mesh, e1, g_vols[vol1]->mesh, Xform::FromArray(X_MN), pl);
  CopyVec3(poly, out_verts);
  for (size_t i = 0; i < faces.size(); ++i) out_faces[i] = faces[i];
  return static_cast<int>(poly.size());
  HO_CATCH(-1)
}
int ho_is_plane_normal_along_pressure_gradient(const double* nhat, int tet, int field) {
  return IsPlaneNormalAlongPressureGradient({nhat[0], nhat[1], nhat[2]}, tet, *g_fields[field]);
}
void ho_calc_polygon_centroid(int n, const int* polygon, const double* normal, int nv, const double* verts, double* out) {
  const Vec3 c = CalcPolygonCentroid(std::vector<int>(polygon, polygon + n), {normal[0], normal[1], normal[2]}, ToVec3(nv, verts));
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
int ho_boxes_overlap(const double* half_a, const double* half_b, const double* X_AB) {
  const Xform X = Xform::FromArray(X_AB);
  return BoxesOverlap({half_a[0], half_a[1], half_a[2]}, {half_b[0], half_b[1], half_b[2]}, X.R, X.p);
}
// Element boxes (center, half width) of a volume (kind 0) or triangle (kind 1) mesh.
int ho_element_boxes(int kind, int h, double* centers, double* half_widths) {
  const std::vector<Aabb>& b = kind == 0 ? Boxes(*g_vols[h]) : Boxes(*g_tris[h]);
  for (size_t i = 0; i < b.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      centers[3 * i + k] = b[i].center[k];
      half_widths[3 * i + k] = b[i].half_width[k];
    }
  return static_cast<int>(b.size());
}
// Reference-shaped Bvh<Aabb> statistics: node count, leaf count, height.
int ho_bvh_stats(int kind, int h, int* num_nodes, int* num_leaves, int* height) {
  const Bvh& b = kind == 0 ? RefBvh(*g_vols[h]) : RefBvh(*g_tris[h]);
  *num_nodes = static_cast<int>(b.nodes.size());
  int leaves = 0;
  for (const auto& n : b.nodes) leaves += n.is_leaf();
  *num_leaves = leaves;
  std::vector<std::pair<int, int>> st{{0, 0}};
  int hmax = 0;
  while (!st.empty()) {
    auto [i, d] = st.back();
    st.pop_back();
    hmax = std::max(hmax, d);
    if (!b.nodes[i].is_leaf()) {
      st.emplace_back(b.nodes[i].left, d + 1);
      st.emplace_back(b.nodes[i].right, d + 1);
    }
  }
  *height = hmax;
  return 0;
}

// ---- candidates ------------------------------------------------------------
// kindB: 0 = volume mesh B, 1 = triangle mesh B. A is always a volume mesh.
// Returns a pair-list handle; ho_pairs_get copies it out.
int ho_candidates(int volA, int kindB, int hB, const double* X_AB, int mode, int64_t* bv_tests) {
  HO_TRY
  const Xform X = Xform::FromArray(X_AB);
  int64_t tests = 0;
  auto c = kindB == 0 ? Candidates(*g_vols[volA], *g_vols[hB], X, mode, &tests)
                      : Candidates(*g_vols[volA], *g_tris[hB], X, mode, &tests);
  if (bv_tests) *bv_tests = tests;
  std::lock_guard<std::mutex> l(g_mu);
  g_pairlists.push_back(std::make_unique<std::vector<std::pair<int, int>>>(std::move(c)));
  return static_cast<int>(g_pairlists.size()) - 1;
  HO_CATCH(-1)
}
// VolumeMeshTopology (volume_mesh_topology.cc:13-68): out[4 * e + f] = neighbor(e, f).
int ho_volume_neighbors(int vol, int* out) {
  HO_TRY
  EnsureSurfaceSide(*g_vols[vol]);
  const auto& nb = *g_vols[vol]->neighbors;
  for (size_t e = 0; e < nb.size(); ++e)
    for (int f = 0; f < 4; ++f) out[4 * e + f] = nb[e][f];
  return 0;
  HO_CATCH(-1)
}
// The tetrahedron pairs of the surface-BVH + BFS intersector in visit order (field_intersection.cc:254-375); a
// pair-list handle. ho_soft_soft over this list is the surface that overload builds.
int ho_surface_bfs_pairs(int volM, int fieldM, int volN, int fieldN, const double* X_MN) {
  HO_TRY
  VolAsset& M = *g_vols[volM];
  VolAsset& N = *g_vols[volN];
  EnsureSurfaceSide(M);
  EnsureSurfaceSide(N);
  const Xform X = Xform::FromArray(X_MN);
  std::vector<std::pair<int, int>> surface_pairs;
  CollideObb(*M.surface_bvh, *N.surface_bvh, X, &surface_pairs);
  auto c = SurfaceBfsTetPairs(M.mesh, *g_fields[fieldM], M.tri_to_tet, *M.neighbors, N.mesh, *g_fields[fieldN],
                              N.tri_to_tet, *N.neighbors, X, surface_pairs);
  std::lock_guard<std::mutex> l(g_mu);
  g_pairlists.push_back(std::make_unique<std::vector<std::pair<int, int>>>(std::move(c)));
  return static_cast<int>(g_pairlists.size()) - 1;
  HO_CATCH(-1)
}
int ho_pairs_count(int h) { return static_cast<int>(g_pairlists[h]->size()); }
void ho_pairs_get(int h, int* out) {
  const auto& v = *g_pairlists[h];
  for (size_t i = 0; i < v.size(); ++i) {
    out[2 * i] = v[i].first;
    out[2 * i + 1] = v[i].second;
  }
}
void ho_pairs_free(int h) { g_pairlists[h].reset(new std::vector<std::pair<int, int>>()); }

// ---- single geometry pair ----------------------------------------------------
// `pairs`/`npairs`: optional explicit candidate list; when null the list comes
// from Candidates(mode). Returns a result handle, -1 for "null surface",
// -2 on error.
int ho_soft_rigid(int vol, int field, int tri, const double* X_WS, const double* X_WR, int64_t id_S, int64_t id_R,
                  int rep, int mode, const int* pairs, int npairs, int filter_face_normal) {
  HO_TRY
  const Xform XS = Xform::FromArray(X_WS), XR = Xform::FromArray(X_WR);
  std::vector<std::pair<int, int>> cand;
  if (pairs) {
    for (int i = 0; i < npairs; ++i) cand.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
  } else {
    cand = Candidates(*g_vols[vol], *g_tris[tri], InvertAndCompose(XS, XR), mode);
  }
  SurfaceResult res;
  if (!ContactSurfaceFromCompliantVolumeRigidSurface(id_S, g_vols[vol]->mesh, *g_fields[field], XS, id_R,
                                                     g_tris[tri]->mesh, XR, cand, static_cast<Representation>(rep), &res,
                                                     filter_face_normal != 0))
    return -1;
  return AddResult(std::move(res));
  HO_CATCH(-2)
}
int ho_soft_soft(int volM, int fieldM, int volN, int fieldN, const double* X_WM, const double* X_WN, int64_t id_M,
                 int64_t id_N, int rep, int mode, const int* pairs, int npairs) {
  HO_TRY
  const Xform XM = Xform::FromArray(X_WM), XN = Xform::FromArray(X_WN);
  std::vector<std::pair<int, int>> cand;
  if (pairs) {
    for (int i = 0; i < npairs; ++i) cand.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
  } else {
    cand = Candidates(*g_vols[volM], *g_vols[volN], InvertAndCompose(XM, XN), mode);
  }
  SurfaceResult res;
  if (!ContactSurfaceFromCompliantVolumes(id_M, g_vols[volM]->mesh, *g_fields[fieldM], XM, id_N, g_vols[volN]->mesh,
                                          *g_fields[fieldN], XN, cand, static_cast<Representation>(rep), &res))
    return -1;
  return AddResult(std::move(res));
  HO_CATCH(-2)
}

// The half-space special cases of CalcRigidCompliant (hydroelastic_calculator.cc:25-41).
int ho_soft_mesh_rigid_half_space(int vol, int field, const double* X_WS, const double* X_WR, int64_t id_S, int64_t id_R,
                                  int rep) {
  HO_TRY
  SurfaceResult res;
  if (!ContactSurfaceFromCompliantVolumeRigidHalfSpace(id_S, g_vols[vol]->mesh, *g_fields[field], Xform::FromArray(X_WS),
                                                       id_R, Xform::FromArray(X_WR), static_cast<Representation>(rep),
                                                       &res))
    return -1;
  return AddResult(std::move(res));
  HO_CATCH(-2)
}
int ho_soft_half_space_rigid_mesh(int tri, const double* X_WS, double pressure_scale, double margin, const double* X_WR,
                                  int64_t id_S, int64_t id_R, int rep, int filter_normals) {
  HO_TRY
  SurfaceResult res;
  if (!ContactSurfaceFromCompliantHalfSpaceRigidMesh(id_S, Xform::FromArray(X_WS), pressure_scale, id_R,
                                                     g_tris[tri]->mesh, Xform::FromArray(X_WR),
                                                     static_cast<Representation>(rep), margin, &res, filter_normals != 0))
    return -1;
  return AddResult(std::move(res));
  HO_CATCH(-2)
}

// A ContactSurface built from explicit arrays already in World (what the reference's traction tests do,
// hydroelastic_traction_calculator_test.cc:149-187): polygon rep: face_data |n|v..|, one gradient per face;
// triangle rep: face_data = 3 indices per face, grads ignored.
int ho_result_from_arrays(int rep, int nv, const double* verts, const double* pressures, int nfd, const int* face_data,
                          const double* grads, int64_t id_M, int64_t id_N) {
  HO_TRY
  MeshBuilder b(static_cast<Representation>(rep));
  b.vertices_B = ToVec3(nv, verts);
  b.pressures.assign(pressures, pressures + nv);
  if (b.rep == Representation::kPolygon) {
    b.face_data.assign(face_data, face_data + nfd);
    int f = 0;
    for (int i = 0; i < nfd; i += face_data[i] + 1, ++f)
      b.grad_e_per_face.push_back(Vec3{grads[3 * f], grads[3 * f + 1], grads[3 * f + 2]});
  } else {
    for (int i = 0; i + 2 < nfd; i += 3) b.faces.push_back({face_data[i], face_data[i + 1], face_data[i + 2]});
  }
  SurfaceResult res;
  MakeMeshAndFieldInWorld(std::move(b), Xform{}, &res, /*already_in_world=*/true);
  res.id_M = id_M;
  res.id_N = id_N;
  res.elem_first.assign(res.num_faces(), 0);
  res.elem_second.assign(res.num_faces(), 0);
  return AddResult(std::move(res));
  HO_CATCH(-2)
}

// HydroelasticTractionCalculator (hydroelastic_traction_calculator.cc:24-83) on a result handle: A / B are
// the bodies of geometries M / N. V = (w, v). out18 = F_Ac_W, F_Ao_W, F_Bo_W, each (tau, f).
int ho_traction(int h, const double* p_WAo, const double* p_WBo, const double* V_WA, const double* V_WB,
                double dissipation, double mu_coulomb, double vslip_regularizer, double* out18) {
  HO_TRY
  const SurfaceResult& r = *g_results[h];
  TractionData d;
  d.p_WAo = {p_WAo[0], p_WAo[1], p_WAo[2]};
  d.p_WBo = {p_WBo[0], p_WBo[1], p_WBo[2]};
  d.V_WA = {Vec3{V_WA[0], V_WA[1], V_WA[2]}, Vec3{V_WA[3], V_WA[4], V_WA[5]}};
  d.V_WB = {Vec3{V_WB[0], V_WB[1], V_WB[2]}, Vec3{V_WB[3], V_WB[4], V_WB[5]}};
  d.p_WC = r.centroid_W;
  const Spatial F_Ac_W = SpatialForceAtCentroid(r, d, dissipation, mu_coulomb, vslip_regularizer);
  Spatial F_Ao_W, F_Bo_W;
  ShiftToBodyOrigins(d, F_Ac_W, &F_Ao_W, &F_Bo_W);
  const Spatial* all[3] = {&F_Ac_W, &F_Ao_W, &F_Bo_W};
  for (int k = 0; k < 3; ++k) {
    out18[6 * k + 0] = all[k]->rot.x, out18[6 * k + 1] = all[k]->rot.y, out18[6 * k + 2] = all[k]->rot.z;
    out18[6 * k + 3] = all[k]->trans.x, out18[6 * k + 4] = all[k]->trans.y, out18[6 * k + 5] = all[k]->trans.z;
  }
  return 0;
  HO_CATCH(-2)
}
double ho_atan_x_over_x_from_x_squared(double x2) { return CalcAtanXOverXFromXSquared(x2); }
double ho_combined_dissipation(double kA, double kB, double dA, double dB) {
  return CombinedHuntCrossleyDissipation(kA, kB, dA, dB);
}
double ho_combined_friction(double mu_A, double mu_B) { return CombinedDynamicFriction(mu_A, mu_B); }

// info[0..7] = rep, nfaces, nverts, nfacedata (poly) or 3*ntris, has_grad_eM, has_grad_eN, -, -
int ho_result_info(int h, int64_t* ids2, int* info8) {
  const SurfaceResult& r = *g_results[h];
  ids2[0] = r.id_M;
  ids2[1] = r.id_N;
  info8[0] = static_cast<int>(r.rep);
  info8[1] = r.num_faces();
  info8[2] = static_cast<int>(r.vertices_W.size());
  info8[3] = r.rep == Representation::kPolygon ? static_cast<int>(r.face_data.size()) : 3 * static_cast<int>(r.tris.size());
  info8[4] = r.has_grad_eM;
  info8[5] = r.has_grad_eN;
  info8[6] = info8[7] = 0;
  return 0;
}
int ho_result_get(int h, double* verts, double* pressures, int* face_data, double* areas, double* normals,
                  double* centroids, double* e_centroid, double* field_grad, double* field_at_Wo, double* grad_eM,
                  double* grad_eN, double* total_area_centroid4, int* elem_first, int* elem_second) {
  const SurfaceResult& r = *g_results[h];
  CopyVec3(r.vertices_W, verts);
  if (pressures) std::memcpy(pressures, r.pressures.data(), r.pressures.size() * sizeof(double));
  if (face_data) {
    if (r.rep == Representation::kPolygon) {
      std::memcpy(face_data, r.face_data.data(), r.face_data.size() * sizeof(int));
    } else {
      for (size_t i = 0; i < r.tris.size(); ++i)
        for (int k = 0; k < 3; ++k) face_data[3 * i + k] = r.tris[i][k];
    }
  }
  if (areas) std::memcpy(areas, r.areas.data(), r.areas.size() * sizeof(double));
  CopyVec3(r.normals_W, normals);
  CopyVec3(r.centroids_W, centroids);
  if (e_centroid) std::memcpy(e_centroid, r.e_centroid.data(), r.e_centroid.size() * sizeof(double));
  if (r.rep == Representation::kPolygon) {
    CopyVec3(r.field_grad_W, field_grad);
    if (field_at_Wo) std::memcpy(field_at_Wo, r.field_at_Wo.data(), r.field_at_Wo.size() * sizeof(double));
  }
  if (r.has_grad_eM) CopyVec3(r.grad_eM_W, grad_eM);
  if (r.has_grad_eN) CopyVec3(r.grad_eN_W, grad_eN);
  if (total_area_centroid4) {
    total_area_centroid4[0] = r.total_area;
    total_area_centroid4[1] = r.centroid_W.x;
    total_area_centroid4[2] = r.centroid_W.y;
    total_area_centroid4[3] = r.centroid_W.z;
  }
  if (elem_first) std::memcpy(elem_first, r.elem_first.data(), r.elem_first.size() * sizeof(int));
  if (elem_second) std::memcpy(elem_second, r.elem_second.data(), r.elem_second.size() * sizeof(int));
  return 0;
}
void ho_result_free(int h) {
  std::lock_guard<std::mutex> l(g_mu);
  g_results[h].reset();
}

// ---- scenes ------------------------------------------------------------------
int ho_scene_create() {
  std::lock_guard<std::mutex> l(g_mu);
  g_scenes.push_back(std::make_unique<Scene>());
  return static_cast<int>(g_scenes.size()) - 1;
}
// type: 0 undefined, 1 rigid (tri), 2 compliant (vol + field).
int ho_scene_add_geometry(int scene, int64_t id, int type, int vol, int field, int tri, int anchored) {
  Scene& s = *g_scenes[scene];
  SceneGeom g{id, type, vol, field, tri, anchored != 0};
  auto it = std::lower_bound(s.geoms.begin(), s.geoms.end(), g, [](const SceneGeom& a, const SceneGeom& b) { return a.id < b.id; });
  const int pos = static_cast<int>(it - s.geoms.begin());
  s.geoms.insert(it, g);
  return pos;
}
// A half space: type 1 rigid or 2 compliant (pressure_scale, margin).
int ho_scene_add_half_space(int scene, int64_t id, int type, double pressure_scale, double margin, int anchored) {
  Scene& s = *g_scenes[scene];
  SceneGeom g{id, type, -1, -1, -1, anchored != 0};
  g.half_space = true;
  g.pressure_scale = pressure_scale;
  g.margin = margin;
  auto it = std::lower_bound(s.geoms.begin(), s.geoms.end(), g, [](const SceneGeom& a, const SceneGeom& b) { return a.id < b.id; });
  const int pos = static_cast<int>(it - s.geoms.begin());
  s.geoms.insert(it, g);
  return pos;
}
// A geometry with no hydroelastic representation; box6 (centre, half widths in its frame) may be null.
int ho_scene_add_undefined(int scene, int64_t id, const double* box6, int anchored) {
  Scene& s = *g_scenes[scene];
  SceneGeom g{id, 0, -1, -1, -1, anchored != 0};
  if (box6) {
    g.has_box = true;
    g.box.center = {box6[0], box6[1], box6[2]};
    g.box.half_width = {box6[3], box6[4], box6[5]};
  }
  auto it = std::lower_bound(s.geoms.begin(), s.geoms.end(), g, [](const SceneGeom& a, const SceneGeom& b) { return a.id < b.id; });
  const int pos = static_cast<int>(it - s.geoms.begin());
  s.geoms.insert(it, g);
  return pos;
}
// The id pairs a WithFallback query of the scene hands to the point-pair narrow phase: out_ids [max_pairs*2].
int ho_scene_fallback_pairs(int scene, const double* poses, int64_t* out_ids, int max_pairs) {
  HO_TRY
  std::vector<std::pair<int64_t, int64_t>> pairs;
  const int rc = ComputeScene(*g_scenes[scene], poses, Representation::kPolygon, 1, true, nullptr, nullptr, &pairs);
  if (rc != kOk) return rc;
  for (size_t k = 0; k < pairs.size() && static_cast<int>(k) < max_pairs; ++k) {
    out_ids[2 * k] = pairs[k].first;
    out_ids[2 * k + 1] = pairs[k].second;
  }
  return static_cast<int>(pairs.size());
  HO_CATCH(-100)
}
// DeformableMeshWithBvh::UpdateVertexPositions for a triangle mesh (deformable_mesh_with_bvh.h:86-93): new vertex
// positions, per-triangle normals / areas recomputed, the reference-shaped BVH(s) refitted with their topology kept.
int ho_tri_update_vertices(int h, const double* verts) {
  HO_TRY
  TriAsset& a = *g_tris[h];
  RefBvh(a);
  Leaf1Bvh(a);
  for (size_t i = 0; i < a.mesh.verts.size(); ++i) a.mesh.verts[i] = {verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]};
  a.mesh.ComputePositionDependentQuantities();
  RefitBvh<TriMesh, 3>(a.bvh_ref.get(), a.mesh, a.mesh.tris);
  RefitBvh<TriMesh, 3>(a.bvh_leaf1.get(), a.mesh, a.mesh.tris);
  a.boxes.reset();
  a.bvh_obb.reset();
  return 0;
  HO_CATCH(-100)
}
// Checks the refitted reference-shaped BVH: every node's box contains the vertices of every triangle below it, and the
// root is exactly the box of the mesh (bvh_updater_test.cc). Returns 1 when it holds.
int ho_tri_bvh_is_fitted(int h) {
  TriAsset& a = *g_tris[h];
  const Bvh& b = RefBvh(a);
  std::vector<std::pair<Vec3, Vec3>> ext(b.nodes.size());
  const double inf = std::numeric_limits<double>::infinity();
  for (int n = static_cast<int>(b.nodes.size()) - 1; n >= 0; --n) {   // children come after their parent
    Vec3 lo{inf, inf, inf}, up{-inf, -inf, -inf};
    auto grow = [&](const Vec3& p) {
      lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
      up = {std::max(up.x, p.x), std::max(up.y, p.y), std::max(up.z, p.z)};
    };
    if (b.nodes[n].is_leaf()) {
      for (int e = 0; e < b.nodes[n].num_index; ++e)
        for (int v = 0; v < 3; ++v) grow(a.mesh.verts[a.mesh.tris[b.nodes[n].indices[e]][v]]);
    } else {
      grow(ext[b.nodes[n].left].first);
      grow(ext[b.nodes[n].left].second);
      grow(ext[b.nodes[n].right].first);
      grow(ext[b.nodes[n].right].second);
    }
    ext[n] = {lo, up};
    const Aabb& bv = b.nodes[n].bv;
    for (int d = 0; d < 3; ++d) {
      const double tol = 4e-16 * (std::fabs(lo[d]) + std::fabs(up[d]) + 1e-300);
      if (bv.center[d] - bv.half_width[d] > lo[d] + tol || bv.center[d] + bv.half_width[d] < up[d] - tol) return 0;
      if (n == 0 && (std::fabs(bv.center[d] - (lo[d] + up[d]) / 2) > tol || std::fabs(bv.half_width[d] - (up[d] - lo[d]) / 2) > tol))
        return 0;
    }
  }
  return 1;
}
void ho_scene_set_normal_filter(int scene, int enabled) { g_scenes[scene]->normal_filter = enabled != 0; }
int ho_scene_num_geometries(int scene) { return static_cast<int>(g_scenes[scene]->geoms.size()); }
void ho_scene_geometry_ids(int scene, int64_t* ids) {
  const Scene& s = *g_scenes[scene];
  for (size_t i = 0; i < s.geoms.size(); ++i) ids[i] = s.geoms[i].id;
}
void ho_scene_exclude_pair(int scene, int64_t a, int64_t b) {
  if (a > b) std::swap(a, b);
  g_scenes[scene]->filtered.insert({a, b});
}
// poses: num_geometries x 12 doubles, ordered by ascending geometry id.
// Returns the number of surfaces (handles written to out_handles), or a
// negative error code (-2 unsupported, -3 rigid-rigid).
int ho_scene_compute(int scene, const double* poses, int rep, int mode, int with_fallback, int* out_handles,
                     int max_handles, int64_t* stats5) {
  HO_TRY
  std::vector<SurfaceResult> out;
  SceneStats st;
  const int rc = ComputeScene(*g_scenes[scene], poses, static_cast<Representation>(rep), mode, with_fallback != 0, &out, &st);
  if (rc != kOk) return rc;
  if (stats5) {
    stats5[0] = st.geometry_pairs;
    stats5[1] = st.bv_tests;
    stats5[2] = st.element_pairs;
    stats5[3] = st.faces;
    stats5[4] = st.surfaces;
  }
  int n = 0;
  for (auto& r : out) {
    if (n >= max_handles) break;
    out_handles[n++] = AddResult(std::move(r));
  }
  return static_cast<int>(out.size());
  HO_CATCH(-100)
}

// CPU-baseline timing: runs `nscenes` instances of the scene (poses:
// nscenes x G x 12) over `nthreads` std::threads, results discarded after
// construction (same work as the reference: poses in -> surfaces built).
// stats6 = geometry pairs, bv tests, element pairs, faces, surfaces, (double bits) total area.
int ho_scene_batch_time(int scene, int nscenes, const double* poses, int rep, int mode, int nthreads, double* seconds,
                        int64_t* stats5, double* total_area) {
  HO_TRY
  const Scene& sc = *g_scenes[scene];
  const int G = static_cast<int>(sc.geoms.size());
  // Warm the lazily built per-asset structures outside the timed region.
  for (const SceneGeom& g : sc.geoms) {
    if (g.half_space) continue;
    if (g.type == 2) { Boxes(*g_vols[g.vol]); RefBvh(*g_vols[g.vol]); if (mode == 3) RefObbBvh(*g_vols[g.vol]); }
    if (g.type == 1) { Boxes(*g_tris[g.tri]); RefBvh(*g_tris[g.tri]); Leaf1Bvh(*g_tris[g.tri]); if (mode == 3) RefObbBvh(*g_tris[g.tri]); }
  }
  std::vector<SceneStats> per_thread(nthreads);
  std::atomic<int> next{0};
  std::atomic<int> err{0};
  const auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    for (;;) {
      const int s = next.fetch_add(1);
      if (s >= nscenes) break;
      std::vector<SurfaceResult> out;
      const int rc = ComputeScene(sc, poses + static_cast<size_t>(s) * G * 12, static_cast<Representation>(rep), mode,
                                  false, &out, &per_thread[tid]);
      if (rc != kOk) err = rc;
    }
  };
  if (nthreads <= 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  *seconds = std::chrono::duration<double>(t1 - t0).count();
  SceneStats tot;
  for (const auto& s : per_thread) {
    tot.geometry_pairs += s.geometry_pairs;
    tot.bv_tests += s.bv_tests;
    tot.element_pairs += s.element_pairs;
    tot.faces += s.faces;
    tot.surfaces += s.surfaces;
    tot.total_area += s.total_area;
  }
  if (stats5) {
    stats5[0] = tot.geometry_pairs;
    stats5[1] = tot.bv_tests;
    stats5[2] = tot.element_pairs;
    stats5[3] = tot.faces;
    stats5[4] = tot.surfaces;
  }
  if (total_area) *total_area = tot.total_area;
  return err.load();
  HO_CATCH(-100)
}

}  // extern "C"
