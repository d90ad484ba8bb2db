/* drake_b200_hydro.h — C-ABI of the B200 hydroelastic contact-surface back end.
 *
 * This is the drop-in boundary for the hot path behind
 *   geometry::internal::ProximityEngine<double>::ComputeContactSurfaces{,WithFallback}
 *   (reference: geometry/proximity_engine.h:309-326, proximity_engine.cc:905-977)
 * batched over many independent scenes. Plain pointers and sizes only; every
 * function returns a dbh_status (0 = ok) and no exception crosses the boundary
 * (the reference reports errors by throwing: proximity_engine.cc:1347-1397).
 *
 * Conventions
 *  - RigidTransform = 12 doubles: rotation column-major (9) then translation
 *    (3) — the memory layout of math::RigidTransform<double>
 *    (math/rigid_transform.h:1004-1007, math/rotation_matrix.h:1185).
 *  - Geometry ids are int64 (GeometryId::get_value()).
 *  - All meshes are expressed in their geometry frame; all results in World.
 *  - Representation: 0 = triangles, 1 = polygons
 *    (HydroelasticContactRepresentation, query_results/contact_surface.h:24).
 */
#ifndef DRAKE_B200_HYDRO_H_
#define DRAKE_B200_HYDRO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t dbh_status;
enum {
  DBH_OK = 0,
  DBH_ERR_INVALID_ARGUMENT = 1,
  DBH_ERR_CUDA = 2,               /* a CUDA runtime call failed; see dbh_last_error() */
  DBH_ERR_NO_DEVICE = 3,          /* no usable sm_100 device: there is no CPU fallback */
  DBH_ERR_DEGENERATE_FIELD = 4,   /* MeshFieldLinear kOkOrThrow: a tetrahedron has |volume| <= eps
                                     (volume_mesh.h:430-434) */
  DBH_ERR_CAPACITY = 5,           /* a work list overflowed its device buffer; enlarge with
                                     dbh_batch_set_capacity() and re-run */
  DBH_ERR_RIGID_RIGID = 6,        /* ContactSurfaceResult::kRigidRigid (hydroelastic_calculator.h:39-57)
                                     in ComputeContactSurfaces (not raised by the WithFallback query) */
  DBH_ERR_UNSUPPORTED = 7,        /* ContactSurfaceResult::kUnsupported: geometry without a
                                     hydroelastic representation */
  DBH_ERR_NOT_READY = 8,          /* results requested before a query completed */
  DBH_ERR_PLANE_NORMAL = 9,       /* Plane ctor: |normal| < 1e-10 (plane.cc:16-29) */
  DBH_ERR_POLYGON_SIZE = 10,      /* tet-triangle polygon with more than 7 vertices (DRAKE_DEMAND,
                                     mesh_intersection.cc:261-262), or a tet-tet clip whose intermediate
                                     polygon outgrew 16 vertices (the reference's own buffer holds 8,
                                     field_intersection.cc:98: beyond that its behaviour is undefined) */
  DBH_ERR_HALF_SPACE_HALF_SPACE = 11,  /* ContactSurfaceResult::kHalfSpaceHalfSpace (hydroelastic_calculator.cc:113-118,
                                          145-147), strict query only */
  DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH = 12, /* kCompliantHalfSpaceCompliantMesh (:120-123), strict query only */
  DBH_ERR_REFERENCE_DEMAND = 13,  /* a DRAKE_DEMAND of the reference would have fired: a triangle edge whose end
                                     points do not lie on different sides of the half-space boundary
                                     (mesh_half_space_intersection.cc:43,66) */
  DBH_ERR_HOST = 14               /* a host-side failure inside the library (out of host memory, ...): no C++
                                     exception crosses this boundary */
};

enum { DBH_TRIANGLE = 0, DBH_POLYGON = 1 };
/* HydroelasticType (geometry/proximity_properties.h:84-88). */
enum { DBH_UNDEFINED = 0, DBH_RIGID = 1, DBH_COMPLIANT = 2 };
/* Query flavour. */
enum { DBH_QUERY_STRICT = 0 /* ComputeContactSurfaces */, DBH_QUERY_WITH_FALLBACK = 1 };

typedef struct dbh_context dbh_context;
typedef struct dbh_batch dbh_batch;

/* -- (1) context: one per device (one process per GPU). ---------------------- */
dbh_status dbh_context_create(int32_t device_ordinal, dbh_context** out);
void dbh_context_destroy(dbh_context* ctx);
/* BVH builder used by the registrations that follow. DBH_BVH_GPU_LBVH (default): linear BVH built on the
 * device (Morton codes + radix sort + Karras' parallel construction + bottom-up fit). DBH_BVH_HOST_MEDIAN_SPLIT:
 * top-down median split on the host, the reference's strategy (bvh.cc:18-122). Query results do not depend on
 * the builder: one element per leaf, leaf boxes bit-identical, candidate sets tree-independent. Measured on the
 * B200 (profiles/README.md): the LBVH's spatial splits give 24-31 % fewer box tests per query than the median
 * split and build a 2 M-tetrahedron asset in half the time. */
enum { DBH_BVH_HOST_MEDIAN_SPLIT = 0, DBH_BVH_GPU_LBVH = 1 };
dbh_status dbh_context_set_bvh_builder(dbh_context* ctx, int32_t builder);
/* Soft-soft intersector used by the compliant-mesh registrations and the batches created afterwards.
 * DBH_INTERSECTOR_VOLUME_BVH (default): the two volume hierarchies are descended against each other -- what
 * ComputeContactSurfaceFromCompliantVolumes runs (field_intersection.cc:218-252,500-520, use_surfaces = false).
 * DBH_INTERSECTOR_SURFACE_BFS: the reference's second IntersectFields overload (field_intersection.cc:254-375,
 * volume_mesh_topology.h): tetrahedron pairs seeded where the two boundary surfaces meet, then a breadth-first walk
 * over the face neighbours of every pair that produced a polygon. Its work is proportional to the contact surface,
 * not to the overlap volume (two 2 M-tetrahedron spheres: ~10^5 instead of ~10^7 pairs per scene). Every polygon it
 * reports is one the volume intersector reports (same values); like the reference's overload it can miss some: a body
 * wholly inside another has no seed (field_intersection_test.cc:296-344), and the walk stops at a pair that produces no
 * polygon (DESIGN.md 4e has the measured rates: none on smooth high-resolution bodies, a few faces on a fifth of the
 * surfaces in dense clutter of grid meshes -- in the reference's own algorithm too). Registration additionally
 * builds the face-neighbour table and a tree over the boundary triangles. At most 2^21 tetrahedra per mesh and
 * 2^22 - 2 geometry pairs per batch. */
enum { DBH_INTERSECTOR_VOLUME_BVH = 0, DBH_INTERSECTOR_SURFACE_BFS = 1 };
dbh_status dbh_context_set_soft_soft_intersector(dbh_context* ctx, int32_t intersector);
/* Thread-local message of the last failing call. */
const char* dbh_last_error(void);
/* ABI version of this library (major * 1000 + minor). */
int32_t dbh_abi_version(void);

/* -- (2) mesh assets (replaces hydroelastic::Geometries::MaybeAddGeometry,
 *        hydroelastic_internal.cc:94-108,145-153): uploads the mesh, derives the
 *        per-tetrahedron field gradient / value-at-origin / inward normals
 *        (mesh_field_linear.h:440-509, volume_mesh.cc:91-138) or per-triangle
 *        normals (triangle_surface_mesh.h:452-487), and builds the BVH. -------- */
typedef struct dbh_asset_info {
  int32_t type;          /* DBH_RIGID or DBH_COMPLIANT */
  int32_t num_vertices;
  int32_t num_elements;
  int32_t num_bvh_nodes;
  int32_t bvh_height;
  int32_t is_half_space; /* 1 for dbh_register_half_space assets (no elements, no BVH) */
  int32_t reserved[2];
  double root_center[3];
  double root_half_width[3];
} dbh_asset_info;

/* VolumeMesh<double> + VolumeMeshFieldLinear<double,double>: verts [nv*3],
 * tets [nt*4], pressure [nv]. */
dbh_status dbh_register_compliant_mesh(dbh_context* ctx, int32_t num_vertices, const double* vertices,
                                       int32_t num_tets, const int32_t* tets, const double* pressure,
                                       int32_t* asset_out);
/* TriangleSurfaceMesh<double>: verts [nv*3], tris [nt*3]. */
dbh_status dbh_register_rigid_mesh(dbh_context* ctx, int32_t num_vertices, const double* vertices,
                                   int32_t num_tris, const int32_t* tris, int32_t* asset_out);
/* HalfSpace shape (z <= 0 of its geometry frame; shape_specification.h): no mesh. type DBH_RIGID, or
 * DBH_COMPLIANT with pressure_scale = hydroelastic_modulus / slab_thickness and margin
 * (hydroelastic_internal.h SoftHalfSpace; hydroelastic_internal.cc MakeSoftRepresentation(HalfSpace)).
 * Pairs with a half space are the special cases of CalcRigidCompliant (hydroelastic_calculator.cc:25-41):
 * compliant mesh x rigid half space (mesh_plane_intersection.cc:251-342) and compliant half space x
 * rigid mesh (mesh_half_space_intersection.cc:338-418); the other combinations fail the strict query
 * like the reference's (DBH_ERR_HALF_SPACE_HALF_SPACE, DBH_ERR_COMPLIANT_HALF_SPACE_COMPLIANT_MESH,
 * DBH_ERR_RIGID_RIGID). Faces of such surfaces carry -1 as the half space's emitting element, and each
 * polygon keeps its own vertices (the reference welds vertices cut from the same mesh edge). */
dbh_status dbh_register_half_space(dbh_context* ctx, int32_t type, double pressure_scale, double margin,
                                   int32_t* asset_out);
/* Deformable reuse: new vertex positions for registered mesh assets that deform from step to step.
 *  - Triangle-mesh assets: the surface mesh of a deformable body in deformable-rigid contact, which the reference keeps
 *    in a DeformableSurfaceMeshWithBvh and intersects with the rigid body's pressure-field volume mesh through the same
 *    tetrahedron-triangle intersector as rigid-compliant contact (deformable_mesh_intersection.cc:85-110): per-triangle
 *    vertices and normals recomputed (triangle_surface_mesh.h:452-487).
 *  - Volume-mesh assets: the deformable bodies of deformable-deformable contact (DeformableVolumeMeshWithBvh +
 *    DeformableGeometry::CalcSignedDistanceField, deformable_contact_geometries.cc:87-92: the vertex values stay, the
 *    field's gradients and values at the origin are rebuilt on the moved mesh): every per-tetrahedron record derived
 *    again with the expressions of registration. A tetrahedron that became degenerate fails the call with
 *    DBH_ERR_DEGENERATE_FIELD, as the reference's field constructor throws.
 * One call = DeformableMeshWithBvh::UpdateVertexPositions for every listed asset (deformable_mesh_with_bvh.h:86-93): the
 * BVH is refitted bottom-up with its topology kept (BvhUpdater::Update, bvh_updater.h:45-103), all on the device.
 * vertices: host, the assets' vertex arrays one after the other ([num_vertices * 3] each, in the order of `assets`).
 * Synchronous; no query of a batch that uses these assets may be in flight. The next queries see the new meshes. */
dbh_status dbh_context_update_vertices(dbh_context* ctx, int32_t num_assets, const int32_t* assets, const double* vertices);
/* The same with the new positions already on this context's device (a GPU-resident deformable solver): one device
 * buffer, the assets' vertex arrays one after the other like `vertices` above. No vertex crosses PCIe; the host's copy of
 * the positions is refreshed lazily, only if the vertex pool has to be regrown by a later registration. */
dbh_status dbh_context_update_vertices_device(dbh_context* ctx, int32_t num_assets, const int32_t* assets,
                                              const double* device_vertices);
dbh_status dbh_asset_get_info(dbh_context* ctx, int32_t asset, dbh_asset_info* out);
/* The asset's BVH for inspection: boxes [(num_bvh_nodes + 1) * 6] = centre, half width per node
 * (node 0 unused, node 1 root), left [(num_bvh_nodes + 1)] = index of the first of the two adjacent
 * children, or ~element for a leaf. Either may be NULL. */
dbh_status dbh_asset_copy_bvh(dbh_context* ctx, int32_t asset, double* boxes, int32_t* left);
/* The surface side of a compliant mesh registered under DBH_INTERSECTOR_SURFACE_BFS, for inspection:
 * neighbors [num_elements * 4] = VolumeMeshTopology::neighbor(e, f), -1 on the boundary (volume_mesh_topology.h);
 * *num_boundary_triangles = triangles of its surface mesh = leaves of its boundary tree. Either may be NULL.
 * DBH_ERR_INVALID_ARGUMENT when the asset has no surface side. */
dbh_status dbh_asset_copy_topology(dbh_context* ctx, int32_t asset, int32_t* neighbors, int32_t* num_boundary_triangles);
/* Copies the derived per-element data back (for tests): compliant: gradients
 * [nt*3], values_at_Mo [nt], inward_normals [nt*12]; rigid: normals [nt*3].
 * Any pointer may be NULL. */
dbh_status dbh_asset_get_derived(dbh_context* ctx, int32_t asset, double* gradients_or_normals,
                                 double* values_at_Mo, double* inward_normals);

/* -- (2b) primitive-shape mesh + pressure-field generators (host, set-up time).
 *        Same topology, element order and vertex order as the reference's
 *        MakeSphereVolumeMesh (make_sphere_mesh.h:543-598), MakeEllipsoidVolumeMesh
 *        (make_ellipsoid_mesh.h:50-73), MakeBoxVolumeMeshWithMaAndSymmetricTriangles
 *        (make_box_mesh.cc:475-504), MakeBoxVolumeMesh (make_box_mesh.cc:870-890),
 *        ConvertVolumeToSurfaceMesh (volume_to_surface_mesh.cc:152-187) and
 *        Make{Sphere,Ellipsoid,Box,Cylinder,Capsule}PressureField. strategy: 0 = kSingleInteriorVertex,
 *        1 = kDenseInteriorVertices (tessellation_strategy.h). ------------------- */
typedef struct dbh_mesh dbh_mesh;
const char* dbh_mesh_last_error(void);
dbh_status dbh_mesh_make_sphere_volume(double radius, double resolution_hint, int32_t strategy, dbh_mesh** out);
dbh_status dbh_mesh_make_ellipsoid_volume(double a, double b, double c, double resolution_hint, int32_t strategy,
                                          dbh_mesh** out);
dbh_status dbh_mesh_make_box_volume_with_ma(double x, double y, double z, dbh_mesh** out);
dbh_status dbh_mesh_make_box_volume(double x, double y, double z, double resolution_hint, dbh_mesh** out);
/* MakeCylinderVolumeMeshWithMa (make_cylinder_mesh.cc:325-425) and MakeCapsuleVolumeMesh (make_capsule_mesh.cc:13-180),
 * both about the z axis; with dbh_mesh_convert_volume_to_surface they are MakeCylinderSurfaceMesh / MakeCapsuleSurfaceMesh
 * (make_cylinder_mesh.h:678-684, make_capsule_mesh.h:117-123), the rigid representations. */
dbh_status dbh_mesh_make_cylinder_volume_with_ma(double radius, double length, double resolution_hint, dbh_mesh** out);
dbh_status dbh_mesh_make_capsule_volume(double radius, double length, double resolution_hint, dbh_mesh** out);
dbh_status dbh_mesh_convert_volume_to_surface(const dbh_mesh* volume, dbh_mesh** out);
dbh_status dbh_mesh_set_sphere_pressure(dbh_mesh* mesh, double radius, double hydroelastic_modulus, double margin);
dbh_status dbh_mesh_set_ellipsoid_pressure(dbh_mesh* mesh, double a, double b, double c, double hydroelastic_modulus,
                                           double margin);
dbh_status dbh_mesh_set_box_pressure(dbh_mesh* mesh, double x, double y, double z, double hydroelastic_modulus,
                                     double margin);
/* A caller's own mesh as a dbh_mesh: vertices [nv*3], elements [ne*vertices_per_element] (4: tetrahedra, 3: triangles).
 * With the next two calls this is the reference's compliant Convex / Mesh: MakeConvexVolumeMesh (make_convex_mesh.cc:
 * 17-60: the closed surface's vertices + the centroid of the enclosed volume, one tetrahedron per triangle) and
 * MakeVolumeMeshPressureField (make_mesh_field.cc:45-96: distance of every interior vertex to the boundary, scaled so that
 * the deepest one carries the modulus; throws the reference's error for a mesh without interior vertices). */
dbh_status dbh_mesh_from_arrays(int32_t num_vertices, const double* vertices, int32_t num_elements, const int32_t* elements,
                                int32_t vertices_per_element, dbh_mesh** out);
dbh_status dbh_mesh_make_convex_volume(const dbh_mesh* closed_surface, dbh_mesh** out);
dbh_status dbh_mesh_set_volume_mesh_pressure(dbh_mesh* mesh, double hydroelastic_modulus, double margin);
/* MakeCylinderPressureField (make_cylinder_field.cc:18-75), MakeCapsulePressureField (make_capsule_field.h:55-90). */
dbh_status dbh_mesh_set_cylinder_pressure(dbh_mesh* mesh, double radius, double length, double hydroelastic_modulus,
                                          double margin);
dbh_status dbh_mesh_set_capsule_pressure(dbh_mesh* mesh, double radius, double length, double hydroelastic_modulus,
                                         double margin);
dbh_status dbh_mesh_get_sizes(const dbh_mesh* mesh, int32_t* num_vertices, int32_t* num_elements,
                              int32_t* vertices_per_element, int32_t* has_pressure);
/* vertices [nv*3], elements [ne*vertices_per_element], pressure [nv]; any may be NULL. */
dbh_status dbh_mesh_copy(const dbh_mesh* mesh, double* vertices, int32_t* elements, double* pressure);
void dbh_mesh_destroy(dbh_mesh* mesh);
/* Registers a generated mesh: volume mesh with pressure -> compliant asset, surface mesh -> rigid asset. */
dbh_status dbh_register_mesh(dbh_context* ctx, const dbh_mesh* mesh, int32_t* asset_out);

/* -- (3) scene batch (the SceneGraph / ProximityEngine registration state for
 *        many independent scenes: AddDynamicGeometry / AddAnchoredGeometry /
 *        collision_filter(), proximity_engine.h:87,122-133). ------------------- */
/* scene_offsets [num_scenes+1] is a CSR over the flat geometry arrays (length
 * G = scene_offsets[num_scenes]). geom_asset: asset handle or -1 for a geometry
 * with no hydroelastic representation (DBH_UNDEFINED). geom_anchored: 0/1.
 * excluded_pairs [num_excluded*2]: flat geometry indices filtered by the
 * CollisionFilter (collision_filter.h:158). */
dbh_status dbh_batch_create(dbh_context* ctx, int32_t num_scenes, const int32_t* scene_offsets,
                            const int32_t* geom_asset, const int64_t* geom_id, const uint8_t* geom_anchored,
                            int32_t num_excluded, const int32_t* excluded_pairs, dbh_batch** out);
void dbh_batch_destroy(dbh_batch* batch);
/* Device work-list capacities (entries). 0 keeps the current value. */
dbh_status dbh_batch_set_capacity(dbh_batch* batch, int64_t max_bv_pairs_per_level, int64_t max_element_pairs,
                                  int64_t max_faces);

/* After dbh_batch_sync() returned DBH_ERR_CAPACITY: enlarges exactly the lists that overflowed, from
 * the device counters of the failed query (they keep counting past a full list). Re-run the query;
 * repeat while it still reports DBH_ERR_CAPACITY (later stages only run once earlier ones fit). */
dbh_status dbh_batch_grow_capacity(dbh_batch* batch);
/* A query walks the batch in `num_chunks` contiguous chunks of scenes that reuse the same device
 * work lists (BV-pair frontier, element-pair lists, polygon staging), so their footprint does not
 * grow with the batch: results are identical for any chunking. dbh_batch_grow_capacity raises the
 * chunk count by itself when the lists would outgrow the memory budget (35 % of device memory, or
 * DBH_WORK_BUDGET_MB). Default 1. */
dbh_status dbh_batch_set_chunks(dbh_batch* batch, int32_t num_chunks);
int32_t dbh_batch_get_chunks(dbh_batch* batch);
/* Number of kernels the last dbh_batch_query launched. */
int32_t dbh_batch_last_launch_count(dbh_batch* batch);

typedef struct dbh_batch_info {
  int32_t num_scenes;
  int32_t num_geometries;
  int32_t num_geometry_pairs;      /* pairs that survive the filter / anchored rule (static) */
  int32_t num_soft_rigid_pairs;
  int32_t num_soft_soft_pairs;
  int32_t num_rigid_rigid_pairs;
  int32_t num_unsupported_pairs;
  int32_t max_traversal_levels;
  int32_t num_half_space_pairs;    /* compliant mesh x rigid half space + compliant half space x rigid mesh */
  int32_t num_rejected_half_space_pairs; /* combinations with a half space that the reference rejects */
} dbh_batch_info;
dbh_status dbh_batch_get_info(dbh_batch* batch, dbh_batch_info* out);

/* -- (4) poses (UpdateWorldPoses, proximity_engine.h:226-227): X_WG for every
 *        geometry, [G*12]. The host variant copies H2D on the batch stream;
 *        the device variant adopts a caller-owned device pointer (no copy). --- */
dbh_status dbh_batch_set_poses_host(dbh_batch* batch, const double* X_WGs);
dbh_status dbh_batch_set_poses_device(dbh_batch* batch, const double* X_WGs_device);

/* -- (5) the query: ComputeContactSurfaces(representation, X_WGs) for every
 *        scene of the batch. Asynchronous on the batch stream;
 *        dbh_batch_sync() waits and reports device-side failures. per_scene_stats_device
 *        (optional, may be NULL): caller-owned device buffer [num_scenes * 6]
 *        int64 = {geometry pairs tested, BV-pair tests, element pairs, faces,
 *        surfaces, reserved}, written by the pack kernel (the NCCL-gathered
 *        payload of the multi-GPU path). -------------------------------------- */
dbh_status dbh_batch_query(dbh_batch* batch, int32_t representation, int32_t query_kind,
                           int64_t* per_scene_stats_device);
dbh_status dbh_batch_sync(dbh_batch* batch);
/* The CUDA stream (cudaStream_t) the batch launches on. */
void* dbh_batch_stream(dbh_batch* batch);

/* -- (5b) a device-side consumer of the surfaces: one spatial force per contact surface, so
 *        that 18 doubles per surface leave the GPU instead of the whole mesh. Replaces
 *        HydroelasticTractionCalculator::ComputeSpatialForcesAtCentroidFromHydroelasticModel and
 *        ShiftSpatialForcesAtCentroidToBodyOrigins (multibody/plant/hydroelastic_traction_calculator.cc:24-83),
 *        fed by HydroelasticEngine::CalcCombinedDissipation (multibody/hydroelastics/hydroelastic_engine.cc:83-104) and
 *        CalcContactFrictionFromSurfaceProperties (coulomb_friction.h:122-139). Body A is the body of geometry
 *        M (id_M), body B the body of geometry N. Data::add_surface_velocity_bias (conveyor-belt surface speeds,
 *        hydroelastic_traction_calculator.h:80-86) is not supported: surfaces move with their bodies. Parity-green on the B200 (tests/test_zz_gpu_wrench.py); timed
 *        as bench.py's e2e.spatial_forces. ---------------------------------- */
/* p_WBo [G*3]: origin of each geometry's body frame in World (NULL: the geometry frame's origin);
 * V_WB [G*6]: spatial velocity (w, v) of that body at its origin, in World (NULL: at rest). Host pointers. */
dbh_status dbh_batch_set_body_kinematics(dbh_batch* batch, const double* p_WBo, const double* V_WB);
/* Per geometry [G]: hydroelastic modulus (+infinity for a rigid geometry), Hunt-Crossley dissipation (s/m),
 * dynamic Coulomb friction. */
dbh_status dbh_batch_set_contact_properties(dbh_batch* batch, const double* hydroelastic_modulus,
                                            const double* hunt_crossley_dissipation, const double* dynamic_friction);
/* Enqueued on the batch stream after dbh_batch_query (no host sync in between is needed). vslip_regularizer:
 * HydroelasticTractionCalculator's constructor argument (default 1e-6 there). */
dbh_status dbh_batch_compute_spatial_forces(dbh_batch* batch, double vslip_regularizer);
/* Waits, then copies [num_surfaces*6] each, (tau, f) = SpatialForce's layout, surfaces in result order:
 * F_Ac_W about the surface centroid, F_Ao_W on body A at Ao, F_Bo_W on body B at Bo. Any may be NULL. */
dbh_status dbh_batch_copy_spatial_forces(dbh_batch* batch, double* F_Ac_W, double* F_Ao_W, double* F_Bo_W);
/* The same without waiting and without splitting: after dbh_batch_sync, enqueues one copy of [num_surfaces*18] =
 * per surface F_Ac_W, F_Ao_W, F_Bo_W (6 doubles each) into out18 (pinned host memory; may be NULL to get the
 * count); another dbh_batch_sync completes it. */
dbh_status dbh_batch_copy_spatial_forces_async(dbh_batch* batch, double* out18, int64_t* num_surfaces);

/* -- (5c) a second device-side consumer: the geometry-only half of
 *        DiscreteUpdateManager::AppendDiscreteContactPairsForHydroelasticContact (multibody/plant/
 *        discrete_update_manager.cc:896-1036), the per-face data the discrete contact solvers (SAP, TAMSI, ...) read
 *        from a ContactSurface with first-order quadrature: one record of 9 doubles per face, in face order (face f
 *        of surface s is record surf_face_offset[s] + f):
 *          p_WC[3] (face centroid), nhat_BA_W[3] (face normal), phi0 = -p0 / g, fn0 = Ae p0, stiffness k = Ae g,
 *        with g = 1 / (1 / gM + 1 / gN), gM = grad_eM . nhat (infinity for a rigid M), gN = -grad_eN . nhat.
 *        stiffness == 0 marks a face the reference skips (Ae <= 1e-14, or gM or gN below 1e-14). Jacobians,
 *        vn0, friction and dissipation need the plant's kinematic tree and stay on the host. Enqueued on the batch
 *        stream after dbh_batch_query; copy after dbh_batch_sync (asynchronous: complete with another
 *        dbh_batch_sync). 72 B per face leave the GPU instead of the complete mesh. ------------------------- */
dbh_status dbh_batch_compute_discrete_pairs(dbh_batch* batch);
dbh_status dbh_batch_copy_discrete_pairs_async(dbh_batch* batch, double* out /* [num_faces*9], may be NULL */,
                                               int64_t* num_faces);

/* -- (5d) a third device-side consumer: the per-contact-point data of DEFORMABLE contact
 *        (geometry/query_results/deformable_contact.{h,cc}). The reference runs deformable pairs through the same
 *        intersectors as hydroelastic pairs:
 *          deformable x rigid       AddDeformableRigidContactSurface, geometry/proximity/deformable_mesh_intersection.cc
 *            :85-183: the volume mesh + pressure field are the RIGID geometry's (dbh_register_compliant_mesh), the
 *            triangle mesh is the deformable body's surface (dbh_register_rigid_mesh);
 *          deformable x deformable  AddDeformableDeformableContactSurface, deformable_field_intersection.cc:46-124: two
 *            volume meshes carrying their signed-distance fields (dbh_register_compliant_mesh with the signed distances
 *            as vertex values). The reference passes the geometry with the LARGER id as the intersector's first field
 *            (:53, :120-124); this library orders a pair by id, smaller first: hand the batch ids in reversed order
 *            (e.g. INT64_MAX - id) for such pairs and map them back.
 *        Deforming meshes move every step: dbh_context_update_vertices, the device-side BvhUpdater.
 *        One record of 24 doubles per face of the last query, in face order:
 *          p_WC[3]   contact point = polygon centroid
 *          nhat_W[3] deformable x rigid: from the rigid body into the deformable one (-face normal of the contact mesh as
 *                    the intersector builds it, whatever the order of the two ids); deformable x deformable: the contact
 *                    mesh's face normal (deformable_contact.cc:58, :96)
 *          R_WC[9]   RotationMatrix::MakeFromOneUnitVector(-nhat_W, z) (math/rotation_matrix.cc:133-242), column-major
 *          deformable x rigid:
 *            pressure     the rigid pressure field at the centroid
 *            grad_W[3]    "pressure_gradient_W" as the reference computes it (:172-174): X_WR * grad_R with the POINT
 *                         transform, i.e. the rotated gradient PLUS p_WR (reproduced, not corrected)
 *            barycentric[3] of the centroid in the deformable triangle (TriangleSurfaceMesh::CalcBarycentric), 0, 0
 *          deformable x deformable:
 *            signed distance at the centroid (the contact field's value there)
 *            barycentric[4] of the centroid in M's tetrahedron, barycentric[4] in N's (VolumeMesh::CalcBarycentric)
 *        (barycentric coordinates through normal equations / Cramer's rule: equal to the reference's QR / LU solves to
 *        rounding, not bitwise) plus, per face, the indices of the two elements that made the polygon (tetrahedron of M;
 *        triangle or tetrahedron of N). Faces of half-space surfaces get NaN records. The index bookkeeping that needs
 *        the deformable meshes' connectivity (contact_vertex_indexes, participating vertices, is_element_inverted) is a
 *        host-side gather over the element indices. Polygon representation only. Enqueued after
 *        dbh_batch_query; copy after dbh_batch_sync (asynchronous: complete with another dbh_batch_sync). ------------ */
dbh_status dbh_batch_compute_deformable_contact(dbh_batch* batch);
dbh_status dbh_batch_copy_deformable_contact_async(dbh_batch* batch, double* out /* [num_faces*24], may be NULL */,
                                                   int32_t* element_a /* [num_faces], may be NULL */,
                                                   int32_t* element_b /* [num_faces], may be NULL */,
                                                   int64_t* num_faces);

/* -- (6) sizes of the last completed query. --------------------------------- */
typedef struct dbh_result_sizes {
  int64_t num_surfaces;
  int64_t num_faces;       /* polygons (DBH_POLYGON) or triangles (DBH_TRIANGLE) */
  int64_t num_vertices;
  int64_t num_polygons;    /* contact polygons (== num_faces for DBH_POLYGON) */
  int64_t face_data_ints;  /* polygon rep: length of the concatenated face_data
                              (polygon_surface_mesh.h:141-185); triangle rep: 3 * num_faces */
  int32_t representation;
  int32_t reserved;
} dbh_result_sizes;
dbh_status dbh_batch_result_sizes(dbh_batch* batch, dbh_result_sizes* out);

/* -- (7) copy-out of packed results into caller buffers (host). Everything a
 *        ContactSurface<double> exposes (query_results/contact_surface.h:160-415).
 *        Surfaces are ordered by (scene, id_M, id_N) with id_M < id_N
 *        (proximity_engine.cc:157-162,932; contact_surface.cc:39-41). Faces and
 *        vertices of a surface are contiguous. Any pointer may be NULL. -------- */
typedef struct dbh_result_buffers {
  /* per surface [num_surfaces] */
  int32_t* surf_scene;
  int64_t* surf_id_M;
  int64_t* surf_id_N;
  int64_t* surf_face_offset;    /* [num_surfaces+1] into the face arrays */
  int64_t* surf_vertex_offset;  /* [num_surfaces+1] into the vertex arrays */
  double* surf_total_area;      /* mesh.total_area() */
  double* surf_centroid;        /* [num_surfaces*3] mesh.centroid() in World */
  uint8_t* surf_flags;          /* bit0: HasGradE_M, bit1: HasGradE_N, bit2: M/N swapped */
  /* per vertex [num_vertices] */
  double* vertices;             /* [num_vertices*3] p_WV */
  double* pressures;            /* e_MN at each vertex */
  /* per face [num_faces] */
  int32_t* face_data;           /* polygon: |n|v0..vn-1| per face, indices local to the surface;
                                   triangle: 3 local vertex indices per face */
  double* areas;
  double* normals;              /* [num_faces*3], unit, out of N into M */
  double* centroids;            /* [num_faces*3] */
  double* pressure_at_centroid; /* poly_e_MN().EvaluateCartesian(f, centroid) or
                                   tri_e_MN().Evaluate(f, (1/3,1/3,1/3)) */
  double* grad_eM;              /* [num_faces*3] EvaluateGradE_M_W; zeros where !HasGradE_M */
  double* grad_eN;              /* [num_faces*3] */
  int32_t* elem_M;              /* emitting element of the first-constructed body (tet of S / tet0) */
  int32_t* elem_N;              /* emitting element of the second (triangle of R / tet1) */
  /* per vertex [num_vertices], optional (leave NULL; only batches with half-space pairs have it): where the vertex
     of a half-space surface comes from in the one mesh of its pair, so that the host can weld the vertices the
     reference shares between polygons (cut_edges, mesh_plane_intersection.cc:205-231; mesh_half_space_
     intersection.cc:52-101): i in 0..3 = a copy of local vertex i of the face's emitting element; 16 + 4 a + b = cut
     from the edge between its local vertices a and b; -1 = the centroid vertex of a triangle fan. Undefined for the
     vertices of mesh-mesh surfaces (they are never shared). drake_b200_proximity_engine.h welds with it. */
  int32_t* vertex_origin;
} dbh_result_buffers;
dbh_status dbh_batch_copy_results(dbh_batch* batch, const dbh_result_buffers* buffers);
/* Same copies enqueued on the batch stream without waiting (destinations should be pinned host
 * memory); dbh_batch_sync() completes them. Lets a caller that alternates two batches overlap
 * one batch's copy-out with the other's query. */
dbh_status dbh_batch_copy_results_async(dbh_batch* batch, const dbh_result_buffers* buffers);

/* Device-resident views of the same arrays, in the same layout, for GPU
 * consumers (valid until the next query on this batch). */
typedef struct dbh_device_results {
  const int32_t* surf_pair;        /* index into the static geometry-pair list */
  const int32_t* surf_scene;
  const int64_t* surf_id_M;
  const int64_t* surf_id_N;
  const int64_t* surf_face_offset;
  const int64_t* surf_vertex_offset;
  const double* surf_total_area;
  const double* surf_centroid;
  const uint8_t* surf_flags;
  const double* vertices;
  const double* pressures;
  const int32_t* face_data;
  const double* areas;
  const double* normals;
  const double* centroids;
  const double* pressure_at_centroid;
  const double* grad_eM;
  const double* grad_eN;
  const int32_t* elem_M;
  const int32_t* elem_N;
  const int32_t* face_pair;        /* per face: index into the static geometry-pair list */
  const int32_t* vertex_origin;    /* per vertex, half-space surfaces only (NULL when the batch has none) */
} dbh_device_results;
dbh_status dbh_batch_device_results(dbh_batch* batch, dbh_device_results* out);

/* -- (8) statistics of the last completed query. ---------------------------- */
typedef struct dbh_stats {
  int64_t geometry_pairs_tested;   /* soft-rigid + soft-soft + half-space pairs whose root boxes were tested */
  int64_t geometry_pairs_overlapping; /* root boxes overlap */
  int64_t bv_pair_tests;
  int64_t tet_tri_pairs;           /* candidate element pairs narrow-phased */
  int64_t tet_tet_pairs;
  int64_t polygons;
  int64_t faces;
  int64_t vertices;
  int64_t surfaces;
  int64_t algorithmic_bytes;       /* SURVEY.md §8(d) accounting */
  double total_area;
  int64_t tet_tri_full_clips;      /* candidates that survived the conservative culls and ran */
  int64_t tet_tet_full_clips;      /* the reference's full clipping sequence */
  int64_t leaf_pairs;              /* leaf pairs passed by the conservative descent (>= candidates) */
  int64_t cull_violations;         /* self-check mode only (dbh_batch_set_verify): polygons emitted by
                                      candidates the conservative cull stage had rejected; must be 0 */
  int64_t exceptional_clips;       /* candidates whose intermediate polygon outgrew 8 vertices */
  int64_t leaf_exact_tests;        /* leaf pairs whose box predicate fp32 could not decide (fp64 test ran) */
  int64_t half_space_elements;     /* tetrahedra sliced by a rigid half space + triangles clipped by a compliant
                                      one (pairs whose root box reaches the boundary); 168 B / 108 B of
                                      algorithmic bytes each */
  int64_t bv_pair_tests_tet_tet;   /* of bv_pair_tests: tests between two tetrahedron trees */
  int64_t leaf_pairs_tet_tet;      /* of leaf_pairs: tetrahedron x tetrahedron */
  int64_t discrete_pairs;          /* faces kept by dbh_batch_compute_discrete_pairs (0 if it did not run) */
  int64_t fallback_pairs;          /* WithFallback query: pairs handed to the point-pair narrow phase
                                      (dbh_batch_copy_fallback_pairs) */
  int64_t bfs_pairs;               /* surface intersector only: tetrahedron pairs handed to the exact stage */
  int64_t bfs_waves;               /* and the breadth-first waves that took (summed over the chunks) */
} dbh_stats;
dbh_status dbh_batch_stats(dbh_batch* batch, dbh_stats* out);

/* Candidate element pairs of the last query (for parity tests): triples
 * (geometry pair index, element of first body, element of second body).
 * kind 0 = tet-tri list, 1 = tet-tet list; kind | 2 = every leaf pair the conservative
 * descent passed (a superset). Pass out == NULL to get the count. */
dbh_status dbh_batch_copy_candidates(dbh_batch* batch, int32_t kind, int32_t* out_triples, int64_t* count);
/* Geometry-level broad phase of the pairs that produce no contact surface. The reference calls
 * MaybeMakeContactSurface only for the pairs FindCollisionCandidates() reports (proximity_engine.cc:867-883,
 * 905-934: FCL world boxes overlap), so a far-away rigid-rigid pair, a geometry with no hydroelastic
 * representation, or a rejected half-space combination neither fails ComputeContactSurfaces nor reaches the
 * point-pair fallback. Here such a pair counts when the two world boxes overlap, built the way FCL builds
 * them (CollisionObject::computeAABB: the local box translated under an identity rotation, else the cube around
 * the box's bounding sphere; a half space is bounded only under the identity rotation): a mesh's box is its root
 * box; for geometries registered with asset -1 the caller provides the box of the shape in the geometry frame,
 * boxes [G*6] = centre, half widths (entries of geometries that have an asset are ignored; a non-finite entry
 * means "no bound": the pair always counts, the default). */
dbh_status dbh_batch_set_geometry_bounds(dbh_batch* batch, const double* boxes);
/* After a DBH_QUERY_WITH_FALLBACK query: the pairs the reference hands to penetration_as_point_pair::
 * MaybeMakePointPair (proximity_engine.cc:958-970) -- rigid-rigid, undefined-type and rejected half-space
 * combinations whose world boxes overlap -- as quads {scene, flat geometry index of the smaller id, flat geometry
 * index of the larger id, kind (2 rigid-rigid, 3 unsupported, 6 half space x half space, 7 compliant half space x
 * compliant mesh, 8 rigid half space x rigid)}, in (scene, id_M, id_N) order. The point-pair arithmetic itself is
 * host-side (include/drake_b200_proximity_engine.h): it needs the primitive shapes, which this library never
 * sees. Pass out == NULL to get the count. */
dbh_status dbh_batch_copy_fallback_pairs(dbh_batch* batch, int32_t* out_quads, int64_t* count);
/* Static geometry-pair list: per pair {scene, flat geom index of first body
 * (S or M), flat geom index of second body (R or N), kind (0 soft-rigid, 1
 * soft-soft, 2 rigid-rigid, 3 unsupported)}. out [num_geometry_pairs*4]. */
dbh_status dbh_batch_copy_geometry_pairs(dbh_batch* batch, int32_t* out);

/* Device time of the last query in milliseconds (CUDA events on the batch stream): {pose+root, per-level
 * descent, narrow phase tet-tri (cull + exact clip), narrow phase tet-tet (cull + exact clip + half-space
 * count pass), scan, pack, reduce, total}. Only recorded when enabled (adds event records). */
dbh_status dbh_batch_enable_timing(dbh_batch* batch, int32_t enabled);
/* Self-check mode: the reference's exact clipping sequence runs for EVERY
 * candidate element pair (the conservative cull stage only flags), and
 * dbh_stats::cull_violations counts flagged candidates that emitted a polygon.
 * Results are those of the un-culled pipeline. Slower; for tests. */
dbh_status dbh_batch_set_verify(dbh_batch* batch, int32_t enabled);
/* Rigid-compliant mesh pairs with or without the face-normal filter: SurfaceVolumeIntersector::
 * SampleVolumeFieldOnSurface(..., filter_face_normal_along_field_gradient) (mesh_intersection.h). On (the default)
 * is what ComputeContactSurfaces does; off is how geometry/benchmarking/mesh_intersection_benchmark.cc runs the
 * intersector (its rigid sphere sits inside the compliant ellipsoid, where the filter would reject every face). */
dbh_status dbh_batch_set_normal_filter(dbh_batch* batch, int32_t enabled);
/* Mid phase of SMALL geometry pairs (both BVHs at most 32767 nodes): one warp per pair descends the dual
 * tree on a node-pair stack in shared memory and runs the leaf predicate and the conservative culls on chip
 * (k_small_pairs), instead of one launch per tree level through global frontier lists. mode 0: never; 1: every
 * eligible pair; 2 (default): when a chunk of scenes holds enough eligible pairs to fill the device (a handful
 * of pairs is faster through the per-level pipeline, which spreads one pair over the whole grid). Results
 * do not depend on the mode. */
dbh_status dbh_batch_set_small_pair_path(dbh_batch* batch, int32_t mode);
/* Keep the candidate element-pair lists of the following queries for dbh_batch_copy_candidates (parity
 * tests). The per-level pipeline always has them (they are its work lists); the on-chip path only writes
 * them out when asked. Off by default. */
dbh_status dbh_batch_set_record_candidates(dbh_batch* batch, int32_t enabled);
dbh_status dbh_batch_kernel_times_ms(dbh_batch* batch, float* out8);
/* The same at stage granularity (first n of 11 values): {pose + root test, per-level descent (k_traverse), cull
 * tet-tri (k_filter + k_small_pairs), cull tet-tet, exact clip tet-tri (k_clip), exact clip tet-tet (+ exceptional
 * candidates), half-space set-up + count pass, scan, pack (k_pack + half-space pack pass), surface totals, whole
 * query}. */
dbh_status dbh_batch_stage_times_ms(dbh_batch* batch, float* out, int32_t n);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* DRAKE_B200_HYDRO_H_ */
