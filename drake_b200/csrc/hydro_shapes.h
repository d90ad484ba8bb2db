// Host-side (set-up time) mesh and pressure-field generators for the
// primitive shapes the hot path is benchmarked on. They reproduce the
// topology, element order and vertex order of the reference generators so a
// scene registered through this library sees the same meshes Drake would build:
//   MakeSphereVolumeMesh        geometry/proximity/make_sphere_mesh.h:543-598
//   MakeEllipsoidVolumeMesh     make_ellipsoid_mesh.h:50-73
//   MakeBoxVolumeMeshWithMaAndSymmetricTriangles   make_box_mesh.cc:475-504
//   MakeBoxVolumeMesh           make_box_mesh.cc:870-890
//   ConvertVolumeToSurfaceMesh  volume_to_surface_mesh.cc:152-187
//   Make{Sphere,Ellipsoid,Box}PressureField  make_sphere_field.h:38-73,
//                               make_ellipsoid_field.h:41-92, make_box_field.cc:15-56
//   MakeCylinderVolumeMeshWithMa  make_cylinder_mesh.cc:47-425;  MakeCylinderPressureField  make_cylinder_field.cc:18-75
//   MakeCapsuleVolumeMesh       make_capsule_mesh.cc:13-180;     MakeCapsulePressureField   make_capsule_field.h:55-90
#pragma once
#include <cstdint>
#include <vector>

namespace dbh {

// Flat mesh container: `verts` is xyz-interleaved; `elems` holds 4 (tets) or 3
// (triangles) vertex indices per element; `pressure` is per vertex (volume
// meshes with a field only).
struct HostMesh {
  int verts_per_elem{4};
  std::vector<double> verts;
  std::vector<int32_t> elems;
  std::vector<double> pressure;
  int num_vertices() const { return static_cast<int>(verts.size() / 3); }
  int num_elements() const { return static_cast<int>(elems.size() / verts_per_elem); }
};

enum TessellationStrategy { kSingleInteriorVertex = 0, kDenseInteriorVertices = 1 };  // tessellation_strategy.h

int SphereRefinementLevel(double radius, double resolution_hint);
HostMesh MakeSphereVolumeMesh(double radius, double resolution_hint, int strategy);
HostMesh MakeEllipsoidVolumeMesh(double a, double b, double c, double resolution_hint, int strategy);
HostMesh MakeBoxVolumeMeshWithMa(double x, double y, double z);
HostMesh MakeBoxVolumeMesh(double x, double y, double z, double resolution_hint);
HostMesh MakeCylinderVolumeMeshWithMa(double radius, double length, double resolution_hint);
HostMesh MakeCapsuleVolumeMesh(double radius, double length, double resolution_hint);
HostMesh ConvertVolumeToSurfaceMesh(const HostMesh& volume);
HostMesh MakeConvexVolumeMesh(const HostMesh& closed_surface);  // make_convex_mesh.cc:17-60

void SetSpherePressureField(HostMesh* mesh, double radius, double modulus, double margin);
void SetEllipsoidPressureField(HostMesh* mesh, double a, double b, double c, double modulus, double margin);
void SetBoxPressureField(HostMesh* mesh, double x, double y, double z, double modulus, double margin);
void SetCylinderPressureField(HostMesh* mesh, double radius, double length, double modulus, double margin);
void SetCapsulePressureField(HostMesh* mesh, double radius, double length, double modulus, double margin);
void SetVolumeMeshPressureField(HostMesh* mesh, double modulus, double margin);  // make_mesh_field.cc:45-96

}  // namespace dbh
