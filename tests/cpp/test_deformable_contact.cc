// The deformable-contact mirror (include/drake_b200_deformable_contact.h) through the reference's own fixtures:
//   geometry/proximity/test/deformable_mesh_intersection_test.cc  OnePolygon :106-178, NoContact :33-104,
//                                                                 InvertedElements :255-330
// and, for two deformable bodies (the reference's fixture hand-writes signed distances, which the Geometries-level
// interface derives itself), the invariants AddDeformableDeformableContactSurface guarantees
// (deformable_field_intersection.cc:46-124, deformable_field_intersection_test.cc:114-173): the contact point is the
// polygon centroid, both signed-distance fields interpolate to the stored signed distance there, the barycentric
// coordinates reproduce the point in both tetrahedra, the normal points out of B into A, every vertex of a
// participating tetrahedron participates.
// Exit code 0 = all checks passed; 3 = no CUDA device (the library has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "drake_b200_deformable_contact.h"

using namespace drake_b200::deformable;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d: EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

static bool Near(double a, double b, double tol) { return std::fabs(a - b) <= tol; }

// The coarsest sphere mesh: an octahedron of 8 tetrahedra around the one interior vertex v0.
static void MakeOctahedron(VolumeMeshd* volume, TriangleSurfaceMeshd* surface, std::vector<int>* s2v, std::vector<int>* tri2tet,
                           double dx) {
  const double v[7][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  const int faces[8][3] = {{1, 2, 5}, {2, 3, 5}, {3, 4, 5}, {4, 1, 5}, {2, 1, 6}, {3, 2, 6}, {4, 3, 6}, {1, 4, 6}};
  for (int i = 0; i < 7; ++i) {
    volume->vertices.insert(volume->vertices.end(), {v[i][0] + dx, v[i][1], v[i][2]});
    if (i > 0) surface->vertices.insert(surface->vertices.end(), {v[i][0] + dx, v[i][1], v[i][2]});
    if (i > 0) s2v->push_back(i);
  }
  for (int t = 0; t < 8; ++t) {
    volume->tetrahedra.insert(volume->tetrahedra.end(), {0, faces[t][0], faces[t][1], faces[t][2]});
    surface->triangles.insert(surface->triangles.end(), {faces[t][0] - 1, faces[t][1] - 1, faces[t][2] - 1});
    tri2tet->push_back(t);
  }
}

// MakeBoxVolumeMeshWithMa(Box::MakeCube(2)) + MakeBoxPressureField(..., 1e5) through the library's generators.
static void MakeCube(VolumeMeshd* mesh, std::vector<double>* pressure) {
  dbh_mesh* m = nullptr;
  EXPECT(dbh_mesh_make_box_volume_with_ma(2.0, 2.0, 2.0, &m) == DBH_OK);
  EXPECT(dbh_mesh_set_box_pressure(m, 2.0, 2.0, 2.0, 1e5, 0.0) == DBH_OK);
  int32_t nv, ne, vpe, hp;
  EXPECT(dbh_mesh_get_sizes(m, &nv, &ne, &vpe, &hp) == DBH_OK && vpe == 4 && hp == 1);
  mesh->vertices.resize(3 * nv);
  mesh->tetrahedra.resize(4 * ne);
  pressure->resize(nv);
  EXPECT(dbh_mesh_copy(m, mesh->vertices.data(), mesh->tetrahedra.data(), pressure->data()) == DBH_OK);
  dbh_mesh_destroy(m);
}

int main() {
  try {
    Geometries probe(0);
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    return 3;
  }
  const VolumeMeshd single{{0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 1, 2, 3}};
  // ConvertVolumeToSurfaceMeshWithBoundaryVertices of the single tetrahedron.
  const TriangleSurfaceMeshd single_surface{single.vertices, {2, 1, 0, 1, 3, 0, 3, 2, 0, 1, 2, 3}};

  {  // OnePolygon
    Geometries g(0);
    g.AddDeformableGeometry(7, single, single_surface, {0, 1, 2, 3}, {0, 0, 0, 0});
    g.AddRigidGeometry(3, single, {123.0, 123.0, 123.0, 123.0}, RigidTransformd::FromTranslation(1.0 / 6.0, 1.0 / 6.0, -0.5));
    EXPECT(g.is_deformable(7) && g.is_rigid(3) && !g.is_rigid(7));
    const DeformableContact c = g.ComputeDeformableContact();
    EXPECT(c.contact_surfaces().size() == 1);
    const DeformableContactSurface& s = c.contact_surfaces()[0];
    EXPECT(s.id_A() == 7 && s.id_B() == 3 && !s.is_B_deformable());
    EXPECT(s.num_contact_points() == 1 && s.pressures().size() == 1);
    EXPECT(Near(s.pressures()[0], 123.0, 123.0 * 2.3e-16));
    const Vector3d p = s.contact_points_W()[0];
    EXPECT(Near(p[0], 1.0 / 3, 1e-14) && Near(p[1], 1.0 / 3, 1e-14) && Near(p[2], 0.0, 1e-14));
    EXPECT((s.tri_contact_vertex_indexes_A()[0] == Vector3i{2, 1, 0}));
    const Vector3d b = s.tri_barycentric_coordinates_A()[0];
    EXPECT(Near(b[0], 1.0 / 3, 1e-14) && Near(b[1], 1.0 / 3, 1e-14) && Near(b[2], 1.0 / 3, 1e-14));
    EXPECT(c.contact_participation(7).num_vertices_in_contact() == 3);
    EXPECT(!s.is_element_inverted()[0]);
    // Rigid into deformable: up; the contact mesh's face normal is the opposite; R_WC's z axis is -nhat.
    EXPECT(Near(s.nhats_W()[0][2], 1.0, 1e-15) && Near(s.contact_mesh_W().face_normals[0][2], -1.0, 1e-15));
    EXPECT(Near(s.R_WCs()[0][8], -1.0, 1e-15));
    // Moved out of contact: nothing.
    g.UpdateRigidWorldPose(3, RigidTransformd::FromTranslation(1.0 / 6.0, 1.0 / 6.0, -1.5));
    EXPECT(g.ComputeDeformableContact().contact_surfaces().empty());
  }

  VolumeMeshd cube;
  std::vector<double> cube_pressure;
  MakeCube(&cube, &cube_pressure);

  {  // NoContact (boxes collide, meshes do not) and InvertedElements
    VolumeMeshd oct;
    TriangleSurfaceMeshd oct_surface;
    std::vector<int> s2v, tri2tet;
    MakeOctahedron(&oct, &oct_surface, &s2v, &tri2tet, 0.0);
    Geometries g(0);
    g.AddDeformableGeometry(1, oct, oct_surface, s2v, tri2tet);
    const double t = 4.0 / 3 + 1e-10;
    g.AddRigidGeometry(2, cube, cube_pressure, RigidTransformd::FromTranslation(t, t, t));
    EXPECT(g.ComputeDeformableContact().contact_surfaces().empty());
    g.UpdateRigidWorldPose(2, RigidTransformd::FromTranslation(1.5, 0, 0));
    auto check_inversion = [&](bool expect_inversion) {
      const DeformableContact c = g.ComputeDeformableContact();
      EXPECT(c.contact_surfaces().size() == 1);
      const DeformableContactSurface& s = c.contact_surfaces()[0];
      EXPECT(s.num_contact_points() > 0);
      for (bool inverted : s.is_element_inverted()) EXPECT(inverted == expect_inversion);
      for (double p : s.pressures()) EXPECT(p >= 0.0);
    };
    check_inversion(false);
    std::vector<double> q = oct.vertices;
    q[0] = 2.0;   // the one interior vertex pulled out through the contact region: every participating tet is inverted
    g.UpdateDeformableVertexPositions(1, q, oct_surface.vertices);
    check_inversion(true);
    q[0] = -2.0;  // pulled out the other side: not inverted
    g.UpdateDeformableVertexPositions(1, q, oct_surface.vertices);
    check_inversion(false);
  }

  {  // Two deformable bodies, tips overlapping along x.
    VolumeMeshd a, b;
    TriangleSurfaceMeshd sa, sb;
    std::vector<int> s2v_a, s2v_b, t2t_a, t2t_b;
    MakeOctahedron(&a, &sa, &s2v_a, &t2t_a, 0.0);
    MakeOctahedron(&b, &sb, &s2v_b, &t2t_b, 1.5);
    Geometries g(0);
    g.AddDeformableGeometry(5, a, sa, s2v_a, t2t_a);
    g.AddDeformableGeometry(9, b, sb, s2v_b, t2t_b);
    const DeformableContact c = g.ComputeDeformableContact();
    EXPECT(c.contact_surfaces().size() == 1);
    const DeformableContactSurface& s = c.contact_surfaces()[0];
    EXPECT(s.id_A() == 5 && s.id_B() == 9 && s.is_B_deformable() && s.num_contact_points() > 0);
    const double center_sd = -1.0 / std::sqrt(3.0);   // distance from the centre to the faces |x| + |y| + |z| = 1
    auto sd = [&](int v) { return v == 0 ? center_sd : 0.0; };
    for (int i = 0; i < s.num_contact_points(); ++i) {
      const Vector3d p = s.contact_points_W()[i];
      EXPECT((p == s.contact_mesh_W().element_centroids[i]));
      EXPECT(Near(p[0], 0.75, 1e-12));                      // the equilibrium surface of two equal bodies: the mid plane
      EXPECT(s.nhats_W()[i][0] < -0.999999);                // out of B (at +x) into A
      const Vector4d bA = s.tet_barycentric_coordinates_A()[i], bB = s.barycentric_coordinates_B()[i];
      const Vector4i iA = s.tet_contact_vertex_indexes_A()[i], iB = s.contact_vertex_indexes_B()[i];
      double qa[3] = {0, 0, 0}, qb[3] = {0, 0, 0}, sda = 0, sdb = 0, sum_a = 0, sum_b = 0;
      for (int k = 0; k < 4; ++k) {
        for (int d = 0; d < 3; ++d) {
          qa[d] += bA[k] * a.vertices[3 * iA[k] + d];
          qb[d] += bB[k] * b.vertices[3 * iB[k] + d];
        }
        sda += bA[k] * sd(iA[k]);
        sdb += bB[k] * sd(iB[k]);
        sum_a += bA[k], sum_b += bB[k];
      }
      EXPECT(Near(sum_a, 1.0, 1e-12) && Near(sum_b, 1.0, 1e-12));
      for (int d = 0; d < 3; ++d) EXPECT(Near(qa[d], p[d], 1e-12) && Near(qb[d], p[d], 1e-12));
      EXPECT(Near(sda, s.signed_distances()[i], 1e-12) && Near(sdb, s.signed_distances()[i], 1e-12));
      EXPECT(s.signed_distances()[i] < 0);
      for (int k = 0; k < 4; ++k) EXPECT(c.contact_participation(5).participates(iA[k]) && c.contact_participation(9).participates(iB[k]));
    }
    // A filter that keeps the two apart.
    EXPECT(g.ComputeDeformableContact([](GeometryId, GeometryId) { return false; }).contact_surfaces().empty());
    // Pulled apart.
    std::vector<double> qv = b.vertices, qs = sb.vertices;
    for (size_t k = 0; k < qv.size(); k += 3) qv[k] += 1.0;
    for (size_t k = 0; k < qs.size(); k += 3) qs[k] += 1.0;
    g.UpdateDeformableVertexPositions(9, qv, qs);
    EXPECT(g.ComputeDeformableContact().contact_surfaces().empty());
  }
  std::printf("ok\n");
  return 0;
}
