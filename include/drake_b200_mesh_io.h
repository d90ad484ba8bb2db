// drake_b200_mesh_io.h — the two mesh file formats the reference's hydroelastic shapes are declared with, host-side
// C++17, no dependencies: what geometry::Mesh("x.vtk") / Mesh("x.obj") / Convex("x.obj") hand to the path.
//
//   ReadVtkToVolumeMesh            geometry/proximity/vtk_to_volume_mesh.cc:17-70 (VTK's vtkUnstructuredGridReader there):
//                                  a legacy-format ASCII UNSTRUCTURED_GRID of tetrahedra (cell type 10), scaled per axis,
//                                  windings reversed under a mirroring scale; any other cell type throws the reference's
//                                  message. Both CELLS layouts are read (count-prefixed, and the 5.x OFFSETS /
//                                  CONNECTIVITY one). BINARY files are rejected (the reference reads them through VTK).
//   ReadObjToTriangleSurfaceMesh   geometry/proximity/obj_to_surface_mesh.cc:17-45 over geometry/read_obj.cc:18-110
//                                  (tinyobjloader there, triangulate = true): every `v` line is a vertex (scaled), every
//                                  `f` line a face (1-based or negative indices, v / v/vt / v//vn / v/vt/vn), windings
//                                  reversed under a mirroring scale. Triangulation as tinyobjloader's: a triangle as
//                                  it is, a quad along its SHORTER diagonal (0-2 only when strictly shorter, else 1-3 --
//                                  which is what the reference's quad_cube.obj expectations show,
//                                  obj_to_surface_mesh_test.cc:63-70); polygons of more than four vertices are fanned
//                                  about their first vertex (tinyobjloader ear-clips them: same surface for a convex
//                                  polygon, possibly other triangles).
// The arrays feed VolumeMeshShape / SurfaceMeshShape of drake_b200_proximity_engine.h.
#ifndef DRAKE_B200_MESH_IO_H_
#define DRAKE_B200_MESH_IO_H_

#include <array>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace drake_b200 {
namespace mesh_io {

struct VolumeMeshData {
  std::vector<double> vertices;      // [nv*3]
  std::vector<int32_t> tetrahedra;   // [nt*4]
};
struct SurfaceMeshData {
  std::vector<double> vertices;      // [nv*3]
  std::vector<int32_t> triangles;    // [nt*3]
};

inline std::string ReadFile(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open '" + path + "'");
  std::ostringstream ss;
  ss << in.rdbuf();
  return ss.str();
}
inline bool Mirrors(const std::array<double, 3>& scale) {
  return (((scale[0] < 0) + (scale[1] < 0) + (scale[2] < 0)) % 2) == 1;
}

inline VolumeMeshData ReadVtkToVolumeMeshFromString(const std::string& contents, const std::array<double, 3>& scale = {1, 1, 1},
                                                    const std::string& description = "<memory>") {
  std::istringstream in(contents);
  std::string line, word;
  if (!std::getline(in, line) || line.rfind("# vtk DataFile Version", 0) != 0)
    throw std::runtime_error("ReadVtkToVolumeMesh(): '" + description + "' is not a legacy VTK file");
  std::getline(in, line);   // title
  in >> word;
  if (word != "ASCII") throw std::runtime_error("ReadVtkToVolumeMesh(): only ASCII legacy VTK files are read here ('" + description + "' is " + word + ")");
  in >> word;
  std::string dataset;
  in >> dataset;
  if (word != "DATASET" || dataset != "UNSTRUCTURED_GRID")
    throw std::runtime_error("ReadVtkToVolumeMesh(): '" + description + "' does not hold an UNSTRUCTURED_GRID");
  VolumeMeshData mesh;
  std::vector<std::vector<int64_t>> cells;
  std::vector<int> types;
  bool have_points = false, have_cells = false, have_types = false;
  while (!(have_points && have_cells && have_types) && (in >> word)) {
    if (word == "POINTS") {
      long long n = 0;
      std::string type;
      in >> n >> type;
      mesh.vertices.resize(3 * static_cast<size_t>(n));
      for (size_t k = 0; k < mesh.vertices.size(); ++k) {
        if (!(in >> mesh.vertices[k])) throw std::runtime_error("ReadVtkToVolumeMesh(): truncated POINTS in '" + description + "'");
        mesh.vertices[k] *= scale[k % 3];
      }
      have_points = true;
    } else if (word == "CELLS") {
      long long n = 0, size = 0;
      in >> n >> size;
      std::streampos after_header = in.tellg();
      in >> word;
      if (word == "OFFSETS") {   // the 5.x layout: n offsets (n - 1 cells), then `size` connectivity entries
        std::string type;
        in >> type;
        std::vector<int64_t> offsets(static_cast<size_t>(n));
        for (auto& o : offsets) in >> o;
        in >> word >> type;
        if (word != "CONNECTIVITY") throw std::runtime_error("ReadVtkToVolumeMesh(): CELLS without CONNECTIVITY in '" + description + "'");
        std::vector<int64_t> conn(static_cast<size_t>(size));
        for (auto& c : conn) in >> c;
        for (size_t c = 0; c + 1 < offsets.size(); ++c) cells.emplace_back(conn.begin() + offsets[c], conn.begin() + offsets[c + 1]);
      } else {
        in.clear();
        in.seekg(after_header);
        for (long long c = 0; c < n; ++c) {
          long long count = 0;
          in >> count;
          std::vector<int64_t> ids(static_cast<size_t>(count));
          for (auto& v : ids) in >> v;
          cells.push_back(std::move(ids));
        }
      }
      if (!in) throw std::runtime_error("ReadVtkToVolumeMesh(): truncated CELLS in '" + description + "'");
      have_cells = true;
    } else if (word == "CELL_TYPES") {
      long long n = 0;
      in >> n;
      types.resize(static_cast<size_t>(n));
      for (auto& t : types) in >> t;
      if (!in) throw std::runtime_error("ReadVtkToVolumeMesh(): truncated CELL_TYPES in '" + description + "'");
      have_types = true;
    }
  }
  if (!(have_points && have_cells && have_types) || types.size() != cells.size())
    throw std::runtime_error("ReadVtkToVolumeMesh(): '" + description + "' lacks POINTS, CELLS or CELL_TYPES");
  const bool reverse_winding = Mirrors(scale);
  const int64_t nv = static_cast<int64_t>(mesh.vertices.size() / 3);
  for (size_t c = 0; c < cells.size(); ++c) {
    if (types[c] != 10 || cells[c].size() != 4)
      throw std::runtime_error("ReadVtkToVolumeMesh(): mesh data should only contain tetrahedra (type id=10). Read cell with "
                               "type id=" + std::to_string(types[c]) + " and number of points " +
                               std::to_string(cells[c].size()) + " in '" + description + "'.");
    std::array<int, 4> local{0, 1, 2, 3};
    if (reverse_winding) std::swap(local[0], local[2]);
    for (int k : local) {
      if (cells[c][k] < 0 || cells[c][k] >= nv) throw std::runtime_error("ReadVtkToVolumeMesh(): vertex index out of range in '" + description + "'");
      mesh.tetrahedra.push_back(static_cast<int32_t>(cells[c][k]));
    }
  }
  return mesh;
}
inline VolumeMeshData ReadVtkToVolumeMesh(const std::string& path, const std::array<double, 3>& scale = {1, 1, 1}) {
  return ReadVtkToVolumeMeshFromString(ReadFile(path), scale, path);
}

inline SurfaceMeshData ReadObjToTriangleSurfaceMeshFromString(const std::string& contents,
                                                              const std::array<double, 3>& scale = {1, 1, 1},
                                                              const std::string& description = "<memory>") {
  SurfaceMeshData mesh;
  std::istringstream in(contents);
  std::string line;
  const bool reverse_winding = Mirrors(scale);
  bool any_face = false;
  auto vertex = [&mesh](int v) { return &mesh.vertices[3 * static_cast<size_t>(v)]; };
  auto squared_distance = [&vertex](int a, int b) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (vertex(a)[k] - vertex(b)[k]) * (vertex(a)[k] - vertex(b)[k]);
    return s;
  };
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tag;
    if (!(ls >> tag)) continue;
    if (tag == "v") {
      double p[3];
      if (!(ls >> p[0] >> p[1] >> p[2])) throw std::runtime_error("Error parsing OBJ '" + description + "': bad vertex line");
      for (int k = 0; k < 3; ++k) mesh.vertices.push_back(p[k] * scale[k]);
    } else if (tag == "f") {
      std::vector<int> face;
      std::string item;
      const int nv = static_cast<int>(mesh.vertices.size() / 3);
      while (ls >> item) {
        const int index = std::stoi(item.substr(0, item.find('/')));
        const int v = index > 0 ? index - 1 : nv + index;   // negative: relative to the vertices read so far
        if (v < 0 || v >= nv) throw std::runtime_error("Error parsing OBJ '" + description + "': vertex index out of range");
        face.push_back(v);
      }
      if (face.size() < 3) throw std::runtime_error("Error parsing OBJ '" + description + "': a face needs three vertices");
      any_face = true;
      std::vector<std::array<int, 3>> tris;
      if (face.size() == 3) {
        tris.push_back({face[0], face[1], face[2]});
      } else if (face.size() == 4) {
        if (squared_distance(face[0], face[2]) < squared_distance(face[1], face[3])) {
          tris.push_back({face[0], face[1], face[2]});
          tris.push_back({face[0], face[2], face[3]});
        } else {
          tris.push_back({face[0], face[1], face[3]});
          tris.push_back({face[1], face[2], face[3]});
        }
      } else {
        for (size_t k = 1; k + 1 < face.size(); ++k) tris.push_back({face[0], face[k], face[k + 1]});
      }
      for (const auto& t : tris) {
        if (reverse_winding) {
          mesh.triangles.insert(mesh.triangles.end(), {t[2], t[1], t[0]});
        } else {
          mesh.triangles.insert(mesh.triangles.end(), {t[0], t[1], t[2]});
        }
      }
    }
  }
  if (!any_face)
    throw std::runtime_error("The OBJ data parsed contains no objects; the data could be corrupt, empty, or not an OBJ file. "
                             "Name: '" + description + "'");
  return mesh;
}
inline SurfaceMeshData ReadObjToTriangleSurfaceMesh(const std::string& path, const std::array<double, 3>& scale = {1, 1, 1}) {
  return ReadObjToTriangleSurfaceMeshFromString(ReadFile(path), scale, path);
}

}  // namespace mesh_io
}  // namespace drake_b200
#endif  // DRAKE_B200_MESH_IO_H_
