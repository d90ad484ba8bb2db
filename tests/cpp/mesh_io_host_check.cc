// Host check of include/drake_b200_mesh_io.h (no GPU, no library):  mesh_io_host_check vtk|obj <file> sx sy sz
// prints `nv ne` then the vertices (17 significant digits) and the elements, or `error: <message>`.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "drake_b200_mesh_io.h"

int main(int argc, char** argv) {
  if (argc != 6) return 2;
  const std::array<double, 3> scale{std::atof(argv[3]), std::atof(argv[4]), std::atof(argv[5])};
  try {
    std::vector<double> v;
    std::vector<int32_t> e;
    int per = 0;
    if (std::string(argv[1]) == "vtk") {
      auto m = drake_b200::mesh_io::ReadVtkToVolumeMesh(argv[2], scale);
      v = m.vertices;
      e = m.tetrahedra;
      per = 4;
    } else {
      auto m = drake_b200::mesh_io::ReadObjToTriangleSurfaceMesh(argv[2], scale);
      v = m.vertices;
      e = m.triangles;
      per = 3;
    }
    std::printf("%zu %zu\n", v.size() / 3, e.size() / per);
    for (size_t k = 0; k < v.size(); k += 3) std::printf("%.17g %.17g %.17g\n", v[k], v[k + 1], v[k + 2]);
    for (size_t k = 0; k < e.size(); k += per) {
      for (int j = 0; j < per; ++j) std::printf("%d ", e[k + j]);
      std::printf("\n");
    }
  } catch (const std::exception& ex) {
    std::printf("error: %s\n", ex.what());
  }
  return 0;
}
