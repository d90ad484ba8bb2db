"""include/drake_b200_mesh_io.h, the two file formats the reference's hydroelastic Mesh / Convex shapes are declared with
(vtk_to_volume_mesh.cc, obj_to_surface_mesh.cc + read_obj.cc), compiled for the host: the reference's own fixtures
(geometry/test/non_convex_mesh.vtk, quad_cube.obj -- restated here, they are a few lines) must give the arrays the
reference's tests expect (vtk_to_volume_mesh_test.cc, obj_to_surface_mesh_test.cc:26-70)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NON_CONVEX_VTK = """# vtk DataFile Version 2.0
Example of non-convex tetrahedral mesh
ASCII
DATASET UNSTRUCTURED_GRID
POINTS 6 double
0.1 0.1 0.1
0.0 0.0 0.0
1.0 0.0 0.0
0.0 1.0 0.0
0.0 0.0 1.0
0.2 0.2 0.2

CELLS 6 30
4   0 5 2 3
4   0 5 3 4
4   0 5 4 2
4   0 1 4 3
4   0 1 3 2
4   0 1 2 4

CELL_TYPES 6
10
10
10
10
10
10
"""
NON_CONVEX_V = np.array([[0.1, 0.1, 0.1], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.2, 0.2, 0.2]])
NON_CONVEX_T = np.array([[0, 5, 2, 3], [0, 5, 3, 4], [0, 5, 4, 2], [0, 1, 4, 3], [0, 1, 3, 2], [0, 1, 2, 4]])

QUAD_CUBE_OBJ = """mtllib quad_cube.mtl

v  1.000000 -1.000000 -1.000000
v  1.000000 -1.000000  1.000000
v -1.000000 -1.000000  1.000000
v -1.000000 -1.000000 -1.000000
v  1.000000  1.000000 -1.000000
v  1.000000  1.000000  1.000000
v -1.000000  1.000000  1.000000
v -1.000000  1.000000 -1.000000
vn 0.0000 -1.0000 0.0000
vn 0.0000 1.0000 0.0000
vn 1.0000 0.0000 0.0000
vn -0.0000 -0.0000 1.0000
vn -1.0000 -0.0000 -0.0000
vn 0.0000 0.0000 -1.0000
usemtl CubeMaterial
f 1//1 2//1 3//1 4//1
f 5//2 8//2 7//2 6//2
f 1//3 5//3 6//3 2//3
f 2//4 6//4 7//4 3//4
f 3//5 7//5 8//5 4//5
f 5//6 1//6 4//6 8//6
"""
# obj_to_surface_mesh_test.cc:52-70: the vertices in file order, and tinyobjloader's split of every quad along its 1-3 diagonal.
QUAD_CUBE_V = np.array([[1, -1, -1], [1, -1, 1], [-1, -1, 1], [-1, -1, -1], [1, 1, -1], [1, 1, 1], [-1, 1, 1], [-1, 1, -1]], float)
QUAD_CUBE_T = np.array([[0, 1, 3], [1, 2, 3], [4, 7, 5], [7, 6, 5], [0, 4, 1], [4, 5, 1], [1, 5, 2], [5, 6, 2], [2, 6, 3], [6, 7, 3],
                        [4, 0, 7], [0, 3, 7]])


def _exe():
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "mesh_io_host_check")
    src = os.path.join(ROOT, "tests", "cpp", "mesh_io_host_check.cc")
    hdr = os.path.join(ROOT, "include", "drake_b200_mesh_io.h")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe],
                       check=True)
    return exe


def _read(tmp_path, kind, text, scale=(1, 1, 1), name=None):
    path = tmp_path / (name or ("mesh." + kind))
    path.write_text(text)
    out = subprocess.run([_exe(), kind, str(path)] + [repr(float(s)) for s in scale], capture_output=True, text=True, check=True).stdout
    if out.startswith("error:"):
        raise RuntimeError(out[len("error: "):].strip())
    rows = out.strip().split("\n")
    nv, ne = (int(x) for x in rows[0].split())
    V = np.array([[float(x) for x in r.split()] for r in rows[1:1 + nv]])
    E = np.array([[int(x) for x in r.split()] for r in rows[1 + nv:1 + nv + ne]])
    return V, E


def _tet_volumes(V, T):
    a, b, c, d = (V[T[:, k]] for k in range(4))
    return np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0


def test_vtk_reference_fixture_scale_and_mirroring(tmp_path):
    V, T = _read(tmp_path, "vtk", NON_CONVEX_VTK)
    np.testing.assert_array_equal(V, NON_CONVEX_V)
    np.testing.assert_array_equal(T, NON_CONVEX_T)
    assert np.all(_tet_volumes(V, T) > 0)
    V2, T2 = _read(tmp_path, "vtk", NON_CONVEX_VTK, (2, 3, 4))
    np.testing.assert_array_equal(V2, NON_CONVEX_V * [2, 3, 4])
    np.testing.assert_array_equal(T2, NON_CONVEX_T)
    # A mirroring scale swaps local vertices 0 and 2 of every tetrahedron (vtk_to_volume_mesh.cc:44-49): volumes stay positive.
    V3, T3 = _read(tmp_path, "vtk", NON_CONVEX_VTK, (-1, 1, 1))
    np.testing.assert_array_equal(T3, NON_CONVEX_T[:, [2, 1, 0, 3]])
    assert np.all(_tet_volumes(V3, T3) > 0)
    V4, T4 = _read(tmp_path, "vtk", NON_CONVEX_VTK, (-1, -1, 1))   # two mirrors: a rotation, no swap
    np.testing.assert_array_equal(T4, NON_CONVEX_T)
    assert np.all(_tet_volumes(V4, T4) > 0)


def test_vtk_offsets_connectivity_layout_and_errors(tmp_path):
    new_layout = NON_CONVEX_VTK.replace("# vtk DataFile Version 2.0", "# vtk DataFile Version 5.1")
    head, rest = new_layout.split("CELLS 6 30")
    cells = "CELLS 7 24\nOFFSETS vtktypeint64\n0 4 8 12 16 20 24\nCONNECTIVITY vtktypeint64\n" + \
        " ".join(" ".join(str(v) for v in t) for t in NON_CONVEX_T) + "\n\nCELL_TYPES 6" + rest.split("CELL_TYPES 6")[1]
    V, T = _read(tmp_path, "vtk", head + cells)
    np.testing.assert_array_equal(V, NON_CONVEX_V)
    np.testing.assert_array_equal(T, NON_CONVEX_T)
    # A hexahedron among the cells: the reference's message (vtk_to_volume_mesh.cc:55-62).
    hexa = NON_CONVEX_VTK.replace("CELLS 6 30", "CELLS 6 30").replace("CELL_TYPES 6\n10", "CELL_TYPES 6\n12")
    with pytest.raises(RuntimeError, match=r"should only contain tetrahedra \(type id=10\)\. Read cell with type id=12"):
        _read(tmp_path, "vtk", hexa)
    with pytest.raises(RuntimeError, match="only ASCII"):
        _read(tmp_path, "vtk", NON_CONVEX_VTK.replace("\nASCII\n", "\nBINARY\n"))
    with pytest.raises(RuntimeError, match="UNSTRUCTURED_GRID"):
        _read(tmp_path, "vtk", NON_CONVEX_VTK.replace("UNSTRUCTURED_GRID", "POLYDATA"))
    with pytest.raises(RuntimeError, match="not a legacy VTK file"):
        _read(tmp_path, "vtk", QUAD_CUBE_OBJ)


def test_obj_reference_fixture(tmp_path):
    """obj_to_surface_mesh_test.cc:26-110: quad_cube.obj scaled by 2 -- the eight vertices in file order, the twelve
    triangles tinyobjloader makes of the six quads."""
    V, T = _read(tmp_path, "obj", QUAD_CUBE_OBJ, (2, 2, 2))
    np.testing.assert_array_equal(V, 2 * QUAD_CUBE_V)
    np.testing.assert_array_equal(T, QUAD_CUBE_T)
    # Outward windings: the enclosed volume is the cube's (CalcEnclosedVolume, meshing_utilities.cc:48-60).
    vol = np.sum(np.einsum("ij,ij->i", np.cross(V[T[:, 0]], V[T[:, 1]]), V[T[:, 2]])) / 6
    assert abs(vol - 64.0) <= 1e-12
    Vm, Tm = _read(tmp_path, "obj", QUAD_CUBE_OBJ, (1, -1, 1))   # a mirroring scale reverses every winding (read_obj.cc:58-68)
    np.testing.assert_array_equal(Tm, QUAD_CUBE_T[:, ::-1])
    assert np.sum(np.einsum("ij,ij->i", np.cross(Vm[Tm[:, 0]], Vm[Tm[:, 1]]), Vm[Tm[:, 2]])) / 6 > 0


def test_obj_index_forms_quads_polygons_and_errors(tmp_path):
    text = "v 0 0 0\nv 2 0 0\nv 2 1 0\nv 0 1 0\nv 1 2 0\nvt 0 0\nvn 0 0 1\n" \
           "f 1/1/1 2/1/1 3/1/1\nf -5 -4 -3 -2\nf 1/1 2/1 3/1 5/1 4/1\n# a comment\n"
    V, T = _read(tmp_path, "obj", text)
    assert len(V) == 5
    # triangle as is; the 2 x 1 rectangle has equal diagonals -> the 1-3 split; the pentagon is fanned about its first vertex
    np.testing.assert_array_equal(T, [[0, 1, 2], [0, 1, 3], [1, 2, 3], [0, 1, 2], [0, 2, 4], [0, 4, 3]])
    # a quad whose 0-2 diagonal is strictly shorter is split along it
    kite = "v 0 0 0\nv 2 -1 0\nv 1 0 0\nv 2 1 0\nf 1 2 3 4\n"
    _, Tk = _read(tmp_path, "obj", kite)
    np.testing.assert_array_equal(Tk, [[0, 1, 2], [0, 2, 3]])
    with pytest.raises(RuntimeError, match="contains no objects"):
        _read(tmp_path, "obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\n")
    with pytest.raises(RuntimeError, match="out of range"):
        _read(tmp_path, "obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n")
