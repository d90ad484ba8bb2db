"""The C++ mirror of the reference interface (include/drake_b200_proximity_engine.h): compiled with
g++ against the C-ABI library; on a box without a GPU it must fail loudly (no CPU fallback), on the
B200 box it must reproduce the reference's known answers (tests/cpp/test_proximity_engine.cc)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_proximity_engine")


def build_exe():
    from drake_b200 import build
    lib = build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_proximity_engine.cc")
    inc = os.path.join(ROOT, "include")
    newest = max([os.path.getmtime(src), os.path.getmtime(lib)] + [os.path.getmtime(os.path.join(inc, h)) for h in os.listdir(inc)])
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                        "-o", EXE, "-L", os.path.dirname(lib), "-ldrake_b200_hydro",
                        "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return EXE


def test_cpp_adapter_compiles_and_has_no_cpu_fallback():
    import torch
    exe = build_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 3, res.stdout
    assert "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_cpp_adapter_known_answers():
    res = subprocess.run([build_exe()], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_cpp_adapter_inflates_compliant_primitives_by_the_margin():
    """A compliant Sphere / Box / Capsule / Cylinder with margin > 0 is meshed and given its pressure field on the shape
    inflated by the margin (hydroelastic_internal.cc:412-508): the adapter's surfaces must match the oracle's inflated
    construction; a rigid Cylinder is the boundary of MakeCylinderVolumeMeshWithMa."""
    import math

    import numpy as np

    from oracle import ho
    R, margin, E = 0.25, 0.01, 1e8
    X_WA = ho.xform(ho.rpy_matrix(math.pi, 0, 0))
    X_WB = ho.xform(None, [0, 0, -(R + R) + 0.125])
    rigid = ho.TriMesh.sphere(R, 0.1)
    args = []
    ms = ho.VolumeMesh.sphere(R + margin, 0.25)
    fs = ho.VolumeField.sphere(ms, R + margin, E, margin)
    mb = ho.VolumeMesh.box_ma(0.5 + 2 * margin, 0.4 + 2 * margin, 0.5 + 2 * margin)
    fb = ho.VolumeField.box(mb, 0.5 + 2 * margin, 0.4 + 2 * margin, 0.5 + 2 * margin, E, margin)
    # Capsule(0.2, 0.3) -> Capsule(radius + margin, length); Cylinder(0.25, 0.4) -> Cylinder(radius + margin, length +
    # 2 margin) (hydroelastic_internal.cc:465-508); the last case puts a RIGID Cylinder(0.2, 0.3) (MakeCylinderSurfaceMesh
    # at the rigid resolution hint 0.1) under the compliant sphere.
    mc = ho.VolumeMesh.capsule(0.2 + margin, 0.3, 0.25)
    fc = ho.VolumeField.capsule(mc, 0.2 + margin, 0.3, E, margin)
    my = ho.VolumeMesh.cylinder_ma(0.25 + margin, 0.4 + 2 * margin, 0.25)
    fy = ho.VolumeField.cylinder(my, 0.25 + margin, 0.4 + 2 * margin, E, margin)
    rigid_cylinder = ho.VolumeMesh.cylinder_ma(0.2, 0.3, 0.1).to_surface()
    for m, f, r in ((ms, fs, rigid), (mb, fb, rigid), (mc, fc, rigid), (my, fy, rigid), (ms, fs, rigid_cylinder)):
        want = ho.soft_rigid(m, f, r, X_WA, X_WB, 1, 2, ho.KPOLYGON)
        assert want is not None and want.num_faces > 4
        args += [str(want.num_faces), repr(float(want.total_area)), repr(float(np.sum(want.e_centroid)))]
    # The nominal (un-inflated) construction gives a different surface: the test can tell the two apart.
    nominal = ho.soft_rigid(ho.VolumeMesh.sphere(R, 0.25), ho.VolumeField.sphere(ho.VolumeMesh.sphere(R, 0.25), R, E, margin),
                            rigid, X_WA, X_WB, 1, 2, ho.KPOLYGON)
    assert abs(nominal.total_area - float(args[1])) > 1e-6 * float(args[1])
    res = subprocess.run([build_exe(), "margin"] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_cpp_adapter_compliant_convex_and_volume_mesh_without_pressure():
    """The reference's compliant Convex (MakeConvexVolumeMesh + MakeVolumeMeshPressureField) and compliant Mesh(.vtk) (the
    given tetrahedra + MakeVolumeMeshPressureField), hydroelastic_internal.cc:555-640, through the C++ mirror: surfaces
    against a rigid sphere equal the oracle's construction."""
    import math

    import numpy as np

    from oracle import ho
    R, E = 0.25, 1e8
    X_WA = ho.xform(ho.rpy_matrix(math.pi, 0, 0))
    X_WB = ho.xform(None, [0, 0, -(R + R) + 0.125])
    rigid = ho.TriMesh.sphere(R, 0.1)
    convex = ho.VolumeMesh.box_ma(0.5, 0.4, 0.5).to_surface().to_convex_volume()
    grid = ho.VolumeMesh.box_grid(0.5, 0.4, 0.5, 0.125)
    args = []
    for m in (convex, grid):
        want = ho.soft_rigid(m, ho.VolumeField.from_mesh(m, E), rigid, X_WA, X_WB, 1, 2, ho.KPOLYGON)
        assert want is not None and want.num_faces > 4
        args += [str(want.num_faces), repr(float(want.total_area)), repr(float(np.sum(want.e_centroid)))]
    # The reference's RIGID Mesh(.vtk) (hydroelastic_internal.cc:375-393): the boundary of the grid mesh, under a compliant sphere.
    ball = ho.VolumeMesh.sphere(R, 0.25)
    want = ho.soft_rigid(ball, ho.VolumeField.sphere(ball, R, E), grid.to_surface(), X_WB, X_WA, 1, 2, ho.KPOLYGON)
    assert want is not None and want.num_faces > 4
    args += [str(want.num_faces), repr(float(want.total_area)), repr(float(np.sum(want.e_centroid)))]
    res = subprocess.run([build_exe(), "convex"] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


def test_traction_mirror_compiles():
    """tests/cpp/test_traction_engine.cc (ComputeHydroelasticContactForces of the C++ mirror) compiles against the
    C-ABI with -Werror; without a GPU it must stop at context creation like every other entry point."""
    import torch
    from drake_b200 import build
    lib = build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_traction_engine")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_traction_engine.cc"), "-o", exe, "-L", os.path.dirname(lib),
                    "-ldrake_b200_hydro", "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    if torch.cuda.is_available():
        pytest.skip("GPU present: executed by tests/test_zz_gpu_wrench.py")
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 3 and "no CPU fallback" in res.stdout


def build_deformable_exe():
    from drake_b200 import build
    lib = build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_deformable_contact")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_deformable_contact.cc")
    inc = os.path.join(ROOT, "include")
    newest = max([os.path.getmtime(src), os.path.getmtime(lib)] + [os.path.getmtime(os.path.join(inc, h)) for h in os.listdir(inc)])
    if not os.path.exists(exe) or os.path.getmtime(exe) < newest:
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                        "-o", exe, "-L", os.path.dirname(lib), "-ldrake_b200_hydro",
                        "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


def test_deformable_mirror_compiles_and_has_no_cpu_fallback():
    """include/drake_b200_deformable_contact.h (deformable::Geometries, DeformableContact, DeformableContactSurface,
    ContactParticipation) compiles warning-free against the C-ABI; without a GPU it fails loudly."""
    import torch
    exe = build_deformable_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 3, res.stdout
    assert "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_deformable_mirror_reference_fixtures():
    """tests/cpp/test_deformable_contact.cc: the reference's OnePolygon / NoContact / InvertedElements fixtures
    (deformable_mesh_intersection_test.cc) and the invariants of deformable x deformable surfaces through the mirror."""
    res = subprocess.run([build_deformable_exe()], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("ok")


def test_registration_property_checks_follow_the_reference():
    """ProximityEngine::ValidateProperties (static, no device): the reference's BadResolutionHint / BadElasticModulus /
    BadSlabThickness suites and its NonNegativeDouble margin check, with its messages
    (hydroelastic_internal_test.cc:905-996,1406-1466; hydroelastic_internal.cc:280-311)."""
    from drake_b200 import build
    lib = build.build()
    src = os.path.join(ROOT, "tests", "cpp", "validate_host_check.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "validate_host_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                    "-o", exe, "-L", os.path.dirname(lib), "-ldrake_b200_hydro", "-Wl,-rpath," + os.path.dirname(lib)],
                   check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout
