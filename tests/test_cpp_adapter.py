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
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib)):
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
