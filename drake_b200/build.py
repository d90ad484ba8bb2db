"""Builds libdrake_b200_hydro.so in-tree with nvcc for sm_100a.

The library is the whole product: hand-written CUDA kernels plus the C-ABI
(include/drake_b200_hydro.h). -fmad=false / -ffp-contract=off: the arithmetic
contract with the reference's scalar CPU path forbids FMA contraction.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdrake_b200_hydro.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "drake_b200_hydro.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libdrake_b200_hydro.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    # DBH_EXTRA_NVCC_FLAGS: debug builds only (e.g. -DDBH_TRACE).
    extra_env = os.environ.get("DBH_EXTRA_NVCC_FLAGS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + extra_env + ["-o", LIB_PATH] + sources()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose and res.stdout:
        print(res.stdout, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ()))
