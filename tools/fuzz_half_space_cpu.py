"""CPU fuzz of the half-space kernel arithmetic (drake_b200/csrc/hydro_plane.cuh compiled for the host,
tests/cpp/plane_host_check.cc) against the oracle: box / sphere / ellipsoid meshes under random AND exactly aligned
poses (identity, axis-aligned rotations, translations that put mesh faces on the boundary).

  python tools/fuzz_half_space_cpu.py [seed] [seconds]

Round-1 finding: every mismatch it reported was a degenerate contact polygon (area ~ 0: its normal, centroid and
the fan hung on that centroid are quotients of rounding noise, different for the oracle's welded vertices and the
kernels' per-polygon ones) -- tests/test_plane_math_cpu.py::compare_with_oracle masks exactly those; surfaces that
consist of noise only (the boundary lying exactly on a zero-pressure mesh face) are skipped here."""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ho
from tests.parity import random_pose
from tests.test_plane_math_cpu import host_mesh_vs_half_space, host_half_space_vs_mesh, compare_with_oracle
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
meshes = []
m = ho.VolumeMesh.box_grid(0.08, 0.06, 0.1, 0.02); meshes.append((m, ho.VolumeField.box(m, 0.08, 0.06, 0.1, 1e5)))
m = ho.VolumeMesh.box_ma(0.3, 0.3, 0.05); meshes.append((m, ho.VolumeField.box(m, 0.3, 0.3, 0.05, 3e4)))
m = ho.VolumeMesh.sphere(0.05, 0.015); meshes.append((m, ho.VolumeField.sphere(m, 0.05, 3e4)))
m = ho.VolumeMesh.ellipsoid(0.03, 0.05, 0.02, 0.01, dense=True); meshes.append((m, ho.VolumeField.ellipsoid(m, 0.03, 0.05, 0.02, 1e5)))
surfs = [mm.to_surface() for mm,_ in meshes]
def special_pose(rng):
    k = rng.integers(0, 4)
    if k == 0: return ho.IDENTITY
    if k == 1: return ho.xform(None, rng.choice([0.0, 0.01, -0.025, 0.05], size=3))
    if k == 2: return ho.xform(ho.rpy_matrix(*(rng.choice([0, math.pi/2, math.pi, -math.pi/2], size=3))), rng.choice([0.0, 0.02, -0.01], size=3))
    return random_pose(rng, 0.05)
t0=time.time(); n=hits=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 40:
    i = int(rng.integers(len(meshes))); mesh, fld = meshes[i]
    X_WS, X_WR = special_pose(rng), special_pose(rng)
    rep = int(rng.integers(2)); swapped = bool(rng.integers(2)); ids = (2,1) if swapped else (1,2)
    ora = ho.soft_mesh_rigid_half_space(mesh, fld, X_WS, X_WR, ids[0], ids[1], rep)
    host = host_mesh_vs_half_space(mesh, fld, X_WS, X_WR, rep, swapped)
    if ora is None: assert host.counts[0]==0
    elif ora.total_area > 1e-9 and np.abs(ora.pressures).max() > 1e-3:
        compare_with_oracle(host, ora, ora.elem_first, ora.grad_eN if swapped else ora.grad_eM); hits+=1
    margin = float(rng.choice([0.0, 1e-3]))
    try:
        ora = ho.soft_half_space_rigid_mesh(surfs[i], X_WS, 2e5, X_WR, ids[0], ids[1], rep, margin=margin)
    except RuntimeError as e:
        print("oracle raised (reference DRAKE_DEMAND):", str(e)[:80]); continue
    host = host_half_space_vs_mesh(surfs[i], X_WS, 2e5, margin, X_WR, rep, swapped)
    if ora is None: assert host.counts[0]==0
    elif ora.total_area > 1e-9 and np.abs(ora.pressures).max() > 1e-3:
        compare_with_oracle(host, ora, ora.elem_second, ora.grad_eN if swapped else ora.grad_eM); hits+=1
    n+=1
print("cases", n, "non-null", hits)
