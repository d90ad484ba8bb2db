"""CPU fuzz of the exact stage's polygon primitives (drake_b200/csrc/hydro_clip.cuh compiled for the host,
tests/cpp/plane_host_check.cc) against the oracle, BIT FOR BIT: whole IntersectTetrahedra and
ClipTriangleByTetrahedron sequences over the candidate pairs of grid-box / sphere meshes under random and exactly
aligned poses (coplanar faces, zero-gradient boundary tetrahedra).

  python tools/fuzz_exact_stage_cpu.py <seed> <seconds>

Round 1: 27 M element pairs / 4.9 M polygons over four 7-minute seeds (one polygon with more than 8 vertices), no
mismatch."""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ho
from tests.parity import random_pose
import tests.test_plane_math_cpu as T
lib = T.harness(); _dp = T._dp
lib.phc_intersect_tetrahedra.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, _dp]
d = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(_dp)
rng = np.random.default_rng(int(sys.argv[1]))
A = ho.VolumeMesh.box_grid(0.08, 0.06, 0.1, 0.02); fA = ho.VolumeField.box(A, 0.08, 0.06, 0.1, 1e5)
B = ho.VolumeMesh.box_grid(0.05, 0.05, 0.05, 0.025); fB = ho.VolumeField.box(B, 0.05, 0.05, 0.05, 2e5)
S = ho.VolumeMesh.sphere(0.05, 0.02); fS = ho.VolumeField.sphere(S, 0.05, 1e5)
R = ho.TriMesh.box(0.2, 0.2, 0.01); Rs = ho.TriMesh.sphere(0.05, 0.02)
def special_pose(rng):
    k = rng.integers(0, 4)
    if k == 0: return ho.xform(None, rng.choice([0.0, 0.02, -0.02, 0.04, 0.01], size=3))
    if k == 1: return ho.xform(ho.rpy_matrix(*(rng.choice([0, math.pi/2, math.pi, -math.pi/2], size=3))), rng.choice([0.0, 0.02, -0.01, 0.03], size=3))
    return random_pose(rng, 0.04)
t0=time.time(); npairs=npoly=nbig=0
while time.time()-t0 < float(sys.argv[2]):
    (m0,f0),(m1,f1) = [(A,fA),(B,fB),(S,fS)][rng.integers(3)], [(A,fA),(B,fB),(S,fS)][rng.integers(3)]
    X_MN = ho.invert_and_compose(special_pose(rng), special_pose(rng))
    pairs,_ = ho.candidates(m0, m1, X_MN)
    if len(pairs) == 0: continue
    v0,t0_,_ = m0.arrays(); v1,t1_,in1 = m1.arrays()
    for e0,e1 in pairs[rng.permutation(len(pairs))[:400]]:
        plane = ho.calc_equilibrium_plane(f0, int(e0), f1, int(e1), X_MN)
        if plane is None: continue
        nhat, disp = plane
        try:
            want,_ = ho.intersect_tetrahedra(m0, int(e0), m1, int(e1), X_MN, nhat, disp); thrown=False
        except RuntimeError: thrown=True
        out = np.zeros((16,3))
        n = lib.phc_intersect_tetrahedra(d(v0[t0_[e0]]), d(v1[t1_[e1]]), d(in1[e1]), d(X_MN), d(nhat), disp, d(out))
        npairs+=1
        if thrown:
            assert n == -2, n; continue
        assert n >= 0 and n == len(want), (n, len(want), e0, e1)
        assert np.array_equal(out[:n], want), (e0,e1)
        npoly += n>0; nbig += n>8
    # tet-tri
    surf = [R, Rs][rng.integers(2)]
    pairs,_ = ho.candidates(m0, surf, X_MN)
    sv, st = surf.arrays()[:2]
    for tet,tri in pairs[rng.permutation(len(pairs))[:400]] if len(pairs) else []:
        want = ho.clip_triangle_by_tetrahedron(m0, int(tet), surf, int(tri), X_MN)
        out = np.zeros((16,3))
        n = lib.phc_clip_triangle_by_tetrahedron(d(v0[t0_[tet]]), d(sv[st[tri]]), d(X_MN), d(out))
        npairs+=1
        assert n == len(want), (n, len(want)); assert np.array_equal(out[:n], want)
        npoly += n>0
print("pairs", npairs, "polygons", npoly, "with more than 8 vertices", nbig)
