"""Diagnostic: C4 scenes, normal vs self-check mode, against the oracle (GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drake_b200 as db
from drake_b200 import workloads as W
from oracle import ho
from tests import parity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = db.Context(0)
wl = W.config4(n)
batch = wl.register(ctx)
osc, order = parity.oracle_scene_for(wl)
want = [osc.compute(wl.poses[s][order], ho.KPOLYGON)[0] for s in range(wl.num_scenes)]
for verify in (False, True):
    batch.set_verify(verify)
    got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
    st = batch.stats()
    print("verify", verify, {k: st[k] for k in ("tet_tri_pairs", "tet_tet_pairs", "tet_tri_full_clips", "tet_tet_full_clips",
                                                 "polygons", "cull_violations", "exceptional_clips", "leaf_pairs")})
    bad = 0
    for s in range(wl.num_scenes):
        assert [(x.id_M, x.id_N) for x in got[s]] == [(x.id_M, x.id_N) for x in want[s]], "surface id lists differ"
        for g, w in zip(got[s], want[s]):
            if g.num_faces != w.num_faces:
                bad += 1
                gk = set(parity.face_keys(g.elem_M, g.elem_N).tolist())
                ok = set(parity.face_keys(w.elem_first, w.elem_second).tolist())
                miss = sorted(ok - gk)
                extra = sorted(gk - ok)
                print(f"  scene {s} pair ({g.id_M},{g.id_N}) faces {g.num_faces} vs {w.num_faces}; missing "
                      f"{[(k >> 32, k & 0xffffffff) for k in miss][:8]} extra {[(k >> 32, k & 0xffffffff) for k in extra][:8]}")
    print("  mismatching surfaces:", bad)

# Where do the missing pairs get lost? (leaf list / candidate flag / exact stage)
batch.set_verify(False)
got = batch.compute_contact_surfaces(wl.poses.reshape(-1, 12), ho.KPOLYGON)
gp = batch.geometry_pairs()
cand = batch.candidates(1)
leaf = batch.candidates(3)
gid = np.asarray(wl.geom_id)
G = wl.geoms_per_scene
for (s, a, b, prs) in [(2, 102, 107, [(84, 104), (84, 105), (101, 102)]), (2, 109, 111, [(155, 129)])]:
    for pi, (sc, ga, gb, kind) in enumerate(gp):
        if sc == s and {gid[ga % G], gid[gb % G]} == {a, b}:
            first = gid[ga % G]
            for (t0, t1) in prs:
                in_leaf = bool(np.any((leaf[:, 0] == pi) & (leaf[:, 1] == t0) & (leaf[:, 2] == t1)))
                in_cand = bool(np.any((cand[:, 0] == pi) & (cand[:, 1] == t0) & (cand[:, 2] == t1)))
                print(f"pair {pi} first body id {first} kind {kind}: ({t0},{t1}) leaf={in_leaf} cand={in_cand}")
# Derived gradients bitwise vs the oracle for every compliant mesh of the palette.
cache = {}
for g, mi in enumerate(wl.geom_mesh):
    m = wl.meshes[mi]
    if m.vertices_per_element != 4 or mi in cache:
        continue
    v, e, p = m.arrays()
    om = ho.VolumeMesh.from_arrays(v, e)
    of = ho.VolumeField.from_values(om, p)
    cache[mi] = 1
    grad, at_mo, inward = ctx.asset_derived(mi)
    _, og, om_, _ = of.arrays()
    print("mesh", mi, "grad bitwise", np.array_equal(grad, og), "p_mo bitwise", np.array_equal(at_mo, om_),
          "max |grad diff|", np.abs(grad - og).max(), "min |grad|", np.linalg.norm(og, axis=1).min())
