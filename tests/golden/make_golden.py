"""Generates tests/golden/*.json from the ORACLE (test infrastructure) -- the reference itself cannot be built or
imported here (DESIGN.md "Oracle"), so these fixtures pin the oracle's own outputs on the BASELINE workloads: a change
of the oracle, of the mesh generators or of the workload RNG shows up as a diff. Run from the repo root:

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from drake_b200 import workloads as W  # noqa: E402
from oracle import ho  # noqa: E402
from tests import parity  # noqa: E402

CASES = {"c3_16_scenes": (lambda: W.config3(16), ho.KPOLYGON), "c3_16_scenes_tri": (lambda: W.config3(16), ho.KTRIANGLE),
         "c4_2_scenes": (lambda: W.config4(2), ho.KPOLYGON), "c2s_2_scenes": (lambda: W.config2_spheres(2, 0, 1.0), ho.KPOLYGON),
         # Half-space ground. "vertices" is the oracle's WELDED count (the reference welds vertices cut
         # from the same mesh edge); the kernels keep vertices per polygon, so GPU tests do not compare it for these surfaces.
         "c3h_16_scenes": (lambda: W.config3_half_space(16), ho.KPOLYGON),
         "c3h_16_scenes_tri": (lambda: W.config3_half_space(16), ho.KTRIANGLE)}


def digest(workload, rep):
    osc, order = parity.oracle_scene_for(workload)
    scenes = []
    for s in range(workload.num_scenes):
        surfaces, _ = osc.compute(workload.poses[s][order], rep, ho.MODE_TREE)
        scenes.append([{"id_M": int(x.id_M), "id_N": int(x.id_N), "faces": int(x.num_faces), "vertices": int(len(x.vertices)),
                        "total_area": float(x.total_area), "centroid": [float(v) for v in x.centroid],
                        "sum_face_pressure": float(np.sum(x.e_centroid)),
                        "key_checksum": int(np.bitwise_xor.reduce(parity.face_keys(x.elem_first, x.elem_second)
                                                                  * np.int64(2654435761) % np.int64(1 << 61)))}
                       for x in surfaces])
    return scenes


if __name__ == "__main__":
    for name, (make, rep) in CASES.items():
        out = {"case": name, "representation": int(rep), "scenes": digest(make(), rep)}
        path = os.path.join(ROOT, "tests", "golden", name + ".json")
        json.dump(out, open(path, "w"), indent=1)
        print(path, sum(len(s) for s in out["scenes"]), "surfaces")
