"""Per-stage device times of a config under different DBH_TUNE launch-bound variants (GPU box).

  python tools/tune.py c3:16384 "traverse=3,filter=3,pack=3" "traverse=4,filter=4,pack=4" ...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import drake_b200 as db
from drake_b200 import engine, workloads as W

name, n = sys.argv[1].split(":")
variants = sys.argv[2:] or [""]
ctx = db.Context(0)
wl = W.CONFIGS[name](int(n))
b = wl.register(ctx)
b.set_poses(wl.poses.reshape(-1, 12))
b.run(wl.representation, engine.STRICT)
b.enable_timing(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
keys = ["pose_root", "traverse", "clip_tet_tri", "clip_tet_tet", "scan", "pack", "surface", "total"]
print("stats", {k: v for k, v in b.stats().items() if k in ("bv_pair_tests", "leaf_pairs", "tet_tri_pairs", "tet_tet_pairs",
                                                           "tet_tri_full_clips", "tet_tet_full_clips", "polygons",
                                                           "leaf_exact_tests")})
print(f"{'variant':40s} " + " ".join(f"{k[:9]:>9s}" for k in keys))
for v in variants:
    os.environ["DBH_TUNE"] = v
    acc = np.zeros(8)
    reps = 8
    for i in range(reps + 2):
        flush.zero_()
        torch.cuda.synchronize()
        b.run(wl.representation, engine.STRICT)
        if i >= 2:
            t = b.kernel_times_ms()
            acc += np.array([t[k] for k in keys])
    acc /= reps
    print(f"{v:40s} " + " ".join(f"{x:9.3f}" for x in acc), flush=True)
