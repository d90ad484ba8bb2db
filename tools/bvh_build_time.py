"""Build time and shape of the two BVH builders per mesh asset (registration = records + BVH + upload), and the box
tests / time of one query over each: the host median-split builder (the reference's strategy, bvh.cc:18-136) against the
device-built linear BVH (hydro_lbvh.cu).   python tools/bvh_build_time.py   (GPU box)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drake_b200 as db  # noqa: E402
from drake_b200 import engine, workloads as W  # noqa: E402

out = []
meshes = {"sphere 512 tets (hint 0.015, sparse)": lambda: engine.Mesh.sphere_volume(0.05, 0.015, 0).set_sphere_pressure(0.05, 3e4),
          "sphere 32768 tets (dense)": lambda: engine.Mesh.sphere_volume(3.0, 0.5, 1).set_sphere_pressure(3.0, 1e5),
          "sphere 2097152 tets (dense)": lambda: engine.Mesh.sphere_volume(1.0, 0.03, 1).set_sphere_pressure(1.0, 1e5)}
for name, make in meshes.items():
    m = make()
    for builder, bname in ((engine.BVH_HOST_MEDIAN_SPLIT, "host median split"), (engine.BVH_GPU_LBVH, "device LBVH")):
        ctx = db.Context(0)
        ctx.set_bvh_builder(builder)
        ctx.register_mesh(m)           # warm (context creation, first CUDA calls)
        t0 = time.perf_counter()
        a = ctx.register_mesh(m)
        dt = time.perf_counter() - t0
        info = ctx.asset_info(a)
        out.append({"mesh": name, "builder": bname, "register_ms": 1e3 * dt, "nodes": info.num_bvh_nodes, "height": info.bvh_height})
for cfg, scenes in (("c3", 16384), ("c5", 16)):
    for builder, bname in ((engine.BVH_HOST_MEDIAN_SPLIT, "host median split"), (engine.BVH_GPU_LBVH, "device LBVH")):
        ctx = db.Context(0)
        ctx.set_bvh_builder(builder)
        wl = W.CONFIGS[cfg](scenes)
        b = wl.register(ctx)
        b.set_poses(wl.poses.reshape(-1, 12))
        b.enable_timing(True)
        for _ in range(3):
            b.run(wl.representation, engine.STRICT)
        ms = []
        for _ in range(5):
            b.run(wl.representation, engine.STRICT)
            ms.append(b.stage_times_ms()["total"])
        st = b.stats()
        out.append({"query": f"{cfg}:{scenes}", "builder": bname, "ms": min(ms), "bv_pair_tests": st["bv_pair_tests"],
                    "leaf_pairs": st["leaf_pairs"], "candidates": st["tet_tri_pairs"] + st["tet_tet_pairs"],
                    "stages": {k: round(v, 4) for k, v in b.stage_times_ms().items()}})
for o in out:
    print(json.dumps(o))
