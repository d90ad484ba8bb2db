"""Steady-state cost of deforming meshes: dbh_context_update_vertices for S deforming assets per call
(host vertices in, per-element records + BVH boxes re-derived on the device), and the query + k_deformable that follow.
GPU box:  python tools/refit_time.py [S]  ->  one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drake_b200 as db
from drake_b200 import engine, workloads as W

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = db.Context(0)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


# Deforming SURFACES against one rigid pressure-field volume per scene (deformable x rigid), S scenes.
vol = W.Mesh.sphere_volume(0.3, 0.1, W.DENSE_INTERIOR_VERTICES).set_sphere_pressure(0.3, 1e6)
surf_src = W.Mesh.sphere_volume(0.25, 0.05, W.DENSE_INTERIOR_VERTICES).to_surface()
sv, st = surf_src.vertices, surf_src.elements
a_vol = ctx.register_mesh(vol)
a_s = [ctx.register_rigid_mesh(sv, st) for _ in range(S)]
rng = np.random.default_rng(7)
X = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], float), (S, 2, 1))
X[:, 0, 9:12] = rng.uniform(-0.05, 0.05, (S, 3)) + [0.1, 0.05, 0.3]
ga = np.array([[a_vol, a_s[s]] for s in range(S)]).reshape(-1)
batch = db.SceneBatch(ctx, np.arange(S + 1) * 2, ga, np.tile([1, 2], S))
new = [sv * (1 + 0.05 * np.sin(s)) for s in range(S)]
flat = [np.ascontiguousarray(v) for v in new]
t_update = timed(lambda: ctx.update_vertices(a_s, flat))
import torch
dev = torch.from_numpy(np.concatenate([v.reshape(-1) for v in flat])).cuda()
torch.cuda.synchronize()
t_update_dev = timed(lambda: ctx.update_vertices_device(a_s, dev.data_ptr()))


def step():
    batch.set_poses(X.reshape(-1, 12))
    batch.run(db.POLYGON, engine.STRICT)        # query + sync (work lists grow during the warm-up calls)
    batch.compute_deformable_contact()
    batch.sync()


t_query = timed(step)
rows, _, _ = batch.copy_deformable_contact()
print(json.dumps({"case": "deformable surfaces x rigid volume", "scenes": S, "triangles_per_surface": int(len(st)),
                  "tets_of_rigid_volume": int(len(vol.elements)), "update_vertices_ms": t_update,
                  "update_vertices_us_per_asset": t_update * 1e3 / S, "update_vertices_device_ms": t_update_dev, "query_plus_contact_data_ms": t_query,
                  "contact_points": int(len(rows))}))

# Deforming VOLUMES (deformable x deformable): two per scene.
va = W.Mesh.sphere_volume(0.25, 0.08, W.SINGLE_INTERIOR_VERTEX)
av, at = va.vertices, va.elements
sdf = np.linalg.norm(av, axis=1) - 0.25
h = [ctx.register_compliant_mesh(av + ([0, 0, 0] if k % 2 == 0 else [0.1, 0.05, 0.42]), at, sdf) for k in range(2 * S)]
b2 = db.SceneBatch(ctx, np.arange(S + 1) * 2, h, np.tile([1, 2], S))
I = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], float), (2 * S, 1))
newv = [np.ascontiguousarray((av * (1 + 0.04 * np.sin(k))) + ([0, 0, 0] if k % 2 == 0 else [0.1, 0.05, 0.42])) for k in range(2 * S)]
t_update2 = timed(lambda: ctx.update_vertices(h, newv))
dev2 = torch.from_numpy(np.concatenate([v.reshape(-1) for v in newv])).cuda()
torch.cuda.synchronize()
t_update2_dev = timed(lambda: ctx.update_vertices_device(h, dev2.data_ptr()))


def step2():
    b2.set_poses(I)
    b2.run(db.POLYGON, engine.STRICT)
    b2.compute_deformable_contact()
    b2.sync()


t_query2 = timed(step2)
rows2, _, _ = b2.copy_deformable_contact()
print(json.dumps({"case": "deformable volumes x deformable volumes", "scenes": S, "tets_per_volume": int(len(at)),
                  "update_vertices_ms": t_update2, "update_vertices_us_per_asset": t_update2 * 1e3 / (2 * S), "update_vertices_device_ms": t_update2_dev,
                  "query_plus_contact_data_ms": t_query2, "contact_points": int(len(rows2))}))
