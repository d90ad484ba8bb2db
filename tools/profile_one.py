"""One warmed-up query of a config inside a cuProfilerStart/Stop range, for ncu captures:

  ncu --profile-from-start off ... python tools/profile_one.py c4:64 [queries]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drake_b200 as db  # noqa: E402
from drake_b200 import engine, workloads as W  # noqa: E402

name, n = sys.argv[1].split(":")
queries = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = db.Context(0)
if os.environ.get("DBH_SOFT_SOFT", "") == "surface":   # the surface-seeded soft-soft intersector (hydro_bfs.cuh)
    ctx.set_soft_soft_intersector(engine.INTERSECTOR_SURFACE_BFS)
wl = W.CONFIGS[name](int(n))
b = wl.register(ctx)
if os.environ.get("DBH_CHUNKS"):
    b.set_chunks(int(os.environ["DBH_CHUNKS"]))
b.set_poses(wl.poses.reshape(-1, 12))
b.run(wl.representation, engine.STRICT)      # grows the work lists
b.run(wl.representation, engine.STRICT)
cuda = ctypes.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
for _ in range(queries):
    b.run(wl.representation, engine.STRICT)
cuda.cuProfilerStop()
print(b.stats())
