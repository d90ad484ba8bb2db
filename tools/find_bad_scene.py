"""Find C4 scenes whose query fails on the GPU (bisection over scene ranges)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drake_b200 as db
from drake_b200 import workloads as W, engine

ctx = db.Context(0)
assets = None


def ok(scene0, n):
    global assets
    wl = W.config4(n, scene0)
    b = wl.register(ctx, assets)
    assets = wl.assets
    b.set_poses(wl.poses.reshape(-1, 12))
    try:
        b.run(wl.representation, engine.STRICT, max_grow=12)
        return True, b.stats()
    except db.HydroError as e:
        return False, str(e)
    finally:
        b.close()


bad = []
def search(s0, n):
    good, info = ok(s0, n)
    if good:
        return
    if n == 1:
        print("BAD scene", s0, info, flush=True)
        bad.append(s0)
        return
    h = n // 2
    search(s0, h)
    search(s0 + h, n - h)

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for s0 in range(first, first + total, 512):
    search(s0, min(512, first + total - s0))
print("bad scenes:", bad)
