"""Runs BASELINE configs once on cuda:0 and prints the query statistics and per-kernel times.

  python tools/config_stats.py c3:16384 c4:256 c2s:64 c5:2
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import drake_b200 as db
from drake_b200 import engine
from drake_b200 import workloads as W


def main():
    ctx = db.Context(0)
    for spec in sys.argv[1:] or ["c3:16384"]:
        name, n = spec.split(":")
        t0 = time.time()
        wl = W.CONFIGS[name](int(n))
        t1 = time.time()
        b = wl.register(ctx)
        t2 = time.time()
        b.set_poses(wl.poses.reshape(-1, 12))
        b.enable_timing(True)
        try:
            for _ in range(3):
                b.run(wl.representation, engine.STRICT)
        except db.HydroError as e:
            print(name, "FAILED:", e, flush=True)
            continue
        st = b.stats()
        ms = b.kernel_times_ms()
        print(f"{name} scenes={n} gen={t1 - t0:.1f}s register={t2 - t1:.1f}s", flush=True)
        print("  stats:", st, flush=True)
        print("  ms:", {k: round(v, 4) for k, v in ms.items()}, flush=True)
        cand = max(1, st["tet_tri_pairs"] + st["tet_tet_pairs"])
        print(f"  full clips / candidates = {(st['tet_tri_full_clips'] + st['tet_tet_full_clips']) / cand:.3f}; "
              f"polygons / candidates = {st['polygons'] / cand:.3f}", flush=True)
        del b


if __name__ == "__main__":
    main()
