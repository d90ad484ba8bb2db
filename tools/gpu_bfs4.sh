#!/bin/bash
# Third visit: hash finaliser + 4x table, k_bfs_wave at 7 / 10 resident blocks per SM.   tools/gpu_bfs4.sh <tag>
tag=${1:-bfs4}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${tag}_tests.log 2>&1
tail -4 gpurun_out/${tag}_tests.log
run() {  # name, tune, bench args
  name=$1; tune=$2; shift; shift
  DBH_TUNE="$tune" timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    c = d["config"]
    print("${name}".ljust(24), "ms", round(d["ms_per_step"], 3), "surf/step", d["surfaces_per_step"], "faces/s %.3g" % d["faces_per_sec"],
          "chunks", c.get("chunks"), "bfs pairs", c.get("bfs_pairs_per_step"), "waves", c.get("bfs_waves_per_step"),
          {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_breakdown"].items() if v > 0.05 and k != "total"})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${tag}_${name}.err").read()[-1500:])
PY
}
run c5_surface_1024_occ7 "" --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 1
run c5_surface_1024_occ10 "bfsocc=1" --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 1
run c5_surface_128_occ7 "" --config c5 --scenes 128 --steps 5 --warmup 3 --soft-soft surface --chunks 1
run c5_surface_128_occ10 "bfsocc=1" --config c5 --scenes 128 --steps 5 --warmup 3 --soft-soft surface --chunks 1
run c4_surface_512_occ10 "bfsocc=1" --config c4 --scenes 512 --steps 5 --warmup 3 --soft-soft surface
