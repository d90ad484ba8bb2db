#!/bin/bash
# Round-end check of the final build: every GPU test, smoke(), the default bench line (other configs included).   tools/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.log 2>&1
tail -6 gpurun_out/${tag}_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | tail -3
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches", d["gpu_launches"])
print("stages", {k: round(v,4) for k,v in d["roofline"]["kernel_ms_breakdown"].items()})
e=d["e2e"]
print("e2e full", e["ms_per_step"], e["value"], e["d2h_bytes_per_step"])
for k in ("faces_only","spatial_forces","discrete_pairs"): print(" ", k, e[k]["ms_per_step"], e[k]["value"])
c=d["cpu_baseline"]; print("cpu obb", c["value"], c["cores"])
for k,v in (d["other_configs"] or {}).items():
    if "error" in v: print(k, v); continue
    print(k, "ms", round(v["ms_per_step"],4), "surf/s", round(v["value"],1), "chunks", v["chunks"], "t", round(v["bench_seconds"],1), v.get("soft_soft_intersector"), v.get("bfs_pairs_per_step"), v.get("bfs_waves_per_step"))
PY
