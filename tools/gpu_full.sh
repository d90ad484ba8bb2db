#!/bin/bash
# The round-end sequence as the driver runs it: GPU tests, smoke, default bench, reference arm.
tag=${1:-full}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.log 2>&1
tail -6 gpurun_out/${tag}_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 1500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | tail -3
tail -3 gpurun_out/${tag}_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err ) 2>&1 | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches", d["gpu_launches"])
print("stages", {k: round(v,4) for k,v in d["roofline"]["kernel_ms_breakdown"].items()})
r=d["roofline"]; print("roofline", r["group"], r["frac"], "issue", r["issue_slots"], "traffic", r["traffic"], r["compulsory_dram"])
e=d["e2e"]
print("e2e full", e["ms_per_step"], e["value"], e["d2h_bytes_per_step"])
for k in ("faces_only","spatial_forces","discrete_pairs"): print(" ", k, e[k]["ms_per_step"], e[k]["value"], e[k]["d2h_bytes_per_step"])
c=d["cpu_baseline"]; print("cpu obb", c["value"], c["cores"], "aabb", c["aabb_hierarchy"]["value"], "1thr", c["single_thread"]["value"])
for k,v in (d["other_configs"] or {}).items():
    if "error" in v: print(k, v); continue
    cb=v.get("cpu_baseline") or {}
    print(k, "ms", round(v["ms_per_step"],4), "surf/s", round(v["value"],1), "pairs/s", "%.3g"%v["element_pairs_per_sec"], "chunks", v["chunks"], "e2e", (v.get("e2e") or {}).get("ms_per_step"), "cpu", cb.get("value"), "cpu1 ms/scene", (cb.get("single_thread") or {}).get("ms_per_scene"), "t", round(v["bench_seconds"],1), "roof", v["roofline"]["group"], round(v["roofline"]["frac"],3))
r=json.load(open("gpurun_out/${tag}_ref.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
PY
