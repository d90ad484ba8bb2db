#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_zz_gpu_wrench.py "tests/test_gpu_parity.py" -k "wrench or discrete or spatial or normal_filter or config3_ball or kat" -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
( time timeout 1500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | tail -3
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"])
print("stages", {k: round(v,4) for k,v in d["roofline"]["kernel_ms_breakdown"].items()})
print("roofline", d["roofline"]["group"], d["roofline"]["frac"], {k: (round(v["ms"],3), round(v["frac"] or 0,3)) for k,v in d["roofline"]["groups"].items()})
e=d["e2e"]
print("e2e full", e["ms_per_step"], e["value"], e["d2h_bytes_per_step"])
for k in ("faces_only","spatial_forces","discrete_pairs"): print(" ", k, e[k]["ms_per_step"], e[k]["value"], e[k]["d2h_bytes_per_step"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k,v in (d["other_configs"] or {}).items():
    if "error" in v: print(k, v); continue
    print(k, "ms", round(v["ms_per_step"],4), "surf/s", round(v["value"],1), "pairs/s", "%.3g"%v["element_pairs_per_sec"], "chunks", v["chunks"], "e2e", (v.get("e2e") or {}).get("ms_per_step"), "cpu", (v.get("cpu_baseline") or {}).get("value"), "cpu1 ms/scene", ((v.get("cpu_baseline") or {}).get("single_thread") or {}).get("ms_per_scene"), "t", round(v["bench_seconds"],1), "roof", v["roofline"]["group"], round(v["roofline"]["frac"],3))
PY
