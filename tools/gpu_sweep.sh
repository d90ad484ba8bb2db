#!/bin/bash
# Variant sweep of one visit: tools/gpu_sweep.sh <tag> "<DBH_TUNE variants separated by ;>" [config]   (bench only, no tests)
tag=$1; variants=$2; cfg=${3:-c3}
mkdir -p gpurun_out
IFS=';' read -ra VS <<< "$variants"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  DBH_TUNE="$v" timeout 300 python bench.py --config $cfg $BENCH_ARGS --no-cpu-baseline --no-e2e --steps 20 --warmup 5 \
    > gpurun_out/${tag}_bench_$i.json 2> gpurun_out/${tag}_bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_$i.json"))
    print("[$i] $v".ljust(28), round(d["ms_per_step"],4), {k: round(v,3) for k,v in d["roofline"]["kernel_ms_breakdown"].items() if v > 0.01 and k != "total"})
except Exception as e:
    print("[$i] $v bench failed", e); print(open("gpurun_out/${tag}_bench_$i.err").read()[-1500:])
PY
done
