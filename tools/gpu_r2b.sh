#!/bin/bash
# Round 2 visit b: the on-chip small-pair mid phase (k_small_pairs): parity in both modes, then A/B bench lines.
tag=${1:-r2b}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${tag}_tests.log
for v in "small=0" "small=2,sblocks=4" "small=2,sblocks=3"; do
  n=$(echo $v | tr -d '=,')
  DBH_TUNE=$v timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5 \
    > gpurun_out/${tag}_bench_$n.json 2> gpurun_out/${tag}_bench_$n.err
  echo "== $v"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_$n.json"))
print(d["ms_per_step"], d["roofline"]["kernel_ms_breakdown"])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches_c3.csv python tools/profile_one.py c3:16384 1 > gpurun_out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_small' \
  -o gpurun_out/${tag}_full_small -f python tools/profile_one.py c3:16384 1 > gpurun_out/${tag}_full_small.log 2>&1
tail -5 gpurun_out/${tag}_tests.log
grep -v "^==PROF\|^$" gpurun_out/${tag}_launches_c3.csv | cut -d, -f5,15 | tail -30
