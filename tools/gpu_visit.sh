#!/bin/bash
# One GPU visit: tools/gpu_visit.sh <tag> "<pytest args or ->" "<DBH_TUNE variants separated by ;>" "<ncu kernel regex or ->" [config:scenes]
tag=$1; tests=$2; variants=$3; ncuk=$4; cfg=${5:-c3:16384}
mkdir -p gpurun_out
if [ "$tests" != "-" ]; then
  (timeout 1200 python -m pytest $tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${tag}_tests.log
  tail -4 gpurun_out/${tag}_tests.log
fi
IFS=';' read -ra VS <<< "$variants"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  DBH_TUNE="$v" timeout 300 python bench.py --config ${cfg%%:*} --no-cpu-baseline --no-e2e --steps 20 --warmup 5 \
    > gpurun_out/${tag}_bench_$i.json 2> gpurun_out/${tag}_bench_$i.err
  echo "== [$i] DBH_TUNE=$v"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_$i.json"))
    print(round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_breakdown"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench_$i.err").read()[-1500:])
PY
done
if [ "$ncuk" != "-" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_one.py $cfg 1 > gpurun_out/${tag}_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$ncuk" \
    -o gpurun_out/${tag}_full -f python tools/profile_one.py $cfg 1 > gpurun_out/${tag}_full.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
for r in rows[1:]: print(r[ik][:48].ljust(50), r[iv])
PY
  tail -2 gpurun_out/${tag}_full.log
fi
