#!/bin/bash
# Bench lines for the other BASELINE configs (parity-test cases; recorded under profiles/ for reference).
tag=${1:-run}
mkdir -p gpurun_out
for cfg in c4 c5 c2s; do
  DBH_DEBUG=1 timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/${tag}_bench_${cfg}.json 2> gpurun_out/${tag}_bench_${cfg}.err
  echo "== $cfg rc=$?"; cat gpurun_out/${tag}_bench_${cfg}.json; tail -5 gpurun_out/${tag}_bench_${cfg}.err
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
