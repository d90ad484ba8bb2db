#!/bin/bash
# Full-size BASELINE configs 4 and 5 (device-resident metric only; the packed results are tens of GB).
tag=${1:-full}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_tests.log
  tail -5 gpurun_out/${tag}_tests.log
fi
for spec in ${SPECS:-c4:4096 c5:1024}; do
  set -- ${spec/:/ }
  DBH_DEBUG=1 timeout 1200 python bench.py --config $1 --scenes $2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
     > gpurun_out/${tag}_bench_$1_$2.json 2> gpurun_out/${tag}_bench_$1_$2.err
  echo "== $1 $2 rc=$?"; cat gpurun_out/${tag}_bench_$1_$2.json; tail -12 gpurun_out/${tag}_bench_$1_$2.err
done
