#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + full capture (C3), full capture of the
# narrow-phase kernels on C5.   usage: tools/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches_c3.csv python tools/profile_one.py c3:16384 2 > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_clip|k_traverse|k_pack|k_filter|k_surface' \
  -o gpurun_out/${tag}_full_c3 -f python tools/profile_one.py c3:16384 1 > gpurun_out/${tag}_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_clip|k_filter|k_pack|k_surface' \
  -o gpurun_out/${tag}_full_c5 -f python tools/profile_one.py c5:16 1 > gpurun_out/${tag}_full_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches_c5.csv python tools/profile_one.py c5:16 1 > /dev/null 2>&1
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err; ls -la gpurun_out | tail -8
