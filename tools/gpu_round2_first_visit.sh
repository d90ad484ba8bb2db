#!/bin/bash
# First GPU visit of round 2: everything round 1 wrote after its GPU minutes ran out, in the order that
# decides what to keep.   usage: tools/gpu_round2_first_visit.sh <tag>      (≈ 6 GPU-minutes)
#  1. the `-m gpu` suite as the driver runs it (must stay green);
#  2. k_wrench (parity-green at the very end of round 1, never timed): its tests again, then (step 5) its bench;
#  3. k_plane_count2 (DBH_TUNE plane=2, never run): the half-space tests under it, then c3h bench lines with
#     the default and the experimental count pass side by side;
#  4. launch list + one `--set full` capture of the half-space kernels on c3h (none exists yet).
tag=${1:-r2a}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log
(DBH_RUN_NEW=1 timeout 200 python -m pytest tests/test_zz_gpu_wrench.py -m gpu -q 2>&1 | tail -40) \
  > gpurun_out/${tag}_tests_wrench.log
(DBH_TUNE=plane=2 timeout 200 python -m pytest tests/test_zz_gpu_halfspace.py -m gpu -q 2>&1 | tail -40) \
  > gpurun_out/${tag}_tests_plane2.log
timeout 120 python bench.py --config c3h --no-cpu-baseline --steps 20 --warmup 5 \
  > gpurun_out/${tag}_bench_c3h.json 2> gpurun_out/${tag}_bench_c3h.err
DBH_TUNE=plane=2 timeout 120 python bench.py --config c3h --no-cpu-baseline --steps 20 --warmup 5 \
  > gpurun_out/${tag}_bench_c3h_plane2.json 2> gpurun_out/${tag}_bench_c3h_plane2.err
(DBH_TUNE=ppack=4 timeout 100 python -m pytest tests/test_zz_gpu_halfspace.py -m gpu -q 2>&1 | tail -5) \
  > gpurun_out/${tag}_tests_ppack4.log
DBH_TUNE=plane=2,ppack=4 timeout 120 python bench.py --config c3h --no-cpu-baseline --steps 20 --warmup 5 \
  > gpurun_out/${tag}_bench_c3h_plane2_ppack4.json 2> gpurun_out/${tag}_bench_c3h_plane2_ppack4.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches_c3h.csv python tools/profile_one.py c3h:16384 2 > gpurun_out/${tag}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_plane' \
  -o gpurun_out/${tag}_full_c3h -f python tools/profile_one.py c3h:16384 1 > gpurun_out/${tag}_full_c3h.log 2>&1
# 5. only if step 2 is green: the e2e figure through the consumer (18 doubles per surface out instead of the mesh).
if tail -1 gpurun_out/${tag}_tests_wrench.log | grep -q "passed" && ! tail -1 gpurun_out/${tag}_tests_wrench.log | grep -q "failed"; then
  timeout 200 python bench.py --no-cpu-baseline --forces --steps 20 --warmup 5 \
    > gpurun_out/${tag}_bench_c3_forces.json 2> gpurun_out/${tag}_bench_c3_forces.err
  echo "== bench c3 --forces"; cut -c1-260 gpurun_out/${tag}_bench_c3_forces.json
fi
for f in tests tests_wrench tests_plane2 tests_ppack4; do echo "== $f"; tail -3 gpurun_out/${tag}_$f.log; done
for f in bench_c3h bench_c3h_plane2 bench_c3h_plane2_ppack4; do echo "== $f"; cut -c1-260 gpurun_out/${tag}_$f.json; done
