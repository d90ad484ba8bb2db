#!/bin/bash
# Second visit of the surface-seeded intersector: tests, config 5 by chunk count, ncu capture of k_bfs_wave.   tools/gpu_bfs3.sh <tag>
tag=${1:-bfs3}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${tag}_tests.log 2>&1
tail -8 gpurun_out/${tag}_tests.log
run() {  # name, bench args
  name=$1; shift
  timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    c = d["config"]
    print("${name}".ljust(22), "ms", round(d["ms_per_step"], 3), "surf/step", d["surfaces_per_step"], "faces/s %.3g" % d["faces_per_sec"],
          "chunks", c.get("chunks"), "bfs pairs", c.get("bfs_pairs_per_step"), "waves", c.get("bfs_waves_per_step"),
          {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_breakdown"].items() if v > 0.05 and k != "total"})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${tag}_${name}.err").read()[-1500:])
PY
}
run c5_surface_1024_c1 --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 1
run c5_surface_1024_c2 --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 2
run c5_surface_128_c1 --config c5 --scenes 128 --steps 5 --warmup 3 --soft-soft surface --chunks 1
run c4_surface_512 --config c4 --scenes 512 --steps 5 --warmup 3 --soft-soft surface
# ncu: four mid-search waves of a 128-scene chunk (launch list of the whole query first).
export DBH_SOFT_SOFT=surface DBH_CHUNKS=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches_c5s.csv python tools/profile_one.py c5:128 1 > gpurun_out/${tag}_launches_c5s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_bfs_wave' \
  --launch-skip 60 --launch-count 4 -o /tmp/${tag}_full_c5s -f python tools/profile_one.py c5:128 1 > gpurun_out/${tag}_full_c5s.log 2>&1
python tools/ncu_summary.py /tmp/${tag}_full_c5s.ncu-rep gpurun_out/${tag}_ncu_full_c5s_summary.csv | tail -3
python tools/ncu_lines.py /tmp/${tag}_full_c5s.ncu-rep "regex:k_bfs_wave" "" 45 > gpurun_out/${tag}_lines_c5s_k_bfs_wave.txt 2>&1
python tools/ncu_summary.py --launches gpurun_out/${tag}_launches_c5s.csv gpurun_out/${tag}_launchlist_c5s.csv | tail -1
head -12 gpurun_out/${tag}_lines_c5s_k_bfs_wave.txt
