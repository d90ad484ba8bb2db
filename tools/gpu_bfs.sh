#!/bin/bash
# Surface-seeded soft-soft intersector on the B200: its GPU tests, a regression slice of the default path, and the
# A/B of the two intersectors on config 5 (128 scenes), config 2 and the headline config.   tools/gpu_bfs.sh <tag>
tag=${1:-bfs}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${tag}_tests.log 2>&1
tail -25 gpurun_out/${tag}_tests.log
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kat or scene_engine or golden" 2>&1 | tail -8 ) > gpurun_out/${tag}_regress.log 2>&1
tail -6 gpurun_out/${tag}_regress.log
run() {  # name, bench args
  name=$1; shift
  timeout 400 python bench.py "$@" --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}".ljust(22), "ms", round(d["ms_per_step"], 4), "surf/step", d["surfaces_per_step"], "faces/s %.3g" % d["faces_per_sec"],
          {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_breakdown"].items() if v > 0.01 and k != "total"})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${tag}_${name}.err").read()[-1200:])
PY
}
run c5_volume_128 --config c5 --scenes 128 --steps 5 --warmup 3
run c5_surface_128 --config c5 --scenes 128 --steps 5 --warmup 3 --soft-soft surface --chunks 1
run c5_surface_128_c4 --config c5 --scenes 128 --steps 5 --warmup 3 --soft-soft surface --chunks 4
run c3_surface --config c3 --steps 10 --warmup 3 --soft-soft surface
run c3_volume --config c3 --steps 10 --warmup 3
