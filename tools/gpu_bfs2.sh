#!/bin/bash
# Config 5 at its full batch size through the surface-seeded intersector, by chunk count; config 4 A/B.   tools/gpu_bfs2.sh <tag>
tag=${1:-bfs2}
mkdir -p gpurun_out
run() {  # name, bench args
  name=$1; shift
  ( time timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err ) 2>&1 | grep real
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
run c5_surface_1024_c2 --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 2
run c5_surface_1024_c8 --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 8
run c5_surface_1024_c1 --config c5 --steps 4 --warmup 3 --soft-soft surface --chunks 1
run c4_surface_512 --config c4 --scenes 512 --steps 5 --warmup 3 --soft-soft surface
run c4_volume_512 --config c4 --scenes 512 --steps 5 --warmup 3
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tail -1
