#!/bin/bash
# tools/gpu_bfs6.sh <tag>: the topology and dense-clutter cases of the surface intersector.
tag=${1:-bfs6}
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q -s -k "topology or clutter" 2>&1 | tail -30 ) > gpurun_out/${tag}_tests.log 2>&1
tail -12 gpurun_out/${tag}_tests.log
