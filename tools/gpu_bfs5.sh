#!/bin/bash
# The surface intersector on deforming meshes + compute-sanitizer memcheck over its small cases.   tools/gpu_bfs5.sh <tag>
tag=${1:-bfs5}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q -k deforming 2>&1 | tail -30 ) > gpurun_out/${tag}_deforming.log 2>&1
tail -5 gpurun_out/${tag}_deforming.log
( timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_surface_bfs.py -m gpu -x -q \
    -k "reference_fixture or multi_chunk or deforming or fully_penetrated" 2>&1 | tail -25 ) > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/${tag}_memcheck.log
