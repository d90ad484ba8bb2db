#!/bin/bash
# Quick GPU visit: parity tests + one short bench (no CPU baseline).
tag=${1:-q}; shift
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
