#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_ab.py r0 default 592 > gpurun_out/ab5_592.log 2>&1; cat gpurun_out/ab5_592.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
timeout 700 python bench.py --warmup 3 --steps 1 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json; tail -8 gpurun_out/bench5.err
