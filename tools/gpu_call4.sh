#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_ab.py r0 default 592 > gpurun_out/ab4_592.log 2>&1; cat gpurun_out/ab4_592.log
timeout 600 python -m pytest tests -m gpu -x -q -k "not freefall" > gpurun_out/pytest_gpu4.log 2>&1; tail -3 gpurun_out/pytest_gpu4.log
timeout 900 python bench.py --warmup 1 --steps 1 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
