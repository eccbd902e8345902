#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_ab.py r0 default 592 > gpurun_out/ab3_592.log 2>&1; cat gpurun_out/ab3_592.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; tail -3 gpurun_out/pytest_gpu3.log
