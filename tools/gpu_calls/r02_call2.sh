#!/bin/bash
# round 2, GPU call 2: blocked Gauss-Jordan A/B, parity tests, stall probe (trace + state dump of stalled cells)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/gpu_ab.py default_r2a default 592 > gpurun_out/ab_gj_592.log 2>&1; cat gpurun_out/ab_gj_592.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python tools/gpu_stall_probe.py 30000 3 default > gpurun_out/stall_probe.log 2>&1; cat gpurun_out/stall_probe.log
