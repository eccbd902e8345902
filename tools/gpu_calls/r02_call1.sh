#!/bin/bash
# round 2, GPU call 1: A/B the build variants, parity tests, second network, full-grid per-cell statistics
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/gpu_ab.py default default_pf 592 > gpurun_out/ab_pf_592.log 2>&1; cat gpurun_out/ab_pf_592.log
timeout 200 python tools/gpu_ab.py default_pf default_pfo 592 > gpurun_out/ab_pfo_592.log 2>&1; cat gpurun_out/ab_pfo_592.log
timeout 200 python tools/gpu_ab.py default default_vs 592 > gpurun_out/ab_vs_592.log 2>&1; cat gpurun_out/ab_vs_592.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
UCLGPU_TEST_SECOND_NETWORK=1 timeout 180 python -m pytest tests/test_gpu_second_network.py -m gpu -q -rxX > gpurun_out/pytest_second_network.log 2>&1; tail -5 gpurun_out/pytest_second_network.log
timeout 300 python tools/gpu_grid_full.py 100000 default > gpurun_out/grid_full.log 2>&1; cat gpurun_out/grid_full.log
