#!/bin/bash
# First GPU call of the next round (one B200, ~8 GPU-minutes):
#   1. A/B the default library against the product-form build variant on a 592-cell stride
#      (same cells, kernel time, per-phase cycles, max |dlog10| between the two)
#   2. parity tests on the default library
#   3. the bench line (device-resident + e2e + bounded CPU baseline)
# If (1) shows the expected solve speed-up (27.6 k -> ~7 k cycles per call) and <= 1e-3 dex difference,
# make the product form the default (drop the #ifdef in engine_core.cuh / engine_la.cuh, rebuild, re-run 2).
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_ab.py default default_pf 592 > gpurun_out/ab_pf_592.log 2>&1; cat gpurun_out/ab_pf_592.log
timeout 200 python tools/gpu_ab.py default_pf default_pfo 592 > gpurun_out/ab_pfo_592.log 2>&1; cat gpurun_out/ab_pfo_592.log
timeout 200 python tools/gpu_ab.py default default_vs 592 > gpurun_out/ab_vs_592.log 2>&1; cat gpurun_out/ab_vs_592.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
UCLGPU_TEST_SECOND_NETWORK=1 timeout 180 python -m pytest tests/test_gpu_second_network.py -m gpu -q -rxX > gpurun_out/pytest_second_network.log 2>&1; tail -3 gpurun_out/pytest_second_network.log
timeout 500 python bench.py --warmup 3 --steps 1 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -8 gpurun_out/bench.err
