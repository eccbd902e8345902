#!/bin/bash
# A/B the round-start build (r0) against the current build, quick parity tests, ncu capture of the new kernel
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_ab.py r0 default 592 > gpurun_out/ab_592.log 2>&1; cat gpurun_out/ab_592.log
timeout 600 python -m pytest tests -m gpu -x -q -k "not freefall" > gpurun_out/pytest_gpu_quick.log 2>&1; tail -3 gpurun_out/pytest_gpu_quick.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/prof2_k_integrate -f python tools/ncu_target.py 148 1e2 > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
