#!/bin/bash
# round 2, GPU call 16: traces of the loose-tolerance cell (crp-photo and default network), whole GPU suite with the new tests
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_loose_trace.py crp_photo 1e7 100 1e3 1e3 1e-5 1e-15 > gpurun_out/loose_crp.log 2>&1; tail -5 gpurun_out/loose_crp.log | cut -c1-600
timeout 300 python tools/gpu_loose_trace.py default 1e7 100 1e3 1e3 1e-5 1e-15 > gpurun_out/loose_default.log 2>&1; tail -5 gpurun_out/loose_default.log | cut -c1-600
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
