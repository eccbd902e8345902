#!/bin/bash
# round 2, GPU call 19: trace of the hot cell at reltol 0.97e-5 (the harness on the CPU sails through, the device does not)
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_loose_trace.py default 1e7 100 1e3 1e3 0.97e-5 1e-15 60000 > gpurun_out/loose_default_097.log 2>&1; tail -3 gpurun_out/loose_default_097.log | cut -c1-400
