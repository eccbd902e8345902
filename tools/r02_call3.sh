#!/bin/bash
# round 2, GPU call 3: overlap on/off A/B after the blocked Gauss-Jordan; blended-transfer experiment on the stall cells
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_ab.py default default_nov 592 > gpurun_out/ab_nov_592.log 2>&1; cat gpurun_out/ab_nov_592.log
timeout 900 python tools/gpu_band_probe.py 100000 0 0.01 0.1 0.5 > gpurun_out/band_probe.log 2>&1; cat gpurun_out/band_probe.log
