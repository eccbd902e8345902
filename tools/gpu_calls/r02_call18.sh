#!/bin/bash
# round 2, GPU call 18: loose-tolerance sweep on the hot cell, both networks
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_loose_sweep.py > gpurun_out/loose_sweep.log 2>&1; cut -c1-250 gpurun_out/loose_sweep.log | tail -20
