#!/bin/bash
# round-1 GPU call: parity tests, per-phase probe, launch list, one full ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=0 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_grid_probe.py 592 > gpurun_out/probe592.log 2>&1; tail -20 gpurun_out/probe592.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/launches_ncu_target.csv python tools/ncu_target.py 148 1e4 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/prof_k_integrate -f python tools/ncu_target.py 148 1e3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
