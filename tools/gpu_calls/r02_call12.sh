#!/bin/bash
# round 2, GPU call 12: the driver's own bench command (runtime check, cost hint), reference arm
set -x
mkdir -p gpurun_out
date
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver_cmd.json 2> gpurun_out/bench_driver_cmd.err; tail -6 gpurun_out/bench_driver_cmd.err; cut -c1-300 gpurun_out/bench_driver_cmd.json
date
timeout 400 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; tail -3 gpurun_out/bench_reference_arm.err; cut -c1-300 gpurun_out/bench_reference_arm.json
date
