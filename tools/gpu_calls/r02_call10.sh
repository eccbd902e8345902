#!/bin/bash
# round 2, GPU call 10: full GPU suite on the final engine, config[4] at default tolerances, config[3] with its budget
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --workload 5 --steps 2 --warmup 1 --cpu-seconds 45 > gpurun_out/bench_w5.json 2> gpurun_out/bench_w5.err; tail -4 gpurun_out/bench_w5.err; cut -c1-300 gpurun_out/bench_w5.json
timeout 400 python bench.py --workload 4 --steps 2 --warmup 1 --cpu-seconds 45 > gpurun_out/bench_w4.json 2> gpurun_out/bench_w4.err; tail -4 gpurun_out/bench_w4.err; cut -c1-300 gpurun_out/bench_w4.json
