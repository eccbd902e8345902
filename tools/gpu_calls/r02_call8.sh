#!/bin/bash
# round 2, GPU call 8: kernel-level probes of the second network (where does it lose its step size?), config[2]
# with a budget that fits hot cores
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_second_network.py -m gpu -q -x > gpurun_out/pytest_second_network.log 2>&1; tail -25 gpurun_out/pytest_second_network.log
timeout 500 python bench.py --workload 3 --steps 1 --warmup 1 --cpu-seconds 45 > gpurun_out/bench_w3.json 2> gpurun_out/bench_w3.err; tail -4 gpurun_out/bench_w3.err; cut -c1-300 gpurun_out/bench_w3.json
