#!/bin/bash
# round 2, GPU call 25 (the last 2 GPU-minutes): two steps of the final bench.py -- step 0 unhinted, step 1 with the
# within-pass neighbourhood hint -- with per-step kernel times on stderr
mkdir -p gpurun_out
UCLCHEM_BENCH_VERBOSE=1 timeout 125 python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/bench_final_policy_2steps.json 2> gpurun_out/bench_final_policy_2steps.err
grep "step\|value" gpurun_out/bench_final_policy_2steps.err | cut -c1-200; cut -c1-200 gpurun_out/bench_final_policy_2steps.json
