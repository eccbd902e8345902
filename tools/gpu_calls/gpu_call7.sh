#!/bin/bash
# round-1 measurement call: bench line, launch list + DRAM traffic of the same command, full ncu capture, reference arm
set -x
mkdir -p gpurun_out
timeout 500 python bench.py --warmup 3 --steps 1 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; cat gpurun_out/bench7.json; tail -6 gpurun_out/bench7.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_bench.csv python bench.py --warmup 3 --steps 1 > gpurun_out/bench7_under_ncu.log 2>&1; tail -30 gpurun_out/launches_bench.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/prof3_k_integrate -f python tools/ncu_target.py 148 1e2 > gpurun_out/ncu_full3.log 2>&1; tail -3 gpurun_out/ncu_full3.log
timeout 300 python bench.py --impl reference --warmup 1 --steps 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err; cat gpurun_out/bench7_ref.json
