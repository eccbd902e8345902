#!/bin/bash
# round 2, GPU call 6: ncu evidence for the round-2 kernel (launch list of the bench command with DRAM bytes, one
# full capture with source), then the bench workloads 1, 3, 4, 5 (configs[0], [2], [3], [4])
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 2 > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -5 gpurun_out/r02_launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/r02_prof \
    python tools/ncu_target.py 148 1e2 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof.ncu-rep --page source --csv > gpurun_out/r02_ncu_source.csv 2>/dev/null
ls -la gpurun_out/r02_*
timeout 600 python -m pytest tests -m gpu -q -k "collapse or jshock or cshock or edge" > gpurun_out/pytest_gpu_models.log 2>&1; tail -5 gpurun_out/pytest_gpu_models.log
for w in 1 5 4 3; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --cpu-seconds 45 > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  tail -4 gpurun_out/bench_w$w.err; cut -c1-400 gpurun_out/bench_w$w.json
done
