#!/bin/bash
# round 2, GPU call 22 (final): whole GPU suite + smoke on the final build, the driver's bench command, the ncu launch
# list of the bench command and one full capture of k_integrate (evidence for the final build), N = 1 / 2 at --steps 8
# for a scaling table with the same hint coverage as the N = 4 / 8 runs
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -3 gpurun_out/bench_final_n1.err | cut -c1-300; cut -c1-300 gpurun_out/bench_final_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 2 > gpurun_out/r02f_bench_under_ncu.log 2>&1
tail -4 gpurun_out/r02f_launches_bench.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/r02f_prof \
    python tools/ncu_target.py 148 1e2 > gpurun_out/r02f_ncu_full.log 2>&1
tail -2 gpurun_out/r02f_ncu_full.log
ncu -i gpurun_out/r02f_prof.ncu-rep --page raw --csv > gpurun_out/r02f_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_prof.ncu-rep --page source --csv > gpurun_out/r02f_ncu_source.csv 2>/dev/null
rm -f gpurun_out/r02f_prof.ncu-rep
timeout 400 python bench.py --gpus 1 --steps 8 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_n1_steps8.json 2> gpurun_out/bench_n1_steps8.err; cut -c1-200 gpurun_out/bench_n1_steps8.json
