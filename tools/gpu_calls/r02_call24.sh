#!/bin/bash
# round 2, GPU call 24: N = 8 and N = 1 at --steps 4 with the neighbour-only cost hint (same conditions for the ratio)
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_nb_n8.json 2> gpurun_out/bench_nb_n8.err; tail -2 gpurun_out/bench_nb_n8.err | cut -c1-300; cut -c1-250 gpurun_out/bench_nb_n8.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_nb_n1_steps4.json 2> gpurun_out/bench_nb_n1_steps4.err; cut -c1-250 gpurun_out/bench_nb_n1_steps4.json
