#!/bin/bash
# round 2, GPU call 20: the headline bench on 8 GPUs (torchrun, one rank per GPU, gather inside the timed region)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -4 gpurun_out/bench_n8.err | cut -c1-300; cut -c1-600 gpurun_out/bench_n8.json
