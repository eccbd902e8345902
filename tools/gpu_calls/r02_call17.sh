#!/bin/bash
# round 2, GPU call 17: the headline bench on 4 GPUs (torchrun, one rank per GPU, gather inside the timed region)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -6 gpurun_out/bench_n4.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_n4.json
