#!/bin/bash
# round 2, GPU call 5 (2 GPUs): the N = 2 bench path (torchrun, NCCL gather) and the in-process two-device sharding
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -8 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "collapse or edge or multi_device" > gpurun_out/pytest_gpu_n2.log 2>&1; tail -15 gpurun_out/pytest_gpu_n2.log
timeout 300 python tools/gpu_two_device.py > gpurun_out/two_device.log 2>&1; cat gpurun_out/two_device.log
