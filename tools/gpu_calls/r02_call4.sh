#!/bin/bash
# round 2, GPU call 4: parity tests (new: componentwise RHS, 1 Myr fixture cells), G3 at full length on the engine,
# blended transfer with the gross-exchange band, first run of the new bench.py
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_g3_full.py > gpurun_out/g3_full_gpu.log 2>&1; cat gpurun_out/g3_full_gpu.log
timeout 600 python tools/gpu_band_probe.py 100000 0 1e-6 1e-4 1e-2 > gpurun_out/band_probe2.log 2>&1; cat gpurun_out/band_probe2.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
