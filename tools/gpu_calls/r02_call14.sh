#!/bin/bash
set -x
mkdir -p gpurun_out
for args in "0 5000 2" "0 200000 30" "1 5000 2" "1 200000 30"; do
  timeout 60 python tools/gpu_pp_probe.py $args >> gpurun_out/pp_probe.log 2>&1; echo "rc=$? ($args)" >> gpurun_out/pp_probe.log
done
cat gpurun_out/pp_probe.log
timeout 300 python -m pytest tests/test_gpu_gar_network.py tests/test_gpu_second_network.py -m gpu -q > gpurun_out/pytest_gpu_networks.log 2>&1; tail -12 gpurun_out/pytest_gpu_networks.log | cut -c1-300
