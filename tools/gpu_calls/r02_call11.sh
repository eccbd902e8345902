#!/bin/bash
# round 2, GPU call 11: the tests added after call 10 (G3 with a failed DVODE call, C-shock through 3 t_diss, overrides)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "g3_full or three_dissipation or overrides or disk_mode or model_api" > gpurun_out/pytest_gpu_new.log 2>&1; tail -30 gpurun_out/pytest_gpu_new.log | cut -c1-300
