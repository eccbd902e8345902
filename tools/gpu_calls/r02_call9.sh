#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python tools/gpu_crp_parity.py crp_photo > gpurun_out/crp_parity.log 2>&1; cat gpurun_out/crp_parity.log
