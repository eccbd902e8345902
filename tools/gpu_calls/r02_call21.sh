#!/bin/bash
# round 2, GPU call 21: the reference's own crp-photo test cell (5 Myr) at its loose tolerances and tighter, engine and oracle
set -x
mkdir -p gpurun_out
timeout 900 python tools/gpu_ref_test_cell.py > gpurun_out/ref_test_cell.log 2>&1; cut -c1-420 gpurun_out/ref_test_cell.log | tail -16
