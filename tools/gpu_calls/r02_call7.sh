#!/bin/bash
# round 2, GPU call 7: why did workload 5 (crp_photo library) hang?  Tight timeouts everywhere.
set -x
mkdir -p gpurun_out
timeout 60 python tools/gpu_crp_probe.py default_sub 4 1.0 > gpurun_out/probe_sub.log 2>&1; echo "rc=$?" >> gpurun_out/probe_sub.log; cat gpurun_out/probe_sub.log
timeout 60 python tools/gpu_crp_probe.py crp_photo 4 1.0 > gpurun_out/probe_crp.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/probe_crp.log; cat gpurun_out/probe_crp.log
if [ $rc -ne 0 ]; then
  timeout 240 compute-sanitizer --tool synccheck --print-limit 10 python tools/gpu_crp_probe.py crp_photo 1 1e-6 > gpurun_out/probe_crp_synccheck.log 2>&1; tail -30 gpurun_out/probe_crp_synccheck.log
else
  timeout 100 python tools/gpu_crp_probe.py crp_photo 296 1e6 > gpurun_out/probe_crp296.log 2>&1; echo "rc=$?" >> gpurun_out/probe_crp296.log; cat gpurun_out/probe_crp296.log
fi
timeout 200 python tools/gpu_ab.py default default_f4 592 > gpurun_out/ab_f4_592.log 2>&1; cat gpurun_out/ab_f4_592.log
timeout 100 python tools/gpu_ab.py default default_f2 592 > gpurun_out/ab_f2_592.log 2>&1; tail -4 gpurun_out/ab_f2_592.log
