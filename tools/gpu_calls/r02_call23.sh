#!/bin/bash
# round 2, GPU call 23: headline bench after the cost hint was restricted to the zeta NEIGHBOURS (a cell's own earlier
# visit is never used): the driver's command line, and the no-hint figure next to it
set -x
mkdir -p gpurun_out
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_nb_n1.json 2> gpurun_out/bench_nb_n1.err; tail -3 gpurun_out/bench_nb_n1.err | cut -c1-300; cut -c1-300 gpurun_out/bench_nb_n1.json
