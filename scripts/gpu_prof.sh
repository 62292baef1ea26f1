#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"PullDensity" -s 1 -c 1 -o gpurun_out/prof_k1 -f python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fast_256.csv python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
