#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -8
grep -E "bit-equal|MGPU" gpurun_out/pytest_gpu.log | tail
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"node_kernel" --kernel-name-base demangled -s 14 -c 1 -o gpurun_out/prof_k1 -f python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_collide_tiled" -s 1 -c 1 -o gpurun_out/prof_k2 -f python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k2.log
ncu -i gpurun_out/prof_k1.ncu-rep --page raw --csv 2>/dev/null | head -3 | cut -c1-300
