#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sc.py tests/test_gpu_cg.py -m gpu -q -k "not graph_replay" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -5
for ty in 4 16 8; do
  echo "== TILE_Y=$ty"
  LBM_TILE_Y=$ty timeout 300 python bench.py --size 256 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_256_ty$ty.json 2> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_256_ty$ty.json
  LBM_TILE_Y=$ty timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_512_ty$ty.json 2>> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_512_ty$ty.json
done
