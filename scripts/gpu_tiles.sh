#!/bin/bash
set -x
mkdir -p gpurun_out
for zc in 32 64 128; do
  echo "== ZCHUNK=$zc"
  LBM_ZCHUNK=$zc timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_512_zc$zc.json 2>> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_512_zc$zc.json
done
LBM_ZCHUNK=64 timeout 300 python bench.py --size 256 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_256_zc64.json 2>> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_256_zc64.json
timeout 600 python -m pytest tests/test_gpu_cg.py -m gpu -q -k "tiled or d3q19 or slab" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -5
