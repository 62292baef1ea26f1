#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 90 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import cases
cases.case_d3q19_periodic(None, n=(10, 16, 32), steps=4)
cases.case_d3q19_periodic(None, n=(70, 16, 64), steps=3)
cases.case_d3q19_sphere(None, n=(12, 16, 96), steps=4)
print('TMA quick ok')
" > gpurun_out/tma_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/tma_quick.log; tail -3 gpurun_out/tma_quick.log
if ! grep -q "TMA quick ok" gpurun_out/tma_quick.log; then echo "ABORT: quick TMA check failed"; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_cg.py tests/test_gpu_fullsize.py -m gpu -q -k "tiled or d3q19 or slab or cfg4" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -5
for tma in 1 0; do
  echo "== TMA=$tma (phi + scalar planes)"
  LBM_PHI_TMA=$tma LBM_SCALAR_TMA=$tma timeout 300 python bench.py --size 256 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_256_tma$tma.json 2>> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_256_tma$tma.json
  LBM_PHI_TMA=$tma LBM_SCALAR_TMA=$tma timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_512_tma$tma.json 2>> gpurun_out/bench_ty.err; python scripts/bench_brief.py gpurun_out/bench_512_tma$tma.json
done
