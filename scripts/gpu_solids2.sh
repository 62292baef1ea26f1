#!/bin/bash
# 1-GPU call: porous bench after the fence fix of the tiled collision kernel, the GPU tests that exercise the tiled
# kernels, and a short run of the headline box (all-fluid variants must be unchanged)
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5"
( timeout 100 $B > $O/sol2_default.json 2> $O/sol2_default.err ); python scripts/bench_brief.py $O/sol2_default.json
( timeout 150 python -u -m pytest tests/test_gpu_cg.py -m gpu -q -rf -k "tiled or open_boundaries or slab or sphere or d3q19" > $O/sol2_tests.log 2>&1; echo "rc=$?" >> $O/sol2_tests.log ); tail -4 $O/sol2_tests.log
( timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $O/sol2_box512.json 2> $O/sol2_box512.err ); python scripts/bench_brief.py $O/sol2_box512.json
