#!/bin/bash
# last 1-GPU call of round 1 (about 5 minutes of box time): the tests of everything added since the previous
# call first, then the whole GPU tier, then the porous bench line and the headline bench
mkdir -p gpurun_out
O=gpurun_out
( timeout 170 python -u -m pytest tests/test_gpu_cg.py tests/test_gpu_sc.py -m gpu -q -rf -k "open_boundaries or d3q19_vs_dense or d3q19_larger" > $O/last_new_tests.log 2>&1; echo "rc=$?" >> $O/last_new_tests.log )
tail -4 $O/last_new_tests.log
( timeout 420 python -u -m pytest tests -m gpu -q -rf --durations=8 > $O/last_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/last_pytest_gpu.log )
tail -15 $O/last_pytest_gpu.log
( timeout 120 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/last_bench_porous.json 2> $O/last_bench_porous.err; echo "rc=$?" >> $O/last_bench_porous.err )
tail -2 $O/last_bench_porous.err; python scripts/bench_brief.py $O/last_bench_porous.json
( timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu > $O/last_bench_512.json 2> $O/last_bench_512.err; echo "rc=$?" >> $O/last_bench_512.err )
tail -2 $O/last_bench_512.err; python scripts/bench_brief.py $O/last_bench_512.json
