#!/bin/bash
# round 2, call 2 (1 GPU): the two tests that failed in call 1, the 2-D configurations timed the way the product runs them
# (graph replay, no per-launch events in the timed region) with and without the persistent kernel, the new bench line.
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -u -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_zz_persistent.py -m gpu -q -rf > $O/c2_pytest.log 2>&1; echo "rc=$?" >> $O/c2_pytest.log ); tail -5 $O/c2_pytest.log
for w in cfg1 cfg2 cfg3; do
  for F in 0 128 8; do
    ( timeout 120 python bench.py --workload $w --steps 4000 --warmup 100 --flags $F > $O/c2_bench_${w}_flags$F.json 2> $O/c2_bench_${w}_flags$F.err ); echo "$w flags $F"; python scripts/bench_brief.py $O/c2_bench_${w}_flags$F.json || tail -3 $O/c2_bench_${w}_flags$F.err
  done
done
( timeout 300 python bench.py --steps 20 --warmup 5 > $O/c2_bench_512.json 2> $O/c2_bench_512.err ); python scripts/bench_brief.py $O/c2_bench_512.json || tail -5 $O/c2_bench_512.err
( timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/c2_bench_ref.json 2> $O/c2_bench_ref.err ); cat $O/c2_bench_ref.json | cut -c1-600
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
