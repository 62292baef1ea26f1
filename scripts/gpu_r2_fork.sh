#!/bin/bash
# round 2, 1-GPU call: config 2 with the open-row chain on a second stream / parallel graph branch beside the density tile
# (LBM_OPEN_FORK=1) against the serial order; parity of the 2-D colour-gradient tier with it.
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu"
for F in 0 1 0 1; do
  ( LBM_OPEN_FORK=$F timeout 200 python bench.py --workload cfg2 $B > $O/s6_cfg2_fork$F.json 2> $O/s6_cfg2_fork$F.err ); echo "cfg2 FORK=$F"; python scripts/bench_brief.py $O/s6_cfg2_fork$F.json 2>&1 | head -1; tail -2 $O/s6_cfg2_fork$F.err
done
( LBM_OPEN_FORK=1 timeout 400 python -u -m pytest tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py tests/test_gpu_classes.py -m gpu -q -x -k "cfg2 or d2q9 or D2Q9 or gold or trajectory or channel or classes or tile" > $O/s6_pytest_fork.log 2>&1; echo "rc=$?" >> $O/s6_pytest_fork.log ); tail -3 $O/s6_pytest_fork.log
