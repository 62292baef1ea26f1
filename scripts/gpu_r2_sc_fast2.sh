#!/bin/bash
# round 2, 1-GPU call: (1) configs 1-3 at 200 steps (the 2-D lattices replay a CUDA graph; short runs are dominated by the capture):
# two-pass Shan-Chen loops, D2Q9 collide tile with the treated open rows folded in (LBM_TILE_2D_ROWS=0: the two patch launches);
# (2) ncu --set full of the two passes of the explicit-forcing loop at config 3's size; (3) parity of what changed.
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu"
for W in cfg1 cfg3; do
  ( timeout 200 python bench.py --workload $W $B --general > $O/s2_${W}_general.json 2> $O/s2_${W}_general.err ); echo "$W general"; python scripts/bench_brief.py $O/s2_${W}_general.json 2>&1 | head -1
  ( timeout 200 python bench.py --workload $W $B > $O/s2_${W}.json 2> $O/s2_${W}.err ); echo "$W two-pass"; python scripts/bench_brief.py $O/s2_${W}.json 2>&1 | head -5
done
( LBM_TILE_2D_ROWS=0 timeout 200 python bench.py --workload cfg2 $B > $O/s2_cfg2_rows0.json 2> $O/s2_cfg2_rows0.err ); echo "cfg2 patch launches"; python scripts/bench_brief.py $O/s2_cfg2_rows0.json 2>&1 | head -1
( timeout 200 python bench.py --workload cfg2 $B > $O/s2_cfg2.json 2> $O/s2_cfg2.err ); echo "cfg2 rows folded"; python scripts/bench_brief.py $O/s2_cfg2.json 2>&1 | head -8
( timeout 300 python -u -m pytest tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py tests/test_gpu_classes.py -m gpu -q -x -k "cfg2 or d2q9 or D2Q9 or gold or trajectory or channel or classes or tile" > $O/s2_pytest.log 2>&1; echo "rc=$?" >> $O/s2_pytest.log ); tail -3 $O/s2_pytest.log
N="python bench.py --workload cfg3 --steps 12 --warmup 2 --no-cpu --flags 8"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"PullCollideOp|PullDensityOp" -s 6 -c 2 -o $O/s2_prof_efs_two_pass -f $N > $O/s2_ncu.log 2>&1 ); tail -2 $O/s2_ncu.log
