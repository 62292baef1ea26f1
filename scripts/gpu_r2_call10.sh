#!/bin/bash
# round 2, 1-GPU call: collision pass with the per-node inputs of the next plane prefetched into thread-private shared memory
# (LBM_POP_PREFETCH=1) against the plain loads, 512^3 and 256^3; parity of the prefetch variant against the C oracle at 128^3.
mkdir -p gpurun_out
O=gpurun_out
for PF in 0 1; do
  ( LBM_POP_PREFETCH=$PF timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $O/c10_box_pf$PF.json 2> $O/c10_box_pf$PF.err ); echo "box 512 PF=$PF"; python scripts/bench_brief.py $O/c10_box_pf$PF.json || tail -3 $O/c10_box_pf$PF.err
  ( LBM_POP_PREFETCH=$PF timeout 200 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu --no-e2e > $O/c10_box256_pf$PF.json 2> $O/c10_box256_pf$PF.err ); echo "box 256 PF=$PF"; python scripts/bench_brief.py $O/c10_box256_pf$PF.json | head -3 || tail -3 $O/c10_box256_pf$PF.err
done
( LBM_POP_PREFETCH=1 timeout 300 python -u -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_cg.py -m gpu -q -x -k "128_cubed or cfg4 or tiled" > $O/c10_pytest_pf.log 2>&1; echo "rc=$?" >> $O/c10_pytest_pf.log ); tail -3 $O/c10_pytest_pf.log
N="python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --size 256"
( LBM_POP_PREFETCH=1 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o $O/c10_prof_collide_pf -f $N > $O/c10_ncu.log 2>&1 ); tail -1 $O/c10_ncu.log
