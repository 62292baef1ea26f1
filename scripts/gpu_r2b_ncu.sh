#!/bin/bash
# round 2, second session: ncu --set full of the two kernels this session changed most, final build -- the perturbation-model
# collision kernel with the prefetch slots (256^3) and the explicit-forcing pull-collide operator at config 3's size.
mkdir -p gpurun_out
O=gpurun_out
( timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cgp_collide_tiled" -s 2 -c 1 -o $O/n_prof_cgp_prefetch -f python bench.py --workload ini3d --steps 3 --warmup 1 --no-cpu --no-e2e > $O/n_ncu1.log 2>&1 ); tail -1 $O/n_ncu1.log
( timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EfsPullCollideOp|ScPullDensityOp" -s 6 -c 2 -o $O/n_prof_efs_final -f python bench.py --workload cfg3 --steps 12 --warmup 2 --no-cpu --flags 8 > $O/n_ncu2.log 2>&1 ); tail -1 $O/n_ncu2.log
