#!/bin/bash
# 1-GPU call: pull-mask / fast-trig variants of the tiled kernels on the porous workload (A/B by environment), the GPU tests
# that exercise them, and ncu --set full captures of the two <solids> kernels
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5"
( timeout 100 $B > $O/sol_default.json 2> $O/sol_default.err ); python scripts/bench_brief.py $O/sol_default.json
( LBM_WETTING_EXACT_TRIG=1 timeout 100 $B > $O/sol_exact_trig.json 2> $O/sol_exact_trig.err ); python scripts/bench_brief.py $O/sol_exact_trig.json | head -3
( LBM_TILE_Y_COLLIDE=8 timeout 100 $B > $O/sol_ty8.json 2> $O/sol_ty8.err ); python scripts/bench_brief.py $O/sol_ty8.json | head -3
( timeout 150 python -u -m pytest tests/test_gpu_cg.py -m gpu -q -rf -k "tiled or open_boundaries or slab or sphere or d3q19" > $O/sol_tests.log 2>&1; echo "rc=$?" >> $O/sol_tests.log ); tail -4 $O/sol_tests.log
N="python bench.py --workload porous --size 256 --nz 192 --steps 3 --warmup 1"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o $O/prof_collide_solids -f $N > $O/ncu_cs.log 2>&1 ); tail -1 $O/ncu_cs.log
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_density_tiled" -s 2 -c 1 -o $O/prof_density_solids -f $N > $O/ncu_ds.log 2>&1 ); tail -1 $O/ncu_ds.log
