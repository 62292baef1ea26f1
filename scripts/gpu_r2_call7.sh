#!/bin/bash
# round 2, 1-GPU call: the collision pass with the normals two planes ahead (one barrier per plane step) against one plane ahead
# (two barriers), on the 512^3 box and the porous workload; GPU tier with the new default; reference Numba-CUDA rows.
mkdir -p gpurun_out
O=gpurun_out
for A in 1 2; do
  ( LBM_COLLIDE_AHEAD=$A timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $O/c7_box_ahead$A.json 2> $O/c7_box_ahead$A.err ); echo "box AHEAD=$A"; python scripts/bench_brief.py $O/c7_box_ahead$A.json || tail -3 $O/c7_box_ahead$A.err
  ( LBM_COLLIDE_AHEAD=$A timeout 150 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/c7_porous_ahead$A.json 2> $O/c7_porous_ahead$A.err ); echo "porous AHEAD=$A"; python scripts/bench_brief.py $O/c7_porous_ahead$A.json | head -4 || tail -3 $O/c7_porous_ahead$A.err
done
( timeout 600 python -u -m pytest tests -m gpu -q -x -rf > $O/c7_pytest.log 2>&1; echo "rc=$?" >> $O/c7_pytest.log ); tail -3 $O/c7_pytest.log
( timeout 400 python scripts/ref_numba_cuda.py 1 2 3 > $O/c7_reference_numba_cuda.jsonl 2> $O/c7_reference_numba_cuda.err ); cat $O/c7_reference_numba_cuda.jsonl | cut -c100-400
N="python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --size 256"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o $O/c7_prof_collide_ahead2 -f $N > $O/c7_ncu.log 2>&1 ); tail -1 $O/c7_ncu.log
N="python bench.py --workload porous --size 256 --nz 192 --steps 3 --warmup 1"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_density_tiled" -s 2 -c 1 -o $O/c7_prof_density_solids -f $N > $O/c7_ncu2.log 2>&1 ); tail -1 $O/c7_ncu2.log
