#!/bin/bash
# round 2, 1-GPU call: perturbation-operator model on the factored fast path (tiled density pass + cgp_collide_tiled_d3q19):
# GPU tier, the reference's 3-D ini configuration at 256^3 on the three kernel paths, ncu of the new kernel.
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -u -m pytest tests -m gpu -q -x -rf > $O/c9_pytest.log 2>&1; echo "rc=$?" >> $O/c9_pytest.log ); tail -5 $O/c9_pytest.log
for F in 0 2 1; do
  ( timeout 200 python bench.py --workload ini3d --size 256 --steps 30 --warmup 5 --flags $F > $O/c9_ini3d_flags$F.json 2> $O/c9_ini3d_flags$F.err ); echo "ini3d 256^3 flags $F"; python scripts/bench_brief.py $O/c9_ini3d_flags$F.json || tail -3 $O/c9_ini3d_flags$F.err
done
N="python bench.py --workload ini3d --size 256 --steps 3 --warmup 1"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cgp_collide_tiled" -s 2 -c 1 -o $O/c9_prof_cgp_collide -f $N > $O/c9_ncu.log 2>&1 ); tail -1 $O/c9_ncu.log
