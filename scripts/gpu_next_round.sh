#!/bin/bash
# First GPU call of the next round (1 GPU, about 6 minutes of box time): everything that was added after the round-1 GPU
# budget ran out, in the order DESIGN.md section 10 names.  Results land in gpurun_out/n_*.
mkdir -p gpurun_out
O=gpurun_out
( timeout 700 python -u -m pytest tests -m gpu -q -rf --durations=8 > $O/n_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/n_pytest_gpu.log ); tail -12 $O/n_pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/n_smoke.log 2>&1; echo "rc=$?" >> $O/n_smoke.log ); tail -2 $O/n_smoke.log
for w in cfg1 cfg2 cfg3; do
  ( timeout 120 python bench.py --workload $w --steps 2000 --warmup 50 > $O/n_bench_$w.json 2> $O/n_bench_$w.err ); python scripts/bench_brief.py $O/n_bench_$w.json || tail -3 $O/n_bench_$w.err
  ( timeout 120 python bench.py --workload $w --steps 2000 --warmup 50 --flags 128 > $O/n_bench_${w}_persistent.json 2>> $O/n_bench_$w.err ); python scripts/bench_brief.py $O/n_bench_${w}_persistent.json | head -1
  ( timeout 120 python bench.py --workload $w --steps 2000 --warmup 50 --flags 64 > $O/n_bench_${w}_ghostplanes.json 2>> $O/n_bench_$w.err ); python scripts/bench_brief.py $O/n_bench_${w}_ghostplanes.json | head -1
done
( timeout 150 python bench.py --workload cfg4 --steps 100 --warmup 10 > $O/n_bench_cfg4.json 2> $O/n_bench_cfg4.err ); python scripts/bench_brief.py $O/n_bench_cfg4.json
( timeout 150 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/n_bench_porous.json 2> $O/n_bench_porous.err ); python scripts/bench_brief.py $O/n_bench_porous.json
( timeout 200 python bench.py --steps 30 --warmup 5 > $O/n_bench_512.json 2> $O/n_bench_512.err ); python scripts/bench_brief.py $O/n_bench_512.json
N="python bench.py --workload porous --size 256 --nz 192 --steps 3 --warmup 1"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o $O/n_prof_collide_solids -f $N > $O/n_ncu_cs.log 2>&1 ); tail -1 $O/n_ncu_cs.log
