#!/bin/bash
# round 2, final 1-GPU call: GPU tier, smoke, the default bench line (with the CPU leg and the host-buffer leg), the reference arm,
# every other workload once, the launch list and one --set full capture of the two kernels of the metric step at 512^3.
mkdir -p gpurun_out
O=gpurun_out
( timeout 700 python -u -m pytest tests -m gpu -q -rf > $O/g_pytest.log 2>&1; echo "rc=$?" >> $O/g_pytest.log ); tail -4 $O/g_pytest.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.log 2>&1; echo "rc=$?" >> $O/g_smoke.log ); tail -2 $O/g_smoke.log
( timeout 400 python bench.py --steps 20 --warmup 5 > $O/g_bench_default.json 2> $O/g_bench_default.err ); python scripts/bench_brief.py $O/g_bench_default.json || tail -5 $O/g_bench_default.err
( timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/g_bench_reference.json 2> $O/g_bench_reference.err ); cut -c1-400 $O/g_bench_reference.json
( timeout 200 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/g_porous.json 2> $O/g_porous.err ); python scripts/bench_brief.py $O/g_porous.json | head -4
( timeout 200 python bench.py --workload ini3d --steps 30 --warmup 5 > $O/g_ini3d.json 2> $O/g_ini3d.err ); python scripts/bench_brief.py $O/g_ini3d.json | head -4
( timeout 200 python bench.py --workload cfg4 --steps 50 --warmup 5 --no-cpu > $O/g_cfg4.json 2> $O/g_cfg4.err ); python scripts/bench_brief.py $O/g_cfg4.json | head -3
for w in cfg1 cfg2 cfg3; do
  ( timeout 300 python bench.py --workload $w --steps 4000 --warmup 100 > $O/g_$w.json 2> $O/g_$w.err ); python scripts/bench_brief.py $O/g_$w.json | head -4
done
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/g_launches_512.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > $O/g_ncu_launches.log 2>&1 ); tail -1 $O/g_ncu_launches.log | cut -c1-200
( timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_(collide|density)_tiled" -s 4 -c 2 -o $O/g_prof_step_512 -f python bench.py --steps 4 --warmup 1 --no-cpu --no-e2e > $O/g_ncu_full.log 2>&1 ); tail -1 $O/g_ncu_full.log
