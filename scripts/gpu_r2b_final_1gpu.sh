#!/bin/bash
# round 2, second session, final 1-GPU call: GPU tier, smoke, the default bench line (CPU leg + host-buffer leg), the reference arm,
# every other workload once with the final defaults, config 3 with / without the forked open-row chain, launch list of the metric step.
mkdir -p gpurun_out
O=gpurun_out
( timeout 700 python -u -m pytest tests -m gpu -q -rf > $O/f_pytest.log 2>&1; echo "rc=$?" >> $O/f_pytest.log ); tail -4 $O/f_pytest.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; echo "rc=$?" >> $O/f_smoke.log ); tail -2 $O/f_smoke.log
( timeout 400 python bench.py --steps 20 --warmup 5 > $O/f_bench_default.json 2> $O/f_bench_default.err ); python scripts/bench_brief.py $O/f_bench_default.json || tail -5 $O/f_bench_default.err
( timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/f_bench_reference.json 2> $O/f_bench_reference.err ); cut -c1-300 $O/f_bench_reference.json
( timeout 200 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/f_porous.json 2> $O/f_porous.err ); python scripts/bench_brief.py $O/f_porous.json | head -4
( timeout 200 python bench.py --workload ini3d --steps 30 --warmup 5 > $O/f_ini3d.json 2> $O/f_ini3d.err ); python scripts/bench_brief.py $O/f_ini3d.json | head -4
( timeout 200 python bench.py --workload cfg4 --steps 50 --warmup 5 --no-cpu > $O/f_cfg4.json 2> $O/f_cfg4.err ); python scripts/bench_brief.py $O/f_cfg4.json | head -3
for w in cfg1 cfg2 cfg3; do
  ( timeout 300 python bench.py --workload $w --steps 2000 --warmup 100 --no-cpu > $O/f_$w.json 2> $O/f_$w.err ); python scripts/bench_brief.py $O/f_$w.json | head -7
done
( LBM_SC_FORK=0 timeout 300 python bench.py --workload cfg3 --steps 2000 --warmup 100 --no-cpu > $O/f_cfg3_fork0.json 2> $O/f_cfg3_fork0.err ); echo "cfg3 LBM_SC_FORK=0"; python scripts/bench_brief.py $O/f_cfg3_fork0.json | head -1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f_launches_512.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > $O/f_ncu_launches.log 2>&1 ); tail -1 $O/f_ncu_launches.log | cut -c1-200
