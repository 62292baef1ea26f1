#!/bin/bash
# round 2, 1-GPU call: GPU tier after the class / solid-normal pipelining of the solids collision kernel and the compact
# wetting-solid list; porous + box lines; ncu of the solids kernel; numba-cuda probe (the unmodified reference kernels on the B200);
# BASELINE configurations 1-3 with the reference on the host cores beside them.
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -u -m pytest tests -m gpu -q -x -rf > $O/c6_pytest.log 2>&1; echo "rc=$?" >> $O/c6_pytest.log ); tail -4 $O/c6_pytest.log
( timeout 150 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/c6_porous.json 2> $O/c6_porous.err ); python scripts/bench_brief.py $O/c6_porous.json || tail -3 $O/c6_porous.err
( timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $O/c6_box.json 2> $O/c6_box.err ); python scripts/bench_brief.py $O/c6_box.json || tail -3 $O/c6_box.err
( timeout 120 python -c "
from numba import cuda
print('numba.cuda available:', cuda.is_available())
cuda.detect()
" > $O/c6_numba_probe.log 2>&1 ); tail -8 $O/c6_numba_probe.log
( timeout 300 python -c "
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import ref_numba
out = ref_numba.run_cg2d(512, 512, 300, target='cuda')
dt = out['step_seconds'][-200:]
print('reference Numba-CUDA on this GPU, cfg 2 (512^2): %.3f ms/step = %.1f MLUPS (%d void nodes)' % (dt.mean()*1e3, out['n_fluid']/dt.mean()/1e6, out['n_fluid']))
" > $O/c6_ref_cuda_cfg2.log 2>&1 ); tail -4 $O/c6_ref_cuda_cfg2.log
for w in cfg1 cfg2 cfg3; do
  ( timeout 400 python bench.py --workload $w --steps 4000 --warmup 100 > $O/c6_bench_$w.json 2> $O/c6_bench_$w.err ); python scripts/bench_brief.py $O/c6_bench_$w.json || tail -3 $O/c6_bench_$w.err
  python -c "
import json; d=json.load(open('$O/c6_bench_$w.json')); print('   cpu_baseline', d['cpu_baseline'])" 2>/dev/null | cut -c1-300
done
N="python bench.py --workload porous --size 256 --nz 192 --steps 3 --warmup 1"
( timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o $O/c6_prof_collide_solids -f $N > $O/c6_ncu_cs.log 2>&1 ); tail -1 $O/c6_ncu_cs.log
