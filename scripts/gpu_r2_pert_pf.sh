#!/bin/bash
# round 2, 1-GPU call: (1) perturbation-model collision kernel with the next plane's per-node inputs prefetched into thread-private
# shared memory (LBM_PERT_PREFETCH=1) against the plain loads: the reference's 3-D ini configuration at 256^3, parity of the variant;
# (2) config 2 with the open-row pre-chain as one launch (one thread per column) or three parallel launches (LBM_OPEN_PRE_SPLIT).
mkdir -p gpurun_out
O=gpurun_out
for PF in 0 1; do
  ( LBM_PERT_PREFETCH=$PF timeout 200 python bench.py --workload ini3d --steps 30 --warmup 5 --no-cpu --no-e2e > $O/s4_ini3d_pf$PF.json 2> $O/s4_ini3d_pf$PF.err ); echo "ini3d PF=$PF"; python scripts/bench_brief.py $O/s4_ini3d_pf$PF.json 2>&1 | head -4
done
( LBM_PERT_PREFETCH=1 timeout 300 python -u -m pytest tests/test_gpu_cg.py tests/test_gpu_classes.py tests/test_gpu_baseline_sizes.py -m gpu -q -x -k "perturb or ini or Perturb" > $O/s4_pytest_pf.log 2>&1; echo "rc=$?" >> $O/s4_pytest_pf.log ); tail -3 $O/s4_pytest_pf.log
for SP in 0 1; do
  ( LBM_OPEN_PRE_SPLIT=$SP timeout 200 python bench.py --workload cfg2 --steps 200 --warmup 10 --no-cpu > $O/s4_cfg2_split$SP.json 2> $O/s4_cfg2_split$SP.err ); echo "cfg2 SPLIT=$SP"; python scripts/bench_brief.py $O/s4_cfg2_split$SP.json 2>&1 | head -7
done
( LBM_PERT_PREFETCH=1 timeout 200 python bench.py --workload ini3d --steps 30 --warmup 5 --no-cpu --no-e2e --nz 512 > $O/s4_ini3d_512_pf1.json 2> $O/s4_ini3d_512_pf1.err ); echo "ini3d 256x256x512 PF=1"; python scripts/bench_brief.py $O/s4_ini3d_512_pf1.json 2>&1 | head -3
( LBM_PERT_PREFETCH=0 timeout 200 python bench.py --workload ini3d --steps 30 --warmup 5 --no-cpu --no-e2e --nz 512 > $O/s4_ini3d_512_pf0.json 2> $O/s4_ini3d_512_pf0.err ); echo "ini3d 256x256x512 PF=0"; python scripts/bench_brief.py $O/s4_ini3d_512_pf0.json 2>&1 | head -3
