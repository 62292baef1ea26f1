#!/bin/bash
# round 2, 1-GPU call: D2Q9 tile kernels -- GPU tier, cfg 2 and the periodic 512^2 / 1024^2 boxes with and without them,
# porous after the deeper pull-mask pipeline.
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -u -m pytest tests -m gpu -q -x -rf > $O/c8_pytest.log 2>&1; echo "rc=$?" >> $O/c8_pytest.log ); tail -6 $O/c8_pytest.log
for T in 1 0; do
  ( LBM_TILE_2D=$T timeout 120 python bench.py --workload cfg2 --steps 4000 --warmup 100 --no-cpu > $O/c8_cfg2_tile$T.json 2> $O/c8_cfg2_tile$T.err ); echo "cfg2 LBM_TILE_2D=$T"; python scripts/bench_brief.py $O/c8_cfg2_tile$T.json || tail -3 $O/c8_cfg2_tile$T.err
  for n in 512 1024 2048; do
    ( LBM_TILE_2D=$T timeout 120 python bench.py --lattice 9 --size $n --steps 2000 --warmup 100 --no-cpu --no-e2e > $O/c8_d2q9_${n}_tile$T.json 2> $O/c8_d2q9_${n}_tile$T.err ); echo "D2Q9 periodic $n LBM_TILE_2D=$T"; python scripts/bench_brief.py $O/c8_d2q9_${n}_tile$T.json | head -4 || tail -3 $O/c8_d2q9_${n}_tile$T.err
  done
done
( timeout 150 python bench.py --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/c8_porous.json 2> $O/c8_porous.err ); python scripts/bench_brief.py $O/c8_porous.json | head -4 || tail -3 $O/c8_porous.err
