#!/bin/bash
# round 2, final P-GPU call (P = $1): mgpu_check with the default (one-sided) exchange, the 512^3 box (with the host-buffer leg),
# the porous workload; P = 8: BASELINE config 5 at its own size and the reference's 3-D ini configuration at 512^3.
mkdir -p gpurun_out
O=gpurun_out
P=${1:-4}
export LBM_PEER_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/f_mgpu_check_p$P.log 2>&1 ); grep -E "MGPU|False" $O/f_mgpu_check_p$P.log || tail -20 $O/f_mgpu_check_p$P.log
( timeout 300 $TR --master-port 29550 bench.py --gpus $P --steps 20 --warmup 5 --no-cpu > $O/f_box_p${P}.json 2> $O/f_box_p${P}.err ); python scripts/bench_brief.py $O/f_box_p${P}.json || tail -5 $O/f_box_p${P}.err
( timeout 200 $TR --master-port 29560 bench.py --gpus $P --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/f_porous_p${P}.json 2> $O/f_porous_p${P}.err ); python scripts/bench_brief.py $O/f_porous_p${P}.json || tail -5 $O/f_porous_p${P}.err
if [ "$P" = "8" ]; then
  ( timeout 400 $TR --master-port 29561 bench.py --gpus 8 --workload porous --size 512 --nz 1024 --steps 30 --warmup 5 > $O/f_cfg5_512x512x1024_p8.json 2> $O/f_cfg5_512x512x1024_p8.err ); python scripts/bench_brief.py $O/f_cfg5_512x512x1024_p8.json || tail -5 $O/f_cfg5_512x512x1024_p8.err
  ( timeout 300 $TR --master-port 29562 bench.py --gpus 8 --workload ini3d --size 512 --steps 30 --warmup 5 > $O/f_ini3d_512_p8.json 2> $O/f_ini3d_512_p8.err ); python scripts/bench_brief.py $O/f_ini3d_512_p8.json || tail -5 $O/f_ini3d_512_p8.err
fi
