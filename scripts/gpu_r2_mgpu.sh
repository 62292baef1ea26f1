#!/bin/bash
# round 2, P-GPU call (P = $1): slab decomposition bit-equal to one GPU for every case of tests/mgpu_check.py with the default
# exchange (one-sided stores into the neighbours' memory + flags) and with LBM_FLAG_NCCL_EXCHANGE (512), then the 512^3 box and
# the porous workload with both.  P = 8 adds BASELINE config 5 at its own size (512 x 512 x 1024).
mkdir -p gpurun_out
O=gpurun_out
P=${1:-2}
export LBM_PEER_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/r2_mgpu_check_p$P.log 2>&1 ); grep -E "MGPU|False" $O/r2_mgpu_check_p$P.log || tail -20 $O/r2_mgpu_check_p$P.log
( LBM_TEST_FLAGS=512 timeout 300 $TR --master-port 29543 tests/mgpu_check.py > $O/r2_mgpu_check_p${P}_nccl.log 2>&1 ); grep -E "MGPU|False" $O/r2_mgpu_check_p${P}_nccl.log || tail -30 $O/r2_mgpu_check_p${P}_nccl.log
for F in 0 512; do
  ( timeout 200 $TR --master-port 2955$((F / 512)) bench.py --gpus $P --steps 20 --warmup 5 --flags $F --no-cpu --no-e2e > $O/r2_box_p${P}_flags$F.json 2> $O/r2_box_p${P}_flags$F.err ); python scripts/bench_brief.py $O/r2_box_p${P}_flags$F.json || tail -5 $O/r2_box_p${P}_flags$F.err
done
( timeout 200 $TR --master-port 29560 bench.py --gpus $P --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/r2_porous_p${P}.json 2> $O/r2_porous_p${P}.err ); python scripts/bench_brief.py $O/r2_porous_p${P}.json || tail -5 $O/r2_porous_p${P}.err
if [ "$P" = "8" ]; then
  ( timeout 400 $TR --master-port 29561 bench.py --gpus 8 --workload porous --size 512 --nz 1024 --steps 30 --warmup 5 > $O/r2_cfg5_512x512x1024_p8.json 2> $O/r2_cfg5_512x512x1024_p8.err ); python scripts/bench_brief.py $O/r2_cfg5_512x512x1024_p8.json || tail -5 $O/r2_cfg5_512x512x1024_p8.err
  ( timeout 300 $TR --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_box_p8_full.json 2> $O/r2_box_p8_full.err ); python scripts/bench_brief.py $O/r2_box_p8_full.json || tail -5 $O/r2_box_p8_full.err
fi
