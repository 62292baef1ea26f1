#!/bin/bash
# round 2, 2-GPU call: slab decomposition bit-equal to one GPU for every case of tests/mgpu_check.py (NCCL exchange, then the
# one-sided exchange LBM_FLAG_PEER_EXCHANGE = 256), then the 512^3 box and the porous workload on two slabs with both exchanges.
mkdir -p gpurun_out
O=gpurun_out
P=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/r2_mgpu_check_p$P.log 2>&1 ); grep -E "bit-equal|MGPU" $O/r2_mgpu_check_p$P.log || tail -20 $O/r2_mgpu_check_p$P.log
( LBM_TEST_FLAGS=256 LBM_PEER_TIMEOUT_MS=8000 timeout 300 $TR --master-port 29543 tests/mgpu_check.py > $O/r2_mgpu_check_p${P}_peer.log 2>&1 ); grep -E "bit-equal|MGPU" $O/r2_mgpu_check_p${P}_peer.log || tail -30 $O/r2_mgpu_check_p${P}_peer.log
for F in 0 256; do
  ( LBM_PEER_TIMEOUT_MS=8000 timeout 200 $TR --master-port 2955$((F / 256)) bench.py --gpus $P --steps 30 --warmup 5 --flags $F --no-cpu --no-e2e > $O/r2_box_p${P}_flags$F.json 2> $O/r2_box_p${P}_flags$F.err ); python scripts/bench_brief.py $O/r2_box_p${P}_flags$F.json || tail -5 $O/r2_box_p${P}_flags$F.err
  ( LBM_PEER_TIMEOUT_MS=8000 timeout 200 $TR --master-port 2956$((F / 256)) bench.py --gpus $P --workload porous --size 256 --nz 512 --steps 30 --warmup 5 --flags $F > $O/r2_porous_p${P}_flags$F.json 2> $O/r2_porous_p${P}_flags$F.err ); python scripts/bench_brief.py $O/r2_porous_p${P}_flags$F.json || tail -5 $O/r2_porous_p${P}_flags$F.err
done
grep -h -o '"checksum": [^]]*]' $O/r2_box_p${P}_flags*.json $O/r2_porous_p${P}_flags*.json
