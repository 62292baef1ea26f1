#!/bin/bash
# 2-GPU call of the next round: slab decomposition on the hardware for the cases added in round 1 after the GPU budget ran
# out (open channels, Shan-Chen models; tests/mgpu_check.py), then the porous workload on two slabs.
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > $O/n_mgpu2.log 2>&1 ); grep -E "bit-equal|MGPU" $O/n_mgpu2.log || tail -20 $O/n_mgpu2.log
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload porous --size 256 --nz 512 --steps 30 --warmup 5 > $O/n_porous_2gpu.json 2> $O/n_porous_2gpu.err ); python scripts/bench_brief.py $O/n_porous_2gpu.json || tail -5 $O/n_porous_2gpu.err
# the one-sided exchange (LBM_FLAG_PEER_EXCHANGE = 256): first run on the hardware.  Bit-equality first (its own short timeout: a
# flag that never arrives would spin), then the 512^3 box on two slabs with and without it.
( LBM_TEST_FLAGS=256 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/mgpu_check.py > $O/n_mgpu2_peer.log 2>&1 ); grep -E "bit-equal|MGPU" $O/n_mgpu2_peer.log || tail -20 $O/n_mgpu2_peer.log
for F in 0 256; do
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((4 + F / 256)) bench.py --gpus 2 --steps 30 --warmup 5 --flags $F --no-cpu --no-e2e > $O/n_box_2gpu_flags$F.json 2> $O/n_box_2gpu_flags$F.err ); python scripts/bench_brief.py $O/n_box_2gpu_flags$F.json || tail -5 $O/n_box_2gpu_flags$F.err
done
