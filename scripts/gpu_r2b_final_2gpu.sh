#!/bin/bash
# round 2, second session, final 2-GPU call: mgpu_check (bit-equality of slabs with one GPU: every model, now with the two-pass
# Shan-Chen loops on slabs over NCCL) and the 512^3 box on two slabs with the final build.
mkdir -p gpurun_out
O=gpurun_out
export LBM_PEER_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/f_mgpu_check_p2.log 2>&1 ); grep -E "MGPU|False" $O/f_mgpu_check_p2.log || tail -20 $O/f_mgpu_check_p2.log
( timeout 300 $TR --master-port 29550 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $O/f_box_p2.json 2> $O/f_box_p2.err ); python scripts/bench_brief.py $O/f_box_p2.json || tail -5 $O/f_box_p2.err
