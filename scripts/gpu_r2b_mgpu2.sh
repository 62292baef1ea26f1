#!/bin/bash
# round 2, second session: mgpu_check on 2 GPUs after the D2Q9 tile path takes phi of the materialised planes from the head operator
# in every schedule (the forked single-slab schedule and the slab schedule had differed by one ulp on the planes a convective outlet reads)
mkdir -p gpurun_out
O=gpurun_out
export LBM_PEER_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/f2_mgpu_check_p2.log 2>&1 ); grep -E "MGPU|False" $O/f2_mgpu_check_p2.log || tail -20 $O/f2_mgpu_check_p2.log
