#!/bin/bash
# round 2, final 2-GPU call: mgpu_check with the final build (perturbation tiled kernel with the peer stores fused), box and
# 3-D ini configuration on two slabs.
mkdir -p gpurun_out
O=gpurun_out
export LBM_PEER_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 tests/mgpu_check.py > $O/h_mgpu_check_p2.log 2>&1 ); grep -E "MGPU|False" $O/h_mgpu_check_p2.log || tail -20 $O/h_mgpu_check_p2.log
( timeout 300 $TR --master-port 29550 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $O/h_box_p2.json 2> $O/h_box_p2.err ); python scripts/bench_brief.py $O/h_box_p2.json || tail -5 $O/h_box_p2.err
for F in 0 512; do
( timeout 300 $TR --master-port 2956$((F/512)) bench.py --gpus 2 --workload ini3d --size 256 --steps 30 --warmup 5 --flags $F > $O/h_ini3d_p2_flags$F.json 2> $O/h_ini3d_p2_flags$F.err ); python scripts/bench_brief.py $O/h_ini3d_p2_flags$F.json || tail -5 $O/h_ini3d_p2_flags$F.err
done
( timeout 100 python bench.py --workload ini3d --size 256 --steps 30 --warmup 5 > $O/h_ini3d_p1.json 2> $O/h_ini3d_p1.err ); python scripts/bench_brief.py $O/h_ini3d_p1.json | tail -1
