#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py > gpurun_out/mgpu_check2.log 2>&1
grep -E "bit-equal|MGPU" gpurun_out/mgpu_check2.log || tail -20 gpurun_out/mgpu_check2.log
p=29560
for fl in 0 32 16; do
  p=$((p+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e --flags $fl > gpurun_out/scale2_f$fl.json 2> gpurun_out/scale2_f$fl.err
  python scripts/bench_brief.py gpurun_out/scale2_f$fl.json || tail -5 gpurun_out/scale2_f$fl.err
done
