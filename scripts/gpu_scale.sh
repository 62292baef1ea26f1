#!/bin/bash
# 8-GPU call: slab-decomposition bit-equality at P=8 and the 1/2/4/8 strong-scaling benches of the 512^3 box
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py > gpurun_out/mgpu_check8.log 2>&1
grep -E "bit-equal|MGPU" gpurun_out/mgpu_check8.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 50 --warmup 5 --no-e2e > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  python scripts/bench_brief.py gpurun_out/scale_n$n.json || tail -5 gpurun_out/scale_n$n.err
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
python scripts/bench_brief.py gpurun_out/scale_n1.json
