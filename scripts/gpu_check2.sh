#!/bin/bash
# 2-GPU call: parity tests (incl. slab decomposition), multi-GPU check and 1/2-GPU benches
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu_check.log 2>&1
grep -E "bit-equal|MGPU" gpurun_out/mgpu_check.log
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_512_n1.json 2> gpurun_out/bench_512_n1.err; cat gpurun_out/bench_512_n1.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_512_n2.json 2> gpurun_out/bench_512_n2.err; tail -3 gpurun_out/bench_512_n2.err; cat gpurun_out/bench_512_n2.json | cut -c1-400
