#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -8
LBM_TEST_FLAGS=32 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/mgpu_check.py > gpurun_out/mgpu_check2_packed.log 2>&1
grep -E "bit-equal|MGPU" gpurun_out/mgpu_check2_packed.log || tail -20 gpurun_out/mgpu_check2_packed.log
