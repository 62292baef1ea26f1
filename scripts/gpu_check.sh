#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python bench.py --size 256 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; tail -2 gpurun_out/bench_256.err; python scripts/bench_brief.py gpurun_out/bench_256.json
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -2 gpurun_out/bench_512.err; python scripts/bench_brief.py gpurun_out/bench_512.json
