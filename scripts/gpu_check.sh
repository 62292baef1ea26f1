#!/bin/bash
# one gpurun call: GPU parity tests, smoke, a short bench and an ncu launch list (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --size 256 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; tail -2 gpurun_out/bench_256.err; cat gpurun_out/bench_256.json
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -2 gpurun_out/bench_512.err; cat gpurun_out/bench_512.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_256.csv python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_collide_tiled|PullDensity" -s 2 -c 2 -o gpurun_out/prof_fast -f python bench.py --size 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
