#!/bin/bash
# 1-GPU call: full GPU test tier, bench lines, launch list and --set full captures of the two fast-path kernels at 512^3
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -2 gpurun_out/bench_512.err; python scripts/bench_brief.py gpurun_out/bench_512.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
for fl in 0; do
  timeout 300 python bench.py --lattice 9 --size 512 --steps 2000 --warmup 20 --no-cpu --no-e2e --flags $fl > gpurun_out/bench_2d512_f$fl.json 2>> gpurun_out/bench_2d.err; python scripts/bench_brief.py gpurun_out/bench_2d512_f$fl.json
done
timeout 300 python bench.py --lattice 9 --size 1024 --steps 2000 --warmup 20 --no-cpu --no-e2e > gpurun_out/bench_2d1024.json 2>> gpurun_out/bench_2d.err; python scripts/bench_brief.py gpurun_out/bench_2d1024.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fast_512.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_density_tiled" -s 2 -c 1 -o gpurun_out/prof_density_512 -f python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k1.log 2>&1; tail -1 gpurun_out/ncu_k1.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"cg_collide_tiled" -s 2 -c 1 -o gpurun_out/prof_collide_512 -f python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
