#!/bin/bash
# round 2, 1-GPU call: programmatic dependent launch (LBM_PDL=1: griddepcontrol in every one-thread-per-node kernel and the D2Q9
# tile kernels, launches with cudaLaunchAttributeProgrammaticStreamSerialization) against plain launches on the launch-bound 2-D
# configurations; parity of the 2-D tiers with it; then the new defaults (split open-row chain, perturbation prefetch).
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu"
for P in 0 1; do
  for W in cfg1 cfg2 cfg3; do
    ( LBM_PDL=$P timeout 200 python bench.py --workload $W $B > $O/s5_${W}_pdl$P.json 2> $O/s5_${W}_pdl$P.err ); echo "$W PDL=$P"; python scripts/bench_brief.py $O/s5_${W}_pdl$P.json 2>&1 | head -1; tail -2 $O/s5_${W}_pdl$P.err
  done
done
( LBM_PDL=1 timeout 400 python -u -m pytest tests/test_gpu_sc.py tests/test_gpu_cg.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py -m gpu -q -x -k "sc or cfg1 or cfg2 or cfg3 or trajectory or chunked or d2q9 or gold or indexing or tracer" > $O/s5_pytest_pdl.log 2>&1; echo "rc=$?" >> $O/s5_pytest_pdl.log ); tail -3 $O/s5_pytest_pdl.log
( LBM_PDL=1 timeout 200 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu --no-e2e > $O/s5_box256_pdl1.json 2> $O/s5_box256_pdl1.err ); echo "box 256 PDL=1"; python scripts/bench_brief.py $O/s5_box256_pdl1.json 2>&1 | head -1
