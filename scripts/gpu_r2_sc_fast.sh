#!/bin/bash
# round 2, 1-GPU call: two-pass form of the Shan-Chen loops (csrc/sc_fast.cuh) -- parity on the GPU (golden vectors in long
# calls, two-pass vs reference-ordered operators at the sizes of BASELINE configs 1 and 3), then configs 1 and 3 with the
# reference-ordered operators (--general) and with the two-pass form at 1 / 2 / 3 resident CTAs per SM of its collision pass.
mkdir -p gpurun_out
O=gpurun_out
( timeout 400 python -u -m pytest tests/test_gpu_sc.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py -m gpu -q -x -k "sc or cfg1 or cfg3 or Shan or shan or d3q19_vs or d3q19_open or d3q19_larger or trajectory or chunked" > $O/s1_pytest.log 2>&1; echo "rc=$?" >> $O/s1_pytest.log ); tail -4 $O/s1_pytest.log
for W in cfg1 cfg3; do
  ( timeout 200 python bench.py --workload $W --general --no-cpu > $O/s1_${W}_general.json 2> $O/s1_${W}_general.err ); echo "$W general"; python scripts/bench_brief.py $O/s1_${W}_general.json | head -5 || tail -3 $O/s1_${W}_general.err
  for OCC in 1 2 3; do
    ( LBM_SC_OCC=$OCC timeout 200 python bench.py --workload $W --no-cpu > $O/s1_${W}_occ$OCC.json 2> $O/s1_${W}_occ$OCC.err ); echo "$W two-pass OCC=$OCC"; python scripts/bench_brief.py $O/s1_${W}_occ$OCC.json | head -5 || tail -3 $O/s1_${W}_occ$OCC.err
  done
done
