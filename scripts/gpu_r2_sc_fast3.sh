#!/bin/bash
# round 2, 1-GPU call: two-pass Shan-Chen operators after the address-arithmetic rewrite (NbTable, both components requested at
# once in the density pass): configs 1 and 3 at 200 steps, the D3Q19 models two-pass vs reference-ordered, parity subset.
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 10 --no-cpu"
for W in cfg1 cfg3; do
  ( timeout 200 python bench.py --workload $W $B > $O/s3_${W}.json 2> $O/s3_${W}.err ); echo "$W two-pass"; python scripts/bench_brief.py $O/s3_${W}.json 2>&1 | head -5
done
( timeout 200 python scripts/sc3d_probe.py 128 40 > $O/s3_sc3d_probe.jsonl 2> $O/s3_sc3d_probe.err ); cat $O/s3_sc3d_probe.jsonl; tail -2 $O/s3_sc3d_probe.err
( timeout 300 python -u -m pytest tests/test_gpu_sc.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py -m gpu -q -x -k "sc or cfg1 or cfg3 or Shan or shan or d3q19_vs or d3q19_open or d3q19_larger or trajectory or chunked" > $O/s3_pytest.log 2>&1; echo "rc=$?" >> $O/s3_pytest.log ); tail -3 $O/s3_pytest.log
