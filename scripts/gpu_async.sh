#!/bin/bash
# last seconds of the round-1 GPU budget: the asynchronous-output test, then its measured effect
mkdir -p gpurun_out
O=gpurun_out
( timeout 40 python -u -m pytest tests/test_gpu_classes.py -m gpu -q -rf -k "asynchronous" > $O/async_test.log 2>&1; echo "rc=$?" >> $O/async_test.log ); tail -3 $O/async_test.log
( timeout 40 python scripts/bench_output.py --size 256 --steps 60 --interval 10 > $O/async_bench.json 2> $O/async_bench.err ); cat $O/async_bench.json; tail -2 $O/async_bench.err
