#!/bin/bash
# 1-GPU call: the whole GPU test tier + smoke()
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -u -m pytest tests -m gpu -q -rf --durations=6 > $O/tier_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/tier_pytest_gpu.log ); tail -14 $O/tier_pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/tier_smoke.log 2>&1; echo "rc=$?" >> $O/tier_smoke.log ); tail -2 $O/tier_smoke.log
