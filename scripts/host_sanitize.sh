#!/bin/bash
# The host test hook (the operator code of liblbmpm.so compiled with g++ -DLBM_HOSTCHECK) under AddressSanitizer +
# UndefinedBehaviorSanitizer: every index the one-thread-per-node operators, the emulated tiled kernels and the slab ring
# compute stays inside its allocation.  CPU only; rebuilds the normal hook afterwards.
set -e
cd "$(dirname "$0")/.."
CS=openlbmpm_b200/csrc
ASAN=$(gcc -print-file-name=libasan.so)
g++ -O1 -g -std=c++17 -DLBM_HOSTCHECK -ffp-contract=off -fPIC -shared -pthread -fsanitize=address,undefined \
    -fno-omit-frame-pointer -x c++ $CS/lbm_api.cu $CS/sc_api.cu $CS/tr_api.cu $CS/cg_fast.cu $CS/host_stubs.cu \
    -o tests/hostcheck/libhostcheck.so
status=0
for f in tests/test_hostcheck_*.py tests/test_host_classes.py tests/test_host_io.py; do
    LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python -m pytest "$f" -x -q -s 2>&1 \
        | grep -i "runtime error\|AddressSanitizer\|passed\|failed" | sort | uniq -c || status=1
done
python tests/hostcheck/build.py --force > /dev/null
exit $status
