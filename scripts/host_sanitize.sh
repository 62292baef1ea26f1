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
# ThreadSanitizer: the emulated CTAs of the tiled kernels (one host thread per CUDA thread, pthread barrier = __syncthreads),
# the cooperative persistent kernels (grid-wide barrier) and the slab ring (one host thread per rank).  Removing one barrier
# of cg_fast_persistent makes it report data races; with the barriers in place it reports none.
mkdir -p gpurun_out
TSAN=$(gcc -print-file-name=libtsan.so)
g++ -O1 -g -std=c++17 -DLBM_HOSTCHECK -ffp-contract=off -fPIC -shared -pthread -fsanitize=thread \
    -x c++ $CS/lbm_api.cu $CS/sc_api.cu $CS/tr_api.cu $CS/cg_fast.cu $CS/host_stubs.cu -o tests/hostcheck/libhostcheck.so
for f in tests/test_hostcheck_tiled.py tests/test_hostcheck_slabs.py tests/test_hostcheck_cg.py tests/test_hostcheck_sc.py; do
    LD_PRELOAD=$TSAN OPENBLAS_NUM_THREADS=1 TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" python -m pytest "$f" -q -s > gpurun_out/host_tsan.log 2>&1 || status=1
    echo "$f: ThreadSanitizer reports naming the hook: $(grep -c libhostcheck gpurun_out/host_tsan.log); $(tail -1 gpurun_out/host_tsan.log)"
    grep -q libhostcheck gpurun_out/host_tsan.log && status=1
done
# FMA contraction on (the closest host stand-in for nvcc's -fmad=true): the tolerances of the oracle / golden comparisons must hold
g++ -O2 -std=c++17 -DLBM_HOSTCHECK -DLBM_HOST_FMA_TEST -march=native -mfma -ffp-contract=fast -fPIC -shared -pthread \
    -x c++ $CS/lbm_api.cu $CS/sc_api.cu $CS/tr_api.cu $CS/cg_fast.cu $CS/host_stubs.cu -o tests/hostcheck/libhostcheck.so
python -m pytest tests/test_hostcheck_cg.py tests/test_hostcheck_sc.py tests/test_hostcheck_tiled.py tests/test_hostcheck_properties.py \
    tests/test_host_classes.py -q 2>&1 | tail -1 || status=1
python tests/hostcheck/build.py --force > /dev/null
exit $status
