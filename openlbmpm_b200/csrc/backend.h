// backend.h -- the few runtime services the library needs (device memory, copies, one-thread-
// per-node launches, an exclusive scan), so that the SAME operator code builds
//   * with nvcc for sm_100a  -> liblbmpm.so, the product, and
//   * with g++ (LBM_HOSTCHECK) -> tests/hostcheck/libhostcheck.so, a test hook that lets the
//     CPU-only test tier check the node arithmetic against the oracle without a GPU.
// The Python package only ever loads liblbmpm.so; there is no CPU fallback in the product.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <typeinfo>
#include <vector>

#ifndef LBM_HOSTCHECK
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#endif

namespace lbm {

#ifndef LBM_HOSTCHECK
// ------------------------------------------------------------------ CUDA backend
typedef cudaStream_t stream_t;

struct BackendError {
    std::string msg;
    bool oom = false;
};

#define LBM_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            throw ::lbm::BackendError{std::string(#expr) + ": " + cudaGetErrorString(_e)};    \
    } while (0)

inline void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw BackendError{"cudaMalloc of " + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e), e == cudaErrorMemoryAllocation};
    }
    return p;
}
inline void dev_free(void* p) {
    if (p) cudaFree(p);
}
inline void dev_zero(void* p, size_t bytes, stream_t s) { LBM_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, s)); }
inline void dev_h2d(void* d, const void* h, size_t bytes, stream_t s) {
    LBM_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
    LBM_CUDA_CHECK(cudaStreamSynchronize(s));
}
inline void dev_d2h(void* h, const void* d, size_t bytes, stream_t s) {
    LBM_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
    LBM_CUDA_CHECK(cudaStreamSynchronize(s));
}
inline void dev_d2d(void* d, const void* s_, size_t bytes, stream_t s) {
    LBM_CUDA_CHECK(cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s));
}
inline void dev_sync(stream_t s) { LBM_CUDA_CHECK(cudaStreamSynchronize(s)); }

// Programmatic dependent launch (sm_90+, LBM_PDL=1): a kernel launched with the attribute may have its CTAs scheduled while
// the kernel in front of it on the stream is still running -- as soon as every CTA of that one has executed
// `griddepcontrol.launch_dependents` -- and blocks in `griddepcontrol.wait` until that kernel has completed and its writes are
// visible.  Every kernel of this library that takes part executes both instructions before it touches memory, so the order of
// all memory operations is that of the plain stream; what overlaps is launch latency and CTA scheduling, which is what the
// launch-bound 2-D configurations (2-6 kernels of a few microseconds per step, replayed from a graph) spend their time on.
// Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("LBM_PDL"); return e ? atoi(e) != 0 : false; }();
    return on;
}
template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Args&... args) {
    if (!pdl_enabled()) {
        kernel<<<grid, block, smem, s>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (e != cudaSuccess) throw BackendError{std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e)};
}

template <class Op>
__global__ void __launch_bounds__(256) node_kernel(const Op op, const int64_t n) {
    pdl_prologue();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) op(i);
}

// one thread per item; counts launches for lbm_get_timing
extern thread_local int64_t g_launch_counter;

// optional per-launch CUDA-event timing on the launching stream (lbm_profile_enable / _report)
struct ProfRecord { const char* name; cudaEvent_t e0, e1; };
struct Profiler {
    bool on = false;
    std::vector<ProfRecord> recs;
    void begin(const char* name, stream_t s) {
        ProfRecord r; r.name = name;
        LBM_CUDA_CHECK(cudaEventCreate(&r.e0)); LBM_CUDA_CHECK(cudaEventCreate(&r.e1));
        LBM_CUDA_CHECK(cudaEventRecord(r.e0, s));
        recs.push_back(r);
    }
    void end(stream_t s) { LBM_CUDA_CHECK(cudaEventRecord(recs.back().e1, s)); }
    void clear() {
        for (auto& r : recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        recs.clear();
    }
};
extern thread_local Profiler g_prof;
inline bool g_prof_active() { return g_prof.on; }

template <class Op>
inline void launch(const Op& op, int64_t n, stream_t s) {
    if (n <= 0) return;
    const int block = 256;
    const int64_t grid = (n + block - 1) / block;
    if (g_prof.on) g_prof.begin(typeid(Op).name(), s);
    launch_kernel(node_kernel<Op>, dim3((unsigned)grid), dim3(block), 0, s, op, n);
    if (g_prof.on) g_prof.end(s);
    LBM_CUDA_CHECK(cudaGetLastError());
    ++g_launch_counter;
}

// ... with at least MINB resident 256-thread CTAs per SM (a register cap for bandwidth-bound operators that hold many values)
template <class Op, int MINB>
__global__ void __launch_bounds__(256, MINB) node_kernel_occ(const Op op, const int64_t n) {
    pdl_prologue();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) op(i);
}
template <int MINB, class Op>
inline void launch_occ(const Op& op, int64_t n, stream_t s) {
    if (n <= 0) return;
    const int block = 256;
    const int64_t grid = (n + block - 1) / block;
    if (g_prof.on) g_prof.begin(typeid(Op).name(), s);
    launch_kernel(node_kernel_occ<Op, MINB>, dim3((unsigned)grid), dim3(block), 0, s, op, n);
    if (g_prof.on) g_prof.end(s);
    LBM_CUDA_CHECK(cudaGetLastError());
    ++g_launch_counter;
}

// Replay of a launch-bound unit of work: small lattices (the 2-D configurations) spend their time in launch
// overhead, so `body` (which only enqueues kernels on `s`) is captured once into a CUDA graph and the graph is
// launched `reps` times.  The executable graph is parked in *keep and destroyed by the next call / the owner.
struct GraphKeep { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; };
inline void graph_release(GraphKeep* k, stream_t s) {
    if (!k->exec && !k->graph) return;
    cudaStreamSynchronize(s);
    if (k->exec) cudaGraphExecDestroy(k->exec);
    if (k->graph) cudaGraphDestroy(k->graph);
    k->exec = nullptr; k->graph = nullptr;
}
template <class F>
inline void replay(int reps, bool use_graph, GraphKeep* keep, stream_t s, F body) {
    if (reps <= 0) return;
    if (!use_graph || reps < 8 || g_prof.on) {
        for (int i = 0; i < reps; ++i) body();
        return;
    }
    graph_release(keep, s);
    const int64_t l0 = g_launch_counter;
    LBM_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
        body();
    } catch (...) {
        cudaGraph_t dead = nullptr;
        cudaStreamEndCapture(s, &dead);
        if (dead) cudaGraphDestroy(dead);
        throw;
    }
    LBM_CUDA_CHECK(cudaStreamEndCapture(s, &keep->graph));
    const int64_t per = g_launch_counter - l0;
    LBM_CUDA_CHECK(cudaGraphInstantiate(&keep->exec, keep->graph, 0));
    for (int i = 0; i < reps; ++i) LBM_CUDA_CHECK(cudaGraphLaunch(keep->exec, s));
    g_launch_counter += per * (reps - 1);
}

inline void exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, stream_t s) {
    size_t tmp_bytes = 0;
    LBM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, s));
    void* tmp = dev_alloc(tmp_bytes);
    LBM_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, s));
    LBM_CUDA_CHECK(cudaStreamSynchronize(s));
    dev_free(tmp);
    ++g_launch_counter;
}

#else
// ------------------------------------------------------------------ host test hook
typedef int stream_t;
struct BackendError {
    std::string msg;
    bool oom = false;
};
inline void* dev_alloc(size_t bytes) {
    void* p = malloc(bytes ? bytes : 8);
    if (!p) throw BackendError{"malloc failed"};
    return p;
}
inline void dev_free(void* p) { free(p); }
inline void dev_zero(void* p, size_t bytes, stream_t) { memset(p, 0, bytes); }
inline void dev_h2d(void* d, const void* h, size_t bytes, stream_t) { memcpy(d, h, bytes); }
inline void dev_d2h(void* h, const void* d, size_t bytes, stream_t) { memcpy(h, d, bytes); }
inline void dev_d2d(void* d, const void* s_, size_t bytes, stream_t) { memmove(d, s_, bytes); }
inline void dev_sync(stream_t) {}
extern thread_local int64_t g_launch_counter;
// Item order of a launch on the host (test hook).  On the GPU the items of one launch run in no particular order, so an
// operator must not read what another item of the SAME launch writes; the forward loop would hide such a hazard.
// LBM_HOST_ORDER=reverse | scatter runs the items backwards / in a strided permutation: the oracle comparisons of the CPU
// tier then fail on any intra-launch dependence (tests/test_hostcheck_order.py).
inline int host_launch_order() {
    static const int order = [] {
        const char* e = getenv("LBM_HOST_ORDER");
        return !e ? 0 : (!strcmp(e, "reverse") ? 1 : (!strcmp(e, "scatter") ? 2 : 0));
    }();
    return order;
}
template <class Op>
inline void launch(const Op& op, int64_t n, stream_t) {
    const int order = host_launch_order();
    if (order == 0) {
        for (int64_t i = 0; i < n; ++i) op(i);
    } else if (order == 1) {
        for (int64_t i = n - 1; i >= 0; --i) op(i);
    } else {
        int64_t stride = 7919;                       // a prime; coprime to n unless n is a multiple of it
        while (n % stride == 0) stride += 2;
        auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
        while (n > 1 && gcd(stride, n) != 1) ++stride;
        int64_t j = n / 3;
        for (int64_t k = 0; k < n; ++k) { op(j); j += stride % n; if (j >= n) j -= n; }
    }
    ++g_launch_counter;
}
template <int MINB, class Op>
inline void launch_occ(const Op& op, int64_t n, stream_t s) { launch(op, n, s); }
inline bool g_prof_active() { return false; }
struct GraphKeep { int unused = 0; };
inline void graph_release(GraphKeep*, stream_t) {}
template <class F>
inline void replay(int reps, bool, GraphKeep*, stream_t, F body) {
    for (int i = 0; i < reps; ++i) body();
}
inline void exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, stream_t) {
    int64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = in[i];
        out[i] = acc;
        acc += v;
    }
}
#endif

}  // namespace lbm
