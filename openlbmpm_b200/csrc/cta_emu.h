// cta_emu.h -- the tiled sm_100a kernels of cg_fast.cu on HOST threads.  TEST HOOK ONLY (tests/hostcheck): one host
// thread per CUDA thread of a CTA, a pthread barrier for __syncthreads, a heap buffer for the dynamic shared memory,
// CTAs one after the other.  Asynchronous copies (cp.async) become immediate copies and the TMA / mbarrier variants are
// not instantiated, so what this checks is the LOGIC of the kernels -- tile indexing, halo windows, slot rotation, the
// barrier protocol (a missing barrier shows up as a data race between real threads), the pull masks, the plane ranges of
// the open-boundary patching -- against the same oracle the GPU tier uses.  Speed is irrelevant (small lattices only).
#pragma once
#ifdef LBM_HOSTCHECK
#include <pthread.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

namespace cta_emu {
struct Dim3 {
    unsigned x, y, z;
    Dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct Tls {
    Dim3 tid, bid, bdim, gdim;
    double* shared = nullptr;
    pthread_barrier_t* bar = nullptr;
    pthread_barrier_t* grid_bar = nullptr;       // cooperative launches only
};
inline Tls& tls() {
    static thread_local Tls t;
    return t;
}
inline void sync() { pthread_barrier_wait(tls().bar); }
inline void grid_sync() { pthread_barrier_wait(tls().grid_bar); }

// cooperative launch: every CTA of the (small) grid is resident at once, grid_sync() is a barrier over all of its threads
template <class F>
inline void launch_cooperative(Dim3 grid, Dim3 block, F kernel) {
    const unsigned nt = block.x * block.y * block.z, nb = grid.x * grid.y * grid.z;
    pthread_barrier_t gbar;
    pthread_barrier_init(&gbar, nullptr, nt * nb);
    std::vector<pthread_barrier_t> bars(nb);
    for (auto& b : bars) pthread_barrier_init(&b, nullptr, nt);
    std::vector<std::thread> threads;
    threads.reserve((size_t)nt * nb);
    for (unsigned b = 0; b < nb; ++b)
        for (unsigned t = 0; t < nt; ++t)
            threads.emplace_back([&, b, t] {
                Tls& l = tls();
                l.tid = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                l.bid = Dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
                l.bdim = block; l.gdim = grid; l.shared = nullptr; l.bar = &bars[b]; l.grid_bar = &gbar;
                kernel();
            });
    for (auto& th : threads) th.join();
    for (auto& b : bars) pthread_barrier_destroy(&b);
    pthread_barrier_destroy(&gbar);
}

template <class F>
inline void launch(Dim3 grid, Dim3 block, size_t smem_bytes, F kernel) {
    const unsigned nt = block.x * block.y * block.z;
    std::vector<double> smem(smem_bytes / sizeof(double) + 32);
    double* base = (double*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                pthread_barrier_t bar;
                pthread_barrier_init(&bar, nullptr, nt);
                std::vector<std::thread> threads;
                threads.reserve(nt);
                for (unsigned t = 0; t < nt; ++t)
                    threads.emplace_back([&, t] {
                        Tls& l = tls();
                        l.tid = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        l.bid = Dim3(bx, by, bz); l.bdim = block; l.gdim = grid; l.shared = base; l.bar = &bar;
                        kernel();
                    });
                for (auto& th : threads) th.join();
                pthread_barrier_destroy(&bar);
            }
}
}  // namespace cta_emu

// ---- the CUDA vocabulary the tiled kernels use ----
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)
typedef cta_emu::Dim3 dim3;
#define threadIdx (cta_emu::tls().tid)
#define blockIdx (cta_emu::tls().bid)
#define blockDim (cta_emu::tls().bdim)
#define gridDim (cta_emu::tls().gdim)
#define LBM_GRID_SYNC() cta_emu::grid_sync()
#define __syncthreads() cta_emu::sync()
#define __syncwarp() ((void)0)          // only in the TMA branches, which are never taken here
#define __pipeline_memcpy_async(dst, src, n) memcpy((dst), (src), (n))
#define __pipeline_commit() ((void)0)
#define __pipeline_wait_prior(n) ((void)0)
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::min;
#endif
