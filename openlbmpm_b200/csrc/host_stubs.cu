// host_stubs.cu -- compiled ONLY into the host test hook (tests/hostcheck): stand-ins for the pieces that exist solely
// as CUDA code.  The tiled kernels are simply absent; the NCCL halo exchange is replaced by an in-process ring.
#ifdef LBM_HOSTCHECK
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "internal.h"

// ---- slab decomposition inside ONE process (test hook) ---------------------------------------------------------
// The CUDA build exchanges ghost planes with ncclSend / ncclRecv between one process per GPU (comm.cu).  Here the
// "ranks" are host threads of the test process, one handle each, and an exchange is: publish my array, barrier, copy
// the neighbours' boundary planes into my ghost planes, barrier.  Same ring, same planes, same `dirs` filter as
// comm.cu::ring_exchange -- so the CPU tier can demand that P slabs reproduce the single-slab run bit for bit for
// every model and kernel path (tests/test_hostcheck_slabs.py), i.e. it checks WHICH planes the step loops exchange
// and WHEN, everything about the decomposition except NCCL itself.
namespace {
struct HostRing {
    int n = 0;
    std::vector<lbm_handle*> members;
    std::vector<void*> pub;
    std::vector<int> ival;
    pthread_barrier_t bar;
    // one-sided exchange: flag words of every slab ("from the slab below", "from the slab above")
    std::unique_ptr<std::atomic<uint64_t>[]> from_down, from_up;
};
struct HostPeerMap { const void* base; double* up; double* down; };
struct HostPeerState { std::vector<HostPeerMap> maps; };
std::mutex g_ring_mutex;
std::map<std::string, HostRing*> g_rings;
uint64_t g_ring_counter = 0;

template <class T>
void host_ring_exchange(lbm_handle* h, T* base, int64_t stride, int narr, int gp, const int8_t* dirs) {
    HostRing* r = (HostRing*)h->nccl;
    const lbm::Grid& g = h->g;
    r->pub[h->rank] = base;
    pthread_barrier_wait(&r->bar);                      // every slab has written its own planes and published its array
    const int up = (h->rank + 1) % r->n, down = (h->rank + r->n - 1) % r->n;
    const T* bup = (const T*)r->pub[up];
    const T* bdown = (const T*)r->pub[down];
    const size_t bytes = (size_t)gp * g.plane * sizeof(T);
    for (int a = 0; a < narr; ++a) {
        T* f = base + a * stride;
        const int dir = dirs ? dirs[a] : 0;
        if (dir == 0 || dir == 1)           // my low ghost <- top planes of the rank below
            memcpy(f + (int64_t)(lbm::NG - gp) * g.plane, bdown + a * stride + (int64_t)(lbm::NG + g.n2 - gp) * g.plane, bytes);
        if (dir == 0 || dir == -1)          // my high ghost <- bottom planes of the rank above
            memcpy(f + (int64_t)(lbm::NG + g.n2) * g.plane, bup + a * stride + (int64_t)lbm::NG * g.plane, bytes);
    }
    pthread_barrier_wait(&r->bar);                      // nobody overwrites its planes before the neighbours have read them
    ++lbm::g_launch_counter;
}
}  // namespace

namespace lbm {
void comm_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs) {
    host_ring_exchange(h, base, stride, narr, gp, dirs);
}
void comm_exchange_u8(lbm_handle* h, uint8_t* base, int gp) { host_ring_exchange(h, base, 0, 1, gp, nullptr); }
// The one-sided exchange of comm.cu on thread ranks: the neighbours' arrays are plain pointers, the push is the same operator,
// the flag words are std::atomic (release store / acquire spin).  Only the FIRST exchange of an array meets at a barrier (the
// pointer swap that cudaIpc does on the GPU); after that the ranks synchronise through the flags alone, so ThreadSanitizer sees
// the protocol as it is.
void comm_peer_pointers(lbm_handle* h, double* base, double** up, double** down) {
    HostRing* r = (HostRing*)h->nccl;
    HostPeerState* ps = (HostPeerState*)h->peer;
    if (!ps) { ps = new HostPeerState(); h->peer = ps; }
    for (const HostPeerMap& k : ps->maps)
        if (k.base == base) { *up = k.up; *down = k.down; return; }
    const int upr = (h->rank + 1) % r->n, downr = (h->rank + r->n - 1) % r->n;
    r->pub[h->rank] = base;
    pthread_barrier_wait(&r->bar);
    HostPeerMap n{base, (double*)r->pub[upr], (double*)r->pub[downr]};
    pthread_barrier_wait(&r->bar);
    ps->maps.push_back(n);
    *up = n.up; *down = n.down;
}
void comm_peer_signal_wait(lbm_handle* h) {
    HostRing* r = (HostRing*)h->nccl;
    const int up = (h->rank + 1) % r->n, down = (h->rank + r->n - 1) % r->n;
    const uint64_t epoch = ++h->peer_epoch;
    r->from_down[up].store(epoch, std::memory_order_release);        // signal
    r->from_up[down].store(epoch, std::memory_order_release);
    // wait -- bounded, like comm.cu::peer_wait: a neighbour that died must not hang the others
    const auto t0 = std::chrono::steady_clock::now();
    auto late = [&] { return std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60); };
    while (r->from_down[h->rank].load(std::memory_order_acquire) < epoch) { sched_yield(); if (late()) throw BackendError{"one-sided exchange: no signal from the slab below"}; }
    while (r->from_up[h->rank].load(std::memory_order_acquire) < epoch) { sched_yield(); if (late()) throw BackendError{"one-sided exchange: no signal from the slab above"}; }
    lbm::g_launch_counter += 2;
}
void comm_peer_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs) {
    PeerPushOp op;
    comm_peer_pointers(h, base, &op.up, &op.down);
    op.g = h->g; op.base = base; op.stride = stride; op.narr = narr; op.gp = gp;
    for (int a = 0; a < 48; ++a) op.dirs[a] = (a < narr && dirs) ? dirs[a] : 0;
    launch(op, op.items(), h->stream);
    comm_peer_signal_wait(h);
}
bool comm_peer_release(lbm_handle* h) {
    // thread ranks: a neighbour's stale pointer is never dereferenced again (nothing is in flight between lbm_step calls, and the
    // next exchange re-publishes the arrays), so there is nothing to wait for
    HostPeerState* ps = (HostPeerState*)h->peer;
    if (ps) ps->maps.clear();
    return true;
}
bool comm_peer_probe(lbm_handle*) { return true; }
void comm_peer_check(lbm_handle*) {}     // the host wait throws by itself
void comm_destroy(lbm_handle* h) { delete (HostPeerState*)h->peer; h->peer = nullptr; }
int comm_allreduce_max(lbm_handle* h, int v) {
    if (h->nranks <= 1) return v;
    HostRing* r = (HostRing*)h->nccl;
    r->ival[h->rank] = v;
    pthread_barrier_wait(&r->bar);
    int m = v;
    for (int k = 0; k < r->n; ++k) m = r->ival[k] > m ? r->ival[k] : m;
    pthread_barrier_wait(&r->bar);
    return m;
}
}  // namespace lbm
extern "C" int lbm_nccl_unique_id(uint8_t* id_out) {
    if (!id_out) return LBM_EINVAL;
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    memset(id_out, 0, 128);
    const uint64_t c = ++g_ring_counter;
    memcpy(id_out, &c, sizeof(c));
    memcpy(id_out + 8, "hostring", 8);
    return LBM_OK;
}
extern "C" int lbm_comm_init(lbm_handle* h, int32_t rank, int32_t nranks, const uint8_t* id) {
    if (!h) return LBM_EINVAL;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id) { h->err = "bad rank / nranks / id"; return LBM_EINVAL; }
    if (h->has_geometry) { h->err = "lbm_comm_init must precede lbm_set_geometry"; return LBM_ESTATE; }
    if (nranks > 1) {
        std::lock_guard<std::mutex> lock(g_ring_mutex);
        HostRing*& r = g_rings[std::string((const char*)id, 128)];
        if (!r) {
            r = new HostRing();
            r->n = nranks; r->members.assign(nranks, nullptr); r->pub.assign(nranks, nullptr); r->ival.assign(nranks, 0);
            pthread_barrier_init(&r->bar, nullptr, (unsigned)nranks);
            r->from_down.reset(new std::atomic<uint64_t>[nranks]); r->from_up.reset(new std::atomic<uint64_t>[nranks]);
            for (int k = 0; k < nranks; ++k) { r->from_down[k].store(0); r->from_up[k].store(0); }
        }
        if (r->n != nranks) { h->err = "nranks differs between the members of one ring"; return LBM_EINVAL; }
        r->members[rank] = h;
        h->nccl = r;
    }
    h->rank = rank; h->nranks = nranks;
    return LBM_OK;
}
// test hook: the two forms of the 3-D Akai wetting correction (reference-ordered / the tiled kernels' fast form)
extern "C" void hostcheck_wetting3(const double* G, const double* ns, double cosT, double sinT, int fast, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        double g[3] = {G[3 * i], G[3 * i + 1], G[3 * i + 2]};
        if (fast) lbm::cg_wetting_akai3_fast(g, ns + 3 * i, cosT, sinT);
        else lbm::cg_wetting<3>(g, ns + 3 * i, cosT, sinT, 2);
        for (int a = 0; a < 3; ++a) out[3 * i + a] = g[a];
    }
}
#endif
