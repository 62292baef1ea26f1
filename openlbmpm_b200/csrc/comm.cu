// comm.cu -- slab decomposition along the flow axis: one handle per rank, one rank per GPU, ghost
// planes exchanged with ncclSend/ncclRecv between ring neighbours over NVLink.  Planes of axis 2 are
// contiguous in every array, so they are sent in place: no packing kernels, no staging buffers.
// (The reference is single-GPU: RKD2Q9.py / ShanChenD2Q9.py hold the whole lattice on one device.)
#ifndef LBM_HOSTCHECK
#include <dlfcn.h>
#include <nccl.h>      // types only: the library is bound at run time (see nccl_api)

#include <vector>

#include "internal.h"

using namespace lbm;

// NCCL is resolved with dlopen/dlsym instead of a link-time dependency: a host process that has already
// loaded its own libnccl.so.2 (PyTorch bundles one) must keep using exactly that copy, and a process
// that never goes multi-GPU should not need NCCL at all.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    if (api.ok) return api;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the process already has
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw BackendError{std::string("cannot load libnccl.so.2: ") + dlerror()};
    auto sym = [&](const char* name) {
        void* p = dlsym(lib, name);
        if (!p) throw BackendError{std::string("libnccl lacks ") + name};
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = true;
    return api;
}
}  // namespace
#define ncclGetUniqueId nccl_api().GetUniqueId
#define ncclCommInitRank nccl_api().CommInitRank
#define ncclCommDestroy nccl_api().CommDestroy
#define ncclGroupStart nccl_api().GroupStart
#define ncclGroupEnd nccl_api().GroupEnd
#define ncclSend nccl_api().Send
#define ncclRecv nccl_api().Recv
#define ncclGetErrorString nccl_api().GetErrorString

#define LBM_NCCL_CHECK(expr)                                                               \
    do {                                                                                   \
        ncclResult_t _r = (expr);                                                          \
        if (_r != ncclSuccess)                                                             \
            throw ::lbm::BackendError{std::string(#expr) + ": " + ncclGetErrorString(_r)}; \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "lbm_nccl_unique_id assumes a 128-byte NCCL id");

extern "C" int lbm_nccl_unique_id(uint8_t id_out[128]) {
    if (!id_out) return LBM_EINVAL;
    try {
        ncclUniqueId id;
        if (ncclGetUniqueId(&id) != ncclSuccess) return LBM_ENCCL;
        memcpy(id_out, &id, 128);
    } catch (const BackendError&) { return LBM_ENCCL; }
    return LBM_OK;
}

extern "C" int lbm_comm_init(lbm_handle* h, int32_t rank, int32_t nranks, const uint8_t id_in[128]) {
    if (!h) return LBM_EINVAL;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id_in) { h->err = "bad rank / nranks / id"; return LBM_EINVAL; }
    if (h->has_geometry) { h->err = "lbm_comm_init must precede lbm_set_geometry"; return LBM_ESTATE; }
    try {
        LBM_CUDA_CHECK(cudaSetDevice(h->cfg.device));
        if (nranks > 1) {
            ncclUniqueId id;
            memcpy(&id, id_in, 128);
            ncclComm_t comm;
            LBM_NCCL_CHECK(ncclCommInitRank(&comm, nranks, id, rank));
            h->nccl = comm;
            LBM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
            LBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
            LBM_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
        }
        h->rank = rank; h->nranks = nranks;
    } catch (const BackendError& e) { h->err = e.msg; return LBM_ENCCL; }
    return LBM_OK;
}

namespace lbm {

// dirs (optional, one entry per array): 0 = ghost planes on both sides, +1 = only the planes that travel upwards
// (needed in the low ghost of the rank above), -1 = only downwards, 2 = none
template <class T>
static void ring_exchange(lbm_handle* h, T* base, int64_t stride, int narr, int gp, ncclDataType_t dt,
                          const int8_t* dirs = nullptr) {
    const Grid& g = h->g;
    ncclComm_t comm = (ncclComm_t)h->nccl;
    cudaStream_t st = h->xstream ? h->xstream : h->stream;
    const int up = (h->rank + 1) % h->nranks, down = (h->rank + h->nranks - 1) % h->nranks;
    const size_t count = (size_t)gp * g.plane;
    LBM_NCCL_CHECK(ncclGroupStart());
    for (int a = 0; a < narr; ++a) {
        T* f = base + a * stride;
        const int dir = dirs ? dirs[a] : 0;
        if (dir == 0 || dir == 1) {
            // my top planes -> low ghost of the rank above; my low ghost <- top planes of the rank below
            LBM_NCCL_CHECK(ncclSend(f + (int64_t)(NG + g.n2 - gp) * g.plane, count, dt, up, comm, st));
            LBM_NCCL_CHECK(ncclRecv(f + (int64_t)(NG - gp) * g.plane, count, dt, down, comm, st));
        }
        if (dir == 0 || dir == -1) {
            // my bottom planes -> high ghost of the rank below; my high ghost <- bottom planes of the rank above
            LBM_NCCL_CHECK(ncclSend(f + (int64_t)NG * g.plane, count, dt, down, comm, st));
            LBM_NCCL_CHECK(ncclRecv(f + (int64_t)(NG + g.n2) * g.plane, count, dt, up, comm, st));
        }
    }
    LBM_NCCL_CHECK(ncclGroupEnd());
    ++g_launch_counter;
}

// Packed variant for the per-step exchanges: the boundary planes of all arrays that travel in one direction are
// gathered into one staging buffer, so an exchange is 2 sends + 2 receives of tens of MB instead of 2 x narr
// messages of one plane each (NCCL point-to-point bandwidth grows with the message size); the pack / unpack
// kernels move the same bytes at HBM speed.
namespace {
constexpr int MAXARR = 48;
struct PackOp {
    Grid g; double* base; int64_t stride; int gp; int n_up, n_dn; int8_t up[MAXARR], dn[MAXARR];
    double *send_up, *send_dn; const double *recv_lo, *recv_hi; int unpack;
    LBM_HD void operator()(int64_t i) const {
        const int64_t per = (int64_t)gp * g.plane;
        const int64_t k = i / per, r = i % per;
        if (k < n_up) {            // arrays travelling upwards: my top planes out, my low ghost in
            double* f = base + up[k] * stride;
            if (!unpack) send_up[k * per + r] = f[(int64_t)(NG + g.n2 - gp) * g.plane + r];
            else f[(int64_t)(NG - gp) * g.plane + r] = recv_lo[k * per + r];
        } else {                   // arrays travelling downwards: my bottom planes out, my high ghost in
            const int64_t kk = k - n_up;
            double* f = base + dn[kk] * stride;
            if (!unpack) send_dn[kk * per + r] = f[(int64_t)NG * g.plane + r];
            else f[(int64_t)(NG + g.n2) * g.plane + r] = recv_hi[kk * per + r];
        }
    }
};
}  // namespace

void comm_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs) {
    // opt-in (LBM_FLAG_PACKED_EXCHANGE): bit-equal to the in-place exchange in tests/mgpu_check.py, but only +0.5 % at
    // 2 GPUs, so the simpler in-place exchange stays the default
    if (narr > MAXARR || !(h->cfg.flags & 32u)) { ring_exchange<double>(h, base, stride, narr, gp, ncclDouble, dirs); return; }
    const Grid& g = h->g;
    PackOp op;
    op.g = g; op.base = base; op.stride = stride; op.gp = gp; op.n_up = op.n_dn = 0;
    for (int a = 0; a < narr; ++a) {
        const int d = dirs ? dirs[a] : 0;
        if (d == 0 || d == 1) op.up[op.n_up++] = (int8_t)a;
        if (d == 0 || d == -1) op.dn[op.n_dn++] = (int8_t)a;
    }
    const int64_t per = (int64_t)gp * g.plane;
    const size_t need = (size_t)(op.n_up + op.n_dn) * per * sizeof(double);
    if (h->stage_bytes < need) {
        dev_sync(h->stream);
        if (h->comm_stream) LBM_CUDA_CHECK(cudaStreamSynchronize(h->comm_stream));
        dev_free(h->stage_send); dev_free(h->stage_recv);
        h->stage_send = (double*)dev_alloc(need); h->stage_recv = (double*)dev_alloc(need);
        h->stage_bytes = need;
    }
    op.send_up = h->stage_send; op.send_dn = h->stage_send + op.n_up * per;
    double* recv_lo = h->stage_recv; double* recv_hi = h->stage_recv + op.n_up * per;
    op.recv_lo = recv_lo; op.recv_hi = recv_hi;
    ncclComm_t comm = (ncclComm_t)h->nccl;
    cudaStream_t st = h->xstream ? h->xstream : h->stream;
    const int up = (h->rank + 1) % h->nranks, down = (h->rank + h->nranks - 1) % h->nranks;
    op.unpack = 0;
    launch(op, (int64_t)(op.n_up + op.n_dn) * per, st);
    LBM_NCCL_CHECK(ncclGroupStart());
    if (op.n_up) {
        LBM_NCCL_CHECK(ncclSend(op.send_up, (size_t)op.n_up * per, ncclDouble, up, comm, st));
        LBM_NCCL_CHECK(ncclRecv(recv_lo, (size_t)op.n_up * per, ncclDouble, down, comm, st));
    }
    if (op.n_dn) {
        LBM_NCCL_CHECK(ncclSend(op.send_dn, (size_t)op.n_dn * per, ncclDouble, down, comm, st));
        LBM_NCCL_CHECK(ncclRecv(recv_hi, (size_t)op.n_dn * per, ncclDouble, up, comm, st));
    }
    LBM_NCCL_CHECK(ncclGroupEnd());
    ++g_launch_counter;
    op.unpack = 1;
    launch(op, (int64_t)(op.n_up + op.n_dn) * per, st);
}
void comm_exchange_u8(lbm_handle* h, uint8_t* base, int gp) { ring_exchange<uint8_t>(h, base, 0, 1, gp, ncclUint8); }

// ------------------------------------------------------------------------------------------------
// One-sided exchange over peer memory (LBM_FLAG_PEER_EXCHANGE, experimental).  Every rank maps the neighbours' images of
// an exchanged allocation once (cudaIpcGetMemHandle / cudaIpcOpenMemHandle; the 64-byte handles travel over the NCCL
// communicator that exists anyway).  An exchange is then
//     PeerPushOp   my boundary planes are STORED into the neighbours' ghost planes (NVLink stores)
//     peer_signal  fence.sys, then the exchange number is released into the neighbours' flag words
//     peer_wait    spins (acquire) until both neighbours' numbers have arrived in my flag words
// on the handle's stream: no rendezvous, no staging.  Write-after-read safety on the fast path comes from its own schedule: the
// factored state is double buffered and its exchange alternates with the exchange of phi, so the neighbour that overwrites a
// ghost plane has waited for a signal its owner only sends after the kernels that read that plane (checked with
// ThreadSanitizer on the thread-rank emulation, host_stubs.cu).  First run on the hardware: tests/mgpu_check.py with
// LBM_TEST_FLAGS=256.
// ------------------------------------------------------------------------------------------------
namespace {
struct PeerMap { const void* base; void* up; void* down; };
struct PeerState {
    std::vector<PeerMap> maps;
    unsigned long long* flags = nullptr;          // [0]: written by the slab below, [1]: by the slab above, [2]: epoch of a timed-out wait; [4..6]: the same triple for comm_peer_release
    unsigned long long *up_flag = nullptr, *down_flag = nullptr;    // the neighbours' words I write
    std::vector<void*> opened;                    // mappings of the neighbours' exchanged arrays (dropped by comm_peer_release)
    std::vector<void*> opened_flags;              // mappings of the neighbours' flag words (live as long as the handle)
    unsigned long long close_gen = 0;             // number of comm_peer_release rounds (words [4], [5]: the neighbours' counts; [6]: timeout)
};
// what travels between neighbours to map one allocation: the IPC handle of the driver allocation that CONTAINS the pointer
// and the pointer's offset in it (cudaMalloc sub-allocates small requests from a larger block; the handle names the block)
struct PeerTicket { cudaIpcMemHandle_t handle; unsigned long long offset; unsigned long long valid; };
unsigned long long offset_in_allocation(const void* p) {
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);      // cuMemGetAddressRange (driver API)
    static range_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return (range_fn)f;
    }();
    unsigned long long base = 0; size_t size = 0;
    if (!fn || fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) throw BackendError{"cuMemGetAddressRange failed"};
    return (unsigned long long)(uintptr_t)p - base;
}

__global__ void peer_signal(unsigned long long* up_from_down, unsigned long long* down_from_up, unsigned long long epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(up_from_down), "l"(epoch) : "memory");
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(down_from_up), "l"(epoch) : "memory");
}
// The wait is bounded (LBM_PEER_TIMEOUT_MS, default 20 s of the global timer): a neighbour that died or fell out of step must
// not leave a kernel spinning on the GPU for ever.  A timeout is recorded in flags[2] and reported by the next synchronising call.
__global__ void peer_wait(unsigned long long* flags, unsigned long long epoch, unsigned long long timeout_ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = 0; k < 2; ++k) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + k) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { flags[2] = epoch; return; }
        } while (true);
    }
    __threadfence_system();
}

// swap one ticket with both ring neighbours and open theirs; -> {image in the slab above, image in the slab below}
void peer_open(lbm_handle* h, std::vector<void*>& opened, void* mine, void** up_img, void** down_img) {
    ncclComm_t comm = (ncclComm_t)h->nccl;
    const int up = (h->rank + 1) % h->nranks, down = (h->rank + h->nranks - 1) % h->nranks;
    PeerTicket tm, tu, td;
    memset(&tm, 0, sizeof(tm));
    // a local failure must not leave the neighbours waiting in the swap: the ticket travels in any case, marked invalid
    std::string local_error;
    try {
        tm.offset = offset_in_allocation(mine);
        LBM_CUDA_CHECK(cudaIpcGetMemHandle(&tm.handle, (char*)mine - tm.offset));
        tm.valid = 1;
    } catch (const BackendError& e) { local_error = e.msg; cudaGetLastError(); }
    char* d = (char*)dev_alloc(3 * sizeof(tm));
    try {
        dev_h2d(d, &tm, sizeof(tm), h->stream);
        LBM_NCCL_CHECK(ncclGroupStart());
        LBM_NCCL_CHECK(ncclSend(d, sizeof(tm), ncclUint8, up, comm, h->stream));
        LBM_NCCL_CHECK(ncclRecv(d + 2 * sizeof(tm), sizeof(tm), ncclUint8, down, comm, h->stream));
        LBM_NCCL_CHECK(ncclSend(d, sizeof(tm), ncclUint8, down, comm, h->stream));
        LBM_NCCL_CHECK(ncclRecv(d + sizeof(tm), sizeof(tm), ncclUint8, up, comm, h->stream));
        LBM_NCCL_CHECK(ncclGroupEnd());
        dev_d2h(&tu, d + sizeof(tm), sizeof(tm), h->stream);
        dev_d2h(&td, d + 2 * sizeof(tm), sizeof(tm), h->stream);
    } catch (...) { dev_free(d); throw; }
    dev_free(d);
    if (!tm.valid) throw BackendError{"one-sided exchange: " + local_error};
    if (!tu.valid || !td.valid) throw BackendError{"one-sided exchange: a neighbour slab could not export its memory"};
    void* base_up = nullptr;
    LBM_CUDA_CHECK(cudaIpcOpenMemHandle(&base_up, tu.handle, cudaIpcMemLazyEnablePeerAccess));
    opened.push_back(base_up);
    *up_img = (char*)base_up + tu.offset;
    if (up == down) { *down_img = *up_img; return; }       // two slabs: one neighbour, one mapping
    void* base_down = nullptr;
    LBM_CUDA_CHECK(cudaIpcOpenMemHandle(&base_down, td.handle, cudaIpcMemLazyEnablePeerAccess));
    opened.push_back(base_down);
    *down_img = (char*)base_down + td.offset;
}
}  // namespace

static PeerState* peer_state(lbm_handle* h) {
    PeerState* ps = (PeerState*)h->peer;
    if (ps) return ps;
    ps = new PeerState();
    h->peer = ps;
    ps->flags = (unsigned long long*)dev_alloc(8 * sizeof(unsigned long long));
    dev_zero(ps->flags, 8 * sizeof(unsigned long long), h->stream);
    dev_sync(h->stream);
    void *fu = nullptr, *fd = nullptr;
    try {
        peer_open(h, ps->opened_flags, ps->flags, &fu, &fd);
    } catch (...) {
        for (void* p : ps->opened_flags) cudaIpcCloseMemHandle(p);
        dev_free(ps->flags);
        delete ps;
        h->peer = nullptr;
        throw;
    }
    ps->up_flag = (unsigned long long*)fu;              // word [0] of the slab above: "from the slab below"
    ps->down_flag = (unsigned long long*)fd + 1;        // word [1] of the slab below: "from the slab above"
    return ps;
}
void comm_peer_pointers(lbm_handle* h, double* base, double** up, double** down) {
    PeerState* ps = peer_state(h);
    for (const PeerMap& k : ps->maps)
        if (k.base == base) { *up = (double*)k.up; *down = (double*)k.down; return; }
    PeerMap n{base, nullptr, nullptr};
    peer_open(h, ps->opened, base, &n.up, &n.down);
    ps->maps.push_back(n);
    *up = (double*)n.up; *down = (double*)n.down;
}
void comm_peer_signal_wait(lbm_handle* h) {
    PeerState* ps = peer_state(h);
    const unsigned long long epoch = ++h->peer_epoch;
    peer_signal<<<1, 1, 0, h->stream>>>(ps->up_flag, ps->down_flag, epoch);
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("LBM_PEER_TIMEOUT_MS");
        return (unsigned long long)(e ? atoll(e) : 20000) * 1000000ull;
    }();
    peer_wait<<<1, 1, 0, h->stream>>>(ps->flags, epoch, timeout_ns);
    LBM_CUDA_CHECK(cudaGetLastError());
    g_launch_counter += 2;
}
void comm_peer_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs) {
    if (narr > 48) throw BackendError{"peer exchange: too many arrays"};
    PeerPushOp op;
    comm_peer_pointers(h, base, &op.up, &op.down);
    op.g = h->g; op.base = base; op.stride = stride; op.narr = narr; op.gp = gp;
    for (int a = 0; a < 48; ++a) op.dirs[a] = (a < narr && dirs) ? dirs[a] : 0;
    launch(op, op.items(), h->stream);
    comm_peer_signal_wait(h);
}

bool comm_peer_probe(lbm_handle* h) {
    int bad = 0;
    try { peer_state(h); } catch (const BackendError&) { bad = 1; cudaGetLastError(); }
    if (comm_allreduce_max(h, bad) == 0) return true;
    // some slab cannot: everybody drops what it has (no barrier -- the slab that failed holds nothing)
    PeerState* ps = (PeerState*)h->peer;
    if (ps) {
        for (void* p : ps->opened_flags) cudaIpcCloseMemHandle(p);
        delete ps;              // ps->flags stays allocated: a neighbour that did map it may close its mapping later
        h->peer = nullptr;
    }
    return false;
}
// after a stream synchronisation: did a wait of the one-sided exchange give up?
void comm_peer_check(lbm_handle* h) {
    PeerState* ps = (PeerState*)h->peer;
    if (!ps) return;
    unsigned long long t = 0;
    dev_d2h(&t, ps->flags + 2, sizeof(t), h->stream);
    if (t) throw BackendError{"one-sided exchange " + std::to_string(t) + ": a neighbour slab's signal did not arrive within LBM_PEER_TIMEOUT_MS"};
}
bool comm_peer_release(lbm_handle* h) {
    PeerState* ps = (PeerState*)h->peer;
    if (!ps || h->nranks <= 1 || ps->maps.empty()) return !h->peer_unreachable;
    cudaStreamSynchronize(h->stream);
    for (void* p : ps->opened) cudaIpcCloseMemHandle(p);
    ps->opened.clear();
    ps->maps.clear();
    // "I no longer map your arrays" into the neighbours' words [4] / [5]; wait (bounded) for theirs
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("LBM_PEER_CLOSE_TIMEOUT_MS");
        return (unsigned long long)(e ? atoll(e) : 3000) * 1000000ull;
    }();
    const unsigned long long gen = ++ps->close_gen;
    unsigned long long late = 1;
    try {
        peer_signal<<<1, 1, 0, h->stream>>>(ps->up_flag + 4, ps->down_flag + 4, gen);
        peer_wait<<<1, 1, 0, h->stream>>>(ps->flags + 4, gen, timeout_ns);
        dev_d2h(&late, ps->flags + 2 + 4, sizeof(late), h->stream);      // peer_wait records a timeout two words behind its pair
        dev_zero(ps->flags + 2 + 4, sizeof(unsigned long long), h->stream);
    } catch (const BackendError&) { cudaGetLastError(); }
    if (late) h->peer_unreachable = true;
    return !h->peer_unreachable;
}

static void peer_destroy(lbm_handle* h) {
    PeerState* ps = (PeerState*)h->peer;
    if (!ps) return;
    comm_peer_release(h);
    cudaStreamSynchronize(h->stream);
    for (void* p : ps->opened_flags) cudaIpcCloseMemHandle(p);
    // ps->flags is NOT freed: a neighbour may still have the words mapped (its own destroy may come later), and freeing under a
    // mapping is undefined; 64 bytes until the process exits
    delete ps;
    h->peer = nullptr;
}
// maximum of an integer over all slabs (used for decisions every rank must take identically)
int comm_allreduce_max(lbm_handle* h, int v) {
    if (h->nranks <= 1) return v;
    int* d = (int*)dev_alloc(sizeof(int));
    int out = v;
    try {
        dev_h2d(d, &v, sizeof(int), h->stream);
        LBM_NCCL_CHECK(nccl_api().AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)h->nccl, h->stream));
        dev_d2h(&out, d, sizeof(int), h->stream);
    } catch (...) { dev_free(d); throw; }
    dev_free(d);
    return out;
}

void comm_destroy(lbm_handle* h) {
    peer_destroy(h);
    if (h->nccl) {
        try { ncclCommDestroy((ncclComm_t)h->nccl); } catch (const BackendError&) {}
        h->nccl = nullptr;
    }
    dev_free(h->stage_send); dev_free(h->stage_recv);
    h->stage_send = h->stage_recv = nullptr; h->stage_bytes = 0;
    if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); h->comm_stream = nullptr; }
    if (h->ev_main) { cudaEventDestroy(h->ev_main); h->ev_main = nullptr; }
    if (h->ev_comm) { cudaEventDestroy(h->ev_comm); h->ev_comm = nullptr; }
}

}  // namespace lbm
#endif
