// cg_fast.cu -- orchestration of the factored colour-gradient fast path (cg_fast_ops.cuh) and the tiled
// sm_100a kernel of its collision pass.
//
// State machine (handle.h): between lbm_step calls the lattice is either STREAMED (fS, rho: what the
// reference holds between loop iterations, RKD2Q9.py:1295) or POST-COLLISION FACTORED (fast->buf[cur],
// streaming pending).  Entering costs one general collision, leaving costs one materialising pull;
// consecutive lbm_step calls stay in factored form.
#include "cg_fast_ops.cuh"
#include "internal.h"
#include <atomic>
#include "coop.h"         // grid-wide barrier; on the host (test hook) also the CUDA vocabulary of the kernels below, cta_emu.h
#ifndef LBM_HOSTCHECK
#include <cuda_pipeline.h>
#endif

namespace lbm {

struct FastState {
    double* buf[2] = {nullptr, nullptr};   // each: gT [Q][vol], kR [vol], a [3][vol]
    int cur = 0;
    int pushed[2] = {0, 0};                // one-sided exchange: 1 = the pass that wrote buf[k] has already stored its boundary planes
                                           // into the neighbour slabs (only the flag handshake is left), 2 = handshake done too
};

static FastFields fast_fields(const lbm_handle* h, int k) {
    const FastState* f = (const FastState*)h->fast;
    FastFields o;
    o.gT = f->buf[k]; o.kR = f->buf[k] + (int64_t)h->Q * h->g.vol; o.a = o.kR + h->g.vol;
    return o;
}

static bool open_box(const lbm_handle* h) { return h->cfg.inlet != LBM_BC_PERIODIC || h->cfg.outlet != LBM_BC_PERIODIC; }

static bool tiled_ok(const lbm_handle* h);

bool cg_fast_eligible(const lbm_handle* h) {
    // open boundaries: the treated planes are patched around the two passes (fast_open_rows_*), which needs the
    // inlet and outlet planes of a slab to be apart.  WettingType 1 turns ANY non-zero |G| into a unit normal
    // (AcceleratedRKGPU2D.py:1700-1706), so its trajectories hang on the exact cancellation of phi between the copied
    // outlet rows; the factored arithmetic rounds differently there, and that combination stays on the
    // reference-ordered kernels.
    if (h->cfg.model != LBM_MODEL_CG || (h->cfg.flags & LBM_FLAG_GENERIC_KERNELS)) return false;
    // the perturbation-operator model has its own collision pass (PullPerturbCollideOp); tracers ride on the CSF flow only
    if (h->cfg.surface_tension_type != LBM_ST_CSF && h->tracer) return false;
    // Tracers read u, G and rho_R of the iteration, none of which the flow collision changes: on the one-thread-per-node fast
    // path (closed boxes) their phase runs right after the collision pass, which then also stores u.  The tiled kernels keep G
    // in shared memory, and the open-row patches rewrite u on a few planes: those combinations stay on the reference-ordered
    // kernels.
    if (h->tracer && (open_box(h) || tiled_ok(h))) return false;
    return !open_box(h) || (h->g.n2 >= 8 && h->cfg.wetting_type != 1);
}

void cg_fast_free(lbm_handle* h) {
    FastState* f = (FastState*)h->fast;
    // one-sided exchange: the neighbour slabs have the buffers mapped.  If they cannot be reached any more (a rank that died),
    // the buffers are deliberately not returned to the driver (the process is on its way out) rather than freed under a mapping.
    const bool safe = f ? comm_peer_release(h) : true;
    if (f) { if (safe) { dev_free(f->buf[0]); dev_free(f->buf[1]); } delete f; }
    h->fast = nullptr;
    h->fast_pending_stream = false;
}

// a new state on the same geometry (lbm_init_equilibrium, lbm_upload_state, lbm_init_spinodal_device): the streamed state is
// authoritative again, the factored buffers stay allocated -- and mapped by the neighbour slabs -- for the next entry (the entry
// collision rewrites every fluid node of buf[cur], the exchange its ghost planes, and the solid nodes were zeroed once and are
// never written).  Saves freeing, re-allocating and zeroing 2 x 25 GB per re-initialisation of a 512^3 lattice.
void cg_fast_reset(lbm_handle* h) {
    FastState* f = (FastState*)h->fast;
    if (f) { f->cur = 0; f->pushed[0] = f->pushed[1] = 0; }
    h->fast_pending_stream = false;
}

static void fast_alloc(lbm_handle* h) {
    if (h->fast) return;
    FastState* f = new FastState();
    h->fast = f;
    const size_t bytes = (size_t)(h->Q + 4) * h->g.vol * sizeof(double);
    for (int k = 0; k < 2; ++k) { f->buf[k] = (double*)dev_alloc(bytes); dev_zero(f->buf[k], bytes, h->stream); }
}

#ifdef LBM_HOSTCHECK
// the TMA variants are never instantiated on the host (cta_emu.h); their calls sit in `if (TMA)` branches and must compile
inline void mbar_init(uint64_t*, int) {}
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
inline void bulk_g2s(void*, const void*, uint32_t, uint64_t*) {}
inline void mbar_wait(uint64_t*, uint32_t) {}
#define LBM_DYN_SMEM(name) double* name = cta_emu::tls().shared
#define LBM_OPAQUE(...) ((void)0)
#define LBM_PDL_PROLOGUE() ((void)0)
#else
#define LBM_DYN_SMEM(name) extern __shared__ __align__(128) double name[]
#define LBM_PDL_PROLOGUE() pdl_prologue()       // backend.h: programmatic dependent launch
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

// ------------------------------------------------------------------------------------------------
// Tiled collision pass for D3Q19.  One CTA owns a TX x TY column of the lattice and marches along z.
// phi is staged in shared memory as a rolling window of 5 planes with a 2-node halo, the interface
// normals n = -+G/|G| (and |G|) of the current 3 planes with a 1-node halo are derived from it in
// shared memory, so that every node's curvature stencil (18 neighbour normals) is served on chip and
// phi is read from HBM once.  The 19 pulled populations of a node are requested before the tile
// synchronises, so the HBM latency overlaps the shared-memory phase.
// ------------------------------------------------------------------------------------------------
// TMA = true: the phi planes are staged with TMA bulk copies (one cp.async.bulk per tile row, 288 B, issued by one
// thread, completing on a per-slot mbarrier) instead of 8-byte cp.async copies issued by every thread.
// PEER = true (LBM_FLAG_PEER_EXCHANGE on slabs): the nodes of the slab's first / last plane store their results a second time,
// through the peer pointers, into the ghost planes of the neighbour slabs -- the halo exchange rides on the pass that produces
// the data (NVLink stores next to the HBM stores); only the flag handshake is left between the passes (comm.cu).
struct PeerPtrs {
    double* up;      // image of the written array in the slab above (its low ghost planes receive my top planes)
    double* down;    // ... in the slab below (its high ghost planes receive my bottom planes)
    int gp;          // ghost planes that travel
};
// AHEAD = how many planes the interface normals run ahead of the collision.  1 (default): the normals of plane z + 1 are derived
// between two barriers of plane step z.  2 (LBM_COLLIDE_AHEAD=2): the normals of plane z + 2 are derived in step z into a ring of
// FOUR slots, so what the curvature stencil of plane z reads was written one step earlier and a step needs ONE barrier, at its
// end.  Measured on the B200 (round 2): the one-barrier form halves the barrier stalls (2.55 -> 1.19 warps per issue) and is
// SLOWER -- 10.87 vs 10.25 ms per 512^3 launch, 3.40 vs 3.18 ms on the porous workload: the wait at the first barrier was where
// the 24 HBM requests of a plane landed, without it the same latency shows up as long-scoreboard stalls on the first use of the
// populations (1.55 -> 1.92).  The kernel is bound by requests in flight per SM (16 warps at 126 registers), not by the barriers.
// PF (all-fluid lattices, LBM_POP_PREFETCH): the 24 per-node inputs of plane z + 1 (19 pulled populations, rhoR, rhoB, lagged force)
// are requested during plane step z with 8-byte cp.async into THREAD-PRIVATE shared-memory slots (no barrier: a thread only ever
// reads what it copied itself) and picked up at the top of step z + 1 -- a request pipeline one full plane step deep that costs no
// registers, where the plain loads of a step have only the shared-memory phase of the same step to land.
template <bool SOLIDS, int TX, int TY, bool TMA, bool PEER = false, int AHEAD = 1, bool PF = false>
__global__ void __launch_bounds__(TX* TY, 512 / (TX * TY) > 0 ? 512 / (TX * TY) : 1)
cg_collide_tiled_d3q19(const CGFields c, const FastFields s, const FastFields o, const int zchunk, const int z_lo, const int z_hi,
                       const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    using L = D3Q19;
    constexpr int NT = TX * TY;
    constexpr int PW = TX + 4, PH = TY + 4;     // phi tile
    constexpr int NW = TX + 2, NH = TY + 2;     // normal tile
    LBM_DYN_SMEM(smem_dyn);
    double (*sphi)[PH][PW] = reinterpret_cast<double (*)[PH][PW]>(smem_dyn);                    // [5]
    constexpr int NS = AHEAD + 2;               // slots of the normal ring
    double (*sn)[4][NH][NW] = reinterpret_cast<double (*)[4][NH][NW]>(smem_dyn + 5 * PH * PW);  // [NS][nx, ny, nz, |G|]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + 5 * PH * PW + NS * 4 * NH * NW);     // [5] one mbarrier per phi slot
    // SOLIDS: solid normals of the near-solid elements of the normal tile, two plane slots (copied one plane step ahead)
    double (*sns)[3][NH][NW] = reinterpret_cast<double (*)[3][NH][NW]>(smem_dyn + 5 * PH * PW + NS * 4 * NH * NW + 8);   // [2]
    double* spop = smem_dyn + 5 * PH * PW + NS * 4 * NH * NW + 8 + (SOLIDS ? 2 * 3 * NH * NW : 0);    // PF: [24][NT] thread-private
    const Grid& g = c.g;
    const int64_t V = g.vol;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = z_lo + blockIdx.z * zchunk;
    const int z_end = min(z_begin + zchunk, z_hi);
    const int x = x0 + tx, y = y0 + ty;

    auto wrapx = [&](int v) { return v < 0 ? v + g.n0 : (v >= g.n0 ? v - g.n0 : v); };
    auto wrapy = [&](int v) { return v < 0 ? v + g.n1 : (v >= g.n1 ? v - g.n1 : v); };

    // phi plane zp -> shared memory with cp.async (LDGSTS): the copy is issued one plane step before the
    // plane is needed, so its HBM/L2 latency never stalls the tile
    auto load_phi_plane = [&](int zp) {
        const int slot = (zp + 10) % 5;
        const double* src = c.phi + (int64_t)(zp + NG) * g.plane;
        if (TMA) {
            // one thread: arm the slot's mbarrier with the plane's byte count, then one bulk copy per row; rows of the
            // first / last tile in x wrap around and are split into two 16-byte-aligned pieces
            if (tid == 0) {
                mbar_arrive_expect_tx(&bars[slot], PH * PW * 8);
                for (int ly = 0; ly < PH; ++ly) {
                    const double* row = src + (int64_t)wrapy(y0 + ly - 2) * g.n0;
                    double* dst = &sphi[slot][ly][0];
                    const int c0 = x0 - 2, cs = c0 < 0 ? 0 : c0, ce = x0 + TX + 2 > g.n0 ? g.n0 : x0 + TX + 2;
                    if (c0 < 0) bulk_g2s(dst, row + g.n0 - 2, 16, &bars[slot]);                    // periodic image on the left
                    bulk_g2s(dst + (cs - c0), row + cs, (uint32_t)(ce - cs) * 8, &bars[slot]);
                    if (x0 + TX + 2 > g.n0) bulk_g2s(dst + PW - 2, row, 16, &bars[slot]);         // periodic image on the right
                }
            }
            return;
        }
        for (int e = tid; e < PH * PW; e += NT) {
            const int ly = e / PW, lx = e - ly * PW;
            __pipeline_memcpy_async(&sphi[slot][ly][lx], src + (int64_t)wrapy(y0 + ly - 2) * g.n0 + wrapx(x0 + lx - 2), 8);
        }
    };
    // elements of the normal tile this thread fills (NE per plane) and their in-plane offsets; with solids the
    // node classes of the NEXT plane are requested together with the populations, before the tile synchronises
    constexpr int NE = (NH * NW + NT - 1) / NT;
    int noff[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = tid + k * NT, ly = e / NW, lx = e - ly * NW;
        noff[k] = e < NH * NW ? wrapy(y0 + ly - 1) * g.n0 + wrapx(x0 + lx - 1) : -1;
    }
    auto class_plane = [&](int zp, uint8_t* cl) {
#pragma unroll
        for (int k = 0; k < NE; ++k) cl[k] = (SOLIDS && noff[k] >= 0) ? c.cls[(int64_t)(zp + NG) * g.plane + noff[k]] : (uint8_t)CLS_FLUID;
    };
    // solid normals of plane zp -> shared memory (cp.async), for the elements whose class says "fluid next to a solid"
    auto ns_plane = [&](int zp, const uint8_t* cl) {
        const int slot = (zp + 8) & 1;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const int e = tid + k * NT;
            if (e >= NH * NW || (cl[k] & (CLS_FLUID | CLS_NEAR)) != (CLS_FLUID | CLS_NEAR)) continue;
            const int ly = e / NW, lx = e - ly * NW;
            const int64_t id = (int64_t)(zp + NG) * g.plane + noff[k];
#pragma unroll
            for (int d = 0; d < 3; ++d) __pipeline_memcpy_async(&sns[slot][d][ly][lx], c.ns + d * V + id, 8);
        }
    };
    // STAGED: the solid normals were copied into shared memory one plane step ahead (ns_plane); otherwise they are read here
    auto normal_plane = [&](int zp, const uint8_t* cl, const bool staged) {
        const int slot = (zp + 12) % NS;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const int e = tid + k * NT;
            if (e >= NH * NW) break;
            const int ly = e / NW, lx = e - ly * NW;
            double G[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0}, gn = 0.0;
            const bool fluid = cl[k] & CLS_FLUID;
            const int64_t id = (int64_t)(zp + NG) * g.plane + noff[k];
            if (fluid) {
                // G = 3 sum_q w_q e_q phi(x + e_q) grouped as (1/6) * face differences + (1/12) * edge differences.
                // (3-D runs WettingType 2 only, whose 1e-8 threshold on |G| makes the exact-cancellation care of the
                // general path unnecessary: plain fused arithmetic here.)
                auto P = [&](int dx, int dy, int dz) { return sphi[(zp + dz + 10) % 5][ly + 1 + dy][lx + 1 + dx]; };
                const double pxy = P(1, 1, 0), mxy = P(-1, -1, 0), pmxy = P(1, -1, 0), mpxy = P(-1, 1, 0);
                const double pxz = P(1, 0, 1), mxz = P(-1, 0, -1), pmxz = P(1, 0, -1), mpxz = P(-1, 0, 1);
                const double pyz = P(0, 1, 1), myz = P(0, -1, -1), pmyz = P(0, 1, -1), mpyz = P(0, -1, 1);
                G[0] = (1.0 / 6.0) * (P(1, 0, 0) - P(-1, 0, 0)) +
                       (1.0 / 12.0) * (((pxy - mxy) + (pmxy - mpxy)) + ((pxz - mxz) + (pmxz - mpxz)));
                G[1] = (1.0 / 6.0) * (P(0, 1, 0) - P(0, -1, 0)) +
                       (1.0 / 12.0) * (((pxy - mxy) - (pmxy - mpxy)) + ((pyz - myz) + (pmyz - mpyz)));
                G[2] = (1.0 / 6.0) * (P(0, 0, 1) - P(0, 0, -1)) +
                       (1.0 / 12.0) * (((pxz - mxz) - (pmxz - mpxz)) + ((pyz - myz) - (pmyz - mpyz)));
                if (SOLIDS && (cl[k] & CLS_NEAR)) {
                    const int nsl = (zp + 8) & 1;
                    const double ns[3] = {staged ? sns[nsl][0][ly][lx] : c.ns[id], staged ? sns[nsl][1][ly][lx] : c.ns[V + id],
                                          staged ? sns[nsl][2][ly][lx] : c.ns[2 * V + id]};
                    if (c.p.exact_trig || c.p.wetting_type != 2) cg_wetting<3>(G, ns, c.p.cosT, c.p.sinT, c.p.wetting_type);
                    else cg_wetting_akai3_fast(G, ns, c.p.cosT, c.p.sinT);
                }
                const double g2 = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
                const double inv = g2 > 0.0 ? rsqrt(g2) : 0.0;
                gn = g2 * inv;
                const double sc = (c.p.wetting_type == 1 ? (gn > 0.0) : (gn > 1.0e-8)) ? (c.p.wetting_type == 1 ? inv : -inv) : 0.0;
                n[0] = sc * G[0]; n[1] = sc * G[1]; n[2] = sc * G[2];
            }
            sn[slot][0][ly][lx] = n[0]; sn[slot][1][ly][lx] = n[1]; sn[slot][2][ly][lx] = n[2];
            sn[slot][3][ly][lx] = gn;
        }
    };

    // x / y neighbour offsets of this thread's column (periodic)
    const int64_t xo[3] = {(int64_t)wrapx(x - 1), (int64_t)x, (int64_t)wrapx(x + 1)};
    const int64_t yo[3] = {(int64_t)wrapy(y - 1) * g.n0, (int64_t)y * g.n0, (int64_t)wrapy(y + 1) * g.n0};

    // use number of a slot: plane p is the ((p - (z_begin - 2)) / 5)-th occupant of its slot -> mbarrier parity
    auto wait_phi_plane = [&](int zp) {
        if (TMA) mbar_wait(&bars[(zp + 10) % 5], (uint32_t)(((zp - (z_begin - 2)) / 5) & 1));
    };
    if (TMA) {
        if (tid == 0) {
            for (int k = 0; k < 5; ++k) mbar_init(&bars[k], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    // Node classes of the normal tile run two plane steps ahead of their use, in registers (1 byte per element): at plane step z
    // cl1 = plane z + AHEAD (consumed by normal_plane in this step), cl2 = plane z + AHEAD + 1 (says which solid normals
    // ns_plane copies in this step), ncl = plane z + AHEAD + 2 (requested in this step).  Before: the class byte and the
    // normals were loaded where they were used, between the two barriers of a plane step -- 30 % of all stall samples of the
    // porous workload sat on those two loads.
    uint8_t ncl[NE], cl1[NE], cl2[NE];
    if (AHEAD == 1) {
        for (int zp = z_begin - 2; zp <= z_begin + 1; ++zp) load_phi_plane(zp);
        __pipeline_commit();
        load_phi_plane(z_begin + 2);
        class_plane(z_begin + 1, cl1);
        if (SOLIDS) ns_plane(z_begin + 1, cl1);
        __pipeline_commit();
        class_plane(z_begin + 2, cl2);
        __pipeline_wait_prior(1);
        for (int zp = z_begin - 2; zp <= z_begin + 1; ++zp) wait_phi_plane(zp);
        __syncthreads();
        class_plane(z_begin - 1, ncl);
        normal_plane(z_begin - 1, ncl, false);
        class_plane(z_begin, ncl);
        normal_plane(z_begin, ncl, false);
        // the first asynchronous copy of the loop lands in the slot of plane z_begin - 2, which normal_plane(z_begin - 1)
        // has just read: every thread must be done with it first (found as a run-to-run difference at 256^3)
        __syncthreads();
    } else {
        for (int zp = z_begin - 2; zp <= z_begin + 2; ++zp) load_phi_plane(zp);      // all five slots
        class_plane(z_begin + 2, cl1);
        if (SOLIDS && z_begin + 2 <= z_end) ns_plane(z_begin + 2, cl1);
        __pipeline_commit();
        class_plane(z_begin + 3, cl2);
        __pipeline_wait_prior(0);
        for (int zp = z_begin - 2; zp <= z_begin + 2; ++zp) wait_phi_plane(zp);
        __syncthreads();
        class_plane(z_begin - 1, ncl);
        normal_plane(z_begin - 1, ncl, false);
        __syncthreads();                                // every thread is done with phi(z_begin - 2): its slot takes plane z_begin + 3
        if (z_begin + 2 <= z_end) load_phi_plane(z_begin + 3);
        __pipeline_commit();
        class_plane(z_begin, ncl);
        normal_plane(z_begin, ncl, false);
        class_plane(z_begin + 1, ncl);
        normal_plane(z_begin + 1, ncl, false);
        __pipeline_wait_prior(0);
        if (z_begin + 2 <= z_end) wait_phi_plane(z_begin + 3);
        __syncthreads();
    }

    const double sgn = c.p.wetting_type == 1 ? 1.0 : -1.0;
    uint32_t pm_next = 1u;
    if (SOLIDS) pm_next = c.pull[(int64_t)(z_begin + NG) * g.plane + yo[1] + xo[1]];
    // PF: request the per-node inputs of plane zp into this thread's slots
    auto prefetch_inputs = [&](int zp) {
        const int64_t pid = (int64_t)(zp + NG) * g.plane + yo[1] + xo[1];
        __pipeline_memcpy_async(&spop[tid], s.gT + pid, 8);
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = (int64_t)(zp - L::d2(q) + NG) * g.plane + yo[1 - L::d1(q)] + xo[1 - L::d0(q)];
            __pipeline_memcpy_async(&spop[q * NT + tid], s.gT + q * V + src, 8);
        }
        __pipeline_memcpy_async(&spop[19 * NT + tid], c.rho[0] + pid, 8);
        __pipeline_memcpy_async(&spop[20 * NT + tid], c.rho[1] + pid, 8);
#pragma unroll
        for (int d = 0; d < 3; ++d) __pipeline_memcpy_async(&spop[(21 + d) * NT + tid], c.F + d * V + pid, 8);
    };
    if (PF) { prefetch_inputs(z_begin); __pipeline_commit(); }
    for (int z = z_begin; z < z_end; ++z) {
        double fT[L::Q];
        double rR = 1.0, rB = 1.0, Fl[3] = {0.0, 0.0, 0.0}, phi0 = 0.0;
        if (PF) {
            __pipeline_wait_prior(0);                   // this thread's inputs of plane z (and every older copy) have landed
#pragma unroll
            for (int q = 0; q < L::Q; ++q) fT[q] = spop[q * NT + tid];
            rR = spop[19 * NT + tid]; rB = spop[20 * NT + tid];
#pragma unroll
            for (int d = 0; d < 3; ++d) Fl[d] = spop[(21 + d) * NT + tid];
#ifndef LBM_HOSTCHECK
            // the values are in registers before the slots are handed to the next copies
            asm volatile("" : "+d"(fT[0]), "+d"(fT[1]), "+d"(fT[2]), "+d"(fT[3]), "+d"(fT[4]), "+d"(fT[5]), "+d"(fT[6]),
                              "+d"(fT[7]), "+d"(fT[8]), "+d"(fT[9]), "+d"(fT[10]), "+d"(fT[11]), "+d"(fT[12]), "+d"(fT[13]),
                              "+d"(fT[14]), "+d"(fT[15]), "+d"(fT[16]), "+d"(fT[17]), "+d"(fT[18]), "+d"(rR), "+d"(rB),
                              "+d"(Fl[0]), "+d"(Fl[1]), "+d"(Fl[2]));
#endif
            if (z + 1 < z_end) prefetch_inputs(z + 1);
            __pipeline_commit();
        }
        const bool more_phi = z + AHEAD + 1 <= z_end;   // the next step derives the normals of plane z + AHEAD + 1
        if (more_phi) load_phi_plane(z + AHEAD + 2);
        if (SOLIDS && more_phi) ns_plane(z + AHEAD + 1, cl2);
        __pipeline_commit();
        // ---- requests to HBM first: pulled populations, densities, lagged force ----
        const int64_t id = (int64_t)(z + NG) * g.plane + yo[1] + xo[1];
        // pull mask (grid.cuh::PullMaskOp): arrived one plane step ago, the next one is requested now
        uint32_t pm = 0xFFFFFFFFu;
        if (SOLIDS) {
            pm = pm_next;
            if (z + 1 < z_end) pm_next = c.pull[id + g.plane];
        }
        const bool fluid = pm & 1u;
        if (z + AHEAD + 2 <= z_end) class_plane(z + AHEAD + 2, ncl);     // node classes of the normal tile, two plane steps ahead
        if (fluid && !PF) {
            fT[0] = __ldcs(s.gT + id);
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const int64_t src = (int64_t)(z - L::d2(q) + NG) * g.plane + yo[1 - L::d1(q)] + xo[1 - L::d0(q)];
                int64_t addr = q * V + src;
                if (SOLIDS && !(pm & (1u << q))) addr = L::opp(q) * V + id;      // half-way bounce back
                fT[q] = __ldcs(s.gT + addr);
            }
            rR = c.rho[0][id]; rB = c.rho[1][id];
#pragma unroll
            for (int d = 0; d < 3; ++d) Fl[d] = c.F[d * V + id];
        }
        if (AHEAD == 1) {
            if (!PF) __pipeline_wait_prior(1);          // plane z + 2 (requested one step ago) has landed (PF: waited for at the top)
            wait_phi_plane(z + 2);
            __syncthreads();
            normal_plane(z + 1, cl1, true);
            __syncthreads();
        } else if (z + AHEAD <= z_end) {
            normal_plane(z + AHEAD, cl1, true);         // reads phi up to plane z + AHEAD + 1: landed and published at the end of the last step
        }
#pragma unroll
        for (int k = 0; k < NE; ++k) { cl1[k] = cl2[k]; cl2[k] = ncl[k]; }
        if (fluid) {
        if (SOLIDS) {
            // The requests above sit in a conditional block; without this fence the compiler sinks the first
            // arithmetic on the loaded values (0.5 * F, the first moment sums) into that block, i.e. in FRONT of the
            // two barriers, and every warp then waits for HBM before the shared-memory phase instead of during it
            // (ncu, porous 256 x 256 x 192: 19.5 % of all stall samples on that one DMUL).  The empty asm makes the
            // values opaque until here.
#ifndef LBM_HOSTCHECK
            asm volatile("" : "+d"(fT[0]), "+d"(fT[1]), "+d"(fT[2]), "+d"(fT[3]), "+d"(fT[4]), "+d"(fT[5]), "+d"(fT[6]),
                              "+d"(fT[7]), "+d"(fT[8]), "+d"(fT[9]), "+d"(fT[10]), "+d"(fT[11]), "+d"(fT[12]), "+d"(fT[13]),
                              "+d"(fT[14]), "+d"(fT[15]), "+d"(fT[16]), "+d"(fT[17]), "+d"(fT[18]), "+d"(rR), "+d"(rB),
                              "+d"(Fl[0]), "+d"(Fl[1]), "+d"(Fl[2]));
#endif
        }
        phi0 = sphi[(z + 10) % 5][ty + 2][tx + 2];
        // ---- curvature and force from the normals in shared memory ----
        const int sl = (z + 12) % NS;
        double n[3] = {sn[sl][0][ty + 1][tx + 1], sn[sl][1][ty + 1][tx + 1], sn[sl][2][ty + 1][tx + 1]};
        const double gn = sn[sl][3][ty + 1][tx + 1];
        double dn[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int sq = (z + L::d2(q) + 12) % NS;
            double nk[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) nk[b] = sn[sq][b][ty + 1 + L::d1(q)][tx + 1 + L::d0(q)];
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (L::c(q, a) != 0)
#pragma unroll
                    for (int b = 0; b < 3; ++b) dn[a][b] += (3.0 * L::w(q) * L::c(q, a)) * nk[b];
        }
        double K = 0.0, nn = 0.0, div = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            nn += n[a] * n[a]; div += dn[a][a];
#pragma unroll
            for (int b = 0; b < 3; ++b) K += n[a] * n[b] * dn[a][b];
        }
        K -= nn * div;
        double G[3], F[3], u[3];
        const double rho = rB + rR, irho = 1.0 / rho;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            G[d] = sgn * gn * n[d];
            F[d] = (c.p.wetting_type == 1 ? 0.5 : -0.5) * c.p.sigma * K * G[d];
        }
        const double tau = c.p.tauR == c.p.tauB ? c.p.tauR : cg_tau(phi0, rR, rB, c.p);   // equal viscosities: tau(phi) is constant
        if (c.p.relax != 0) {
            // MRT: the momentum is three of the moments, so the transform is done once
            double m[L::NMOM];
            L::to_moments(fT, m);
            u[0] = (m[3] + 0.5 * Fl[0]) * irho; u[1] = (m[5] + 0.5 * Fl[1]) * irho; u[2] = (m[7] + 0.5 * Fl[2]) * irho;
            L::relax_moments(m, rho, u, F, 1.0 / tau);
            L::from_moments(m, fT);
        } else {
            double mom[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 1; q < L::Q; ++q)
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if (L::c(q, d) != 0) mom[d] += L::c(q, d) * fT[q];
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] = (mom[d] + 0.5 * Fl[d]) * irho;
            cg_collide<L>(fT, rho, u, F, tau, 0);
        }
        const double amp = gn > 1.0e-8 ? c.p.beta * rR * rB * irho : 0.0;   // a = amp G / |G| = amp sgn n
#pragma unroll
        for (int q = 0; q < L::Q; ++q) __stcs(o.gT + q * V + id, fT[q]);
        o.kR[id] = rR * irho;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            o.a[d * V + id] = amp * sgn * n[d];
            c.F[d * V + id] = F[d];
        }
        if (PEER) {
            // the factored state travels one plane deep: populations moving up into the low ghost of the slab above, those moving
            // down into the high ghost of the slab below, the four recolouring scalars both ways (factored_dirs)
            if (z == g.n2 - 1 && pp.up) {
                const int64_t gid = id - (int64_t)g.n2 * g.plane;
#pragma unroll
                for (int q = 1; q < L::Q; ++q)
                    if (L::d2(q) == 1) pp.up[q * V + gid] = fT[q];
                pp.up[L::Q * V + gid] = rR * irho;
#pragma unroll
                for (int d = 0; d < 3; ++d) pp.up[(L::Q + 1 + d) * V + gid] = amp * sgn * n[d];
            }
            if (z == 0 && pp.down) {
                const int64_t gid = id + (int64_t)g.n2 * g.plane;
#pragma unroll
                for (int q = 1; q < L::Q; ++q)
                    if (L::d2(q) == -1) pp.down[q * V + gid] = fT[q];
                pp.down[L::Q * V + gid] = rR * irho;
#pragma unroll
                for (int d = 0; d < 3; ++d) pp.down[(L::Q + 1 + d) * V + gid] = amp * sgn * n[d];
            }
        }
        }   // fluid
        if (AHEAD != 1) {
            // the copies requested at the top of this step (phi plane, solid normals) have had the whole step to land; after
            // the barrier they are visible to every thread, and everybody is done with the slots the next step overwrites
            __pipeline_wait_prior(0);
            if (more_phi) wait_phi_plane(z + AHEAD + 2);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled density pass for D3Q19 (pass 1).  Same column-marching layout as the collision pass: the four
// recolouring scalars (kR, a) of the neighbours are served from a rolling 5-plane shared-memory window
// with a 1-node halo (each value is read from L2/HBM once per CTA instead of 19 times), and the 19
// pulled populations of the NEXT plane are requested before the current plane is reduced, so one full
// plane of HBM requests per thread is always in flight.
// ------------------------------------------------------------------------------------------------
// TMA = true: the scalar planes are staged with TMA bulk copies (rows widened to a 2-node halo in x so that they start
// 16-byte aligned), issued by four threads (one per field) and completing on a per-slot mbarrier.
template <bool SOLIDS, int TX, int TY, bool TMA, bool PEER = false>
__global__ void __launch_bounds__(TX* TY, 512 / (TX * TY) > 0 ? 512 / (TX * TY) : 1)
cg_density_tiled_d3q19(const CGFields c, const FastFields s, const int zchunk, const int z_lo, const int z_hi,
                       const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    using L = D3Q19;
    constexpr int NT = TX * TY;
    constexpr int XH = TMA ? 2 : 1;                 // halo columns kept in shared memory
    constexpr int NW = TX + 2 * XH, NH = TY + 2;
    LBM_DYN_SMEM(smem_dyn);
    double (*ss)[4][NH][NW] = reinterpret_cast<double (*)[4][NH][NW]>(smem_dyn);   // [5 slots][kR, ax, ay, az]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + 5 * 4 * NH * NW);      // [5] one mbarrier per slot
    const Grid& g = c.g;
    const int64_t V = g.vol;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = z_lo + blockIdx.z * zchunk;
    const int z_end = min(z_begin + zchunk, z_hi);
    const int x = x0 + tx, y = y0 + ty;
    auto wrapx = [&](int v) { return v < 0 ? v + g.n0 : (v >= g.n0 ? v - g.n0 : v); };
    auto wrapy = [&](int v) { return v < 0 ? v + g.n1 : (v >= g.n1 ? v - g.n1 : v); };
    const double* fld[4] = {s.kR, s.a, s.a + V, s.a + 2 * V};

    auto load_scalar_plane = [&](int zp) {
        const int slot = (zp + 10) % 5;
        const int64_t base = (int64_t)(zp + NG) * g.plane;
        if (TMA) {
            if (tid < 32) {
                if (tid == 0) mbar_arrive_expect_tx(&bars[slot], 4 * NH * NW * 8);
                __syncwarp();
                if (tid < 4) {
                    const int c0 = x0 - 2, cs = c0 < 0 ? 0 : c0, ce = x0 + TX + 2 > g.n0 ? g.n0 : x0 + TX + 2;
                    for (int ly = 0; ly < NH; ++ly) {
                        const double* row = fld[tid] + base + (int64_t)wrapy(y0 + ly - 1) * g.n0;
                        double* dst = &ss[slot][tid][ly][0];
                        if (c0 < 0) bulk_g2s(dst, row + g.n0 - 2, 16, &bars[slot]);
                        bulk_g2s(dst + (cs - c0), row + cs, (uint32_t)(ce - cs) * 8, &bars[slot]);
                        if (x0 + TX + 2 > g.n0) bulk_g2s(dst + NW - 2, row, 16, &bars[slot]);
                    }
                }
            }
            return;
        }
        for (int e = tid; e < NH * NW; e += NT) {
            const int ly = e / NW, lx = e - ly * NW;
            const int64_t off = base + (int64_t)wrapy(y0 + ly - 1) * g.n0 + wrapx(x0 + lx - XH);
#pragma unroll
            for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(&ss[slot][k][ly][lx], fld[k] + off, 8);
        }
    };
    auto wait_scalar_plane = [&](int zp, int first) {
        if (TMA) mbar_wait(&bars[(zp + 10) % 5], (uint32_t)(((zp - first) / 5) & 1));
    };
    const int64_t xo[3] = {(int64_t)wrapx(x - 1), (int64_t)x, (int64_t)wrapx(x + 1)};
    const int64_t yo[3] = {(int64_t)wrapy(y - 1) * g.n0, (int64_t)y * g.n0, (int64_t)wrapy(y + 1) * g.n0};

    // request the pulled populations of plane z: values + bit mask of the directions whose upstream node is fluid
    // `pm` = the node's pull mask (grid.cuh::PullMaskOp), requested one plane step before the populations it steers
    auto request = [&](int z, uint32_t pm, double* f, unsigned& mask, bool& fluid) {
        const int64_t id = (int64_t)(z + NG) * g.plane + yo[1] + xo[1];
        mask = SOLIDS ? pm : 0xFFFFFFFFu;
        fluid = mask & 1u;
        if (!fluid) return;
        f[0] = __ldcs(s.gT + id);
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = (int64_t)(z - L::d2(q) + NG) * g.plane + yo[1 - L::d1(q)] + xo[1 - L::d0(q)];
            int64_t addr = q * V + src;
            if (SOLIDS && !(mask & (1u << q))) addr = L::opp(q) * V + id;        // half-way bounce back
            f[q] = __ldcs(s.gT + addr);
        }
    };
    auto pull_mask = [&](int z) -> uint32_t {
        return (SOLIDS && z < z_end) ? c.pull[(int64_t)(z + NG) * g.plane + yo[1] + xo[1]] : 0u;
    };

    if (TMA) {
        if (tid == 0) {
            for (int k = 0; k < 5; ++k) mbar_init(&bars[k], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    load_scalar_plane(z_begin - 1);
    load_scalar_plane(z_begin);
    __pipeline_commit();
    load_scalar_plane(z_begin + 1);
    __pipeline_commit();
    wait_scalar_plane(z_begin - 1, z_begin - 1);
    wait_scalar_plane(z_begin, z_begin - 1);
    const bool pert = c.p.st_type == LBM_ST_PERTURBATION;      // recolouring weights w_i / |e_i| (cg_fast_ops.cuh::cg_red_part)
    double cur[L::Q], nxt[L::Q];
    unsigned mcur = 0, mnxt = 0;
    bool fcur = true, fnxt = true;
    // the pull masks run TWO plane steps ahead of the requests they steer (a plane step of this pass is short: one step ahead
    // left 15 % of all stall samples of the porous workload on the mask's arrival)
    uint32_t pm1 = pull_mask(z_begin + 1), pm2 = pull_mask(z_begin + 2), pm3 = 0u;
    request(z_begin, pull_mask(z_begin), cur, mcur, fcur);
    for (int z = z_begin; z < z_end; ++z) {
        if (z + 1 < z_end) load_scalar_plane(z + 2);    // cp.async, needed by the next plane step
        __pipeline_commit();
        pm3 = pull_mask(z + 3);
        if (z + 1 < z_end) request(z + 1, pm1, nxt, mnxt, fnxt);
        pm1 = pm2; pm2 = pm3;
        __pipeline_wait_prior(1);                       // plane z + 1 has landed
        wait_scalar_plane(z + 1, z_begin - 1);
        __syncthreads();
        if (fcur) {
            const int64_t id = (int64_t)(z + NG) * g.plane + yo[1] + xo[1];
            const int s0 = (z + 10) % 5;
            const double kR0 = ss[s0][0][ty + 1][tx + XH];
            const double a0[3] = {ss[s0][1][ty + 1][tx + XH], ss[s0][2][ty + 1][tx + XH], ss[s0][3][ty + 1][tx + XH]};
            double accR = kR0 * cur[0], accB = cur[0] - accR;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                double fr;
                if (!SOLIDS || (mcur & (1u << q))) {
                    const int sq = (z - L::d2(q) + 10) % 5;
                    const int ly = ty + 1 - L::d1(q), lx = tx + XH - L::d0(q);
                    double ea = 0.0;
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        if (L::c(q, d) != 0) ea += L::c(q, d) * ss[sq][1 + d][ly][lx];
                    fr = ss[sq][0][ly][lx] * cur[q] + cg_red_weight<L>(q, pert) * ea;
                } else {
                    fr = cg_red_part<L>(L::opp(q), cur[q], kR0, a0, pert);
                }
                accR += fr; accB += cur[q] - fr;
            }
            c.rho[0][id] = accR; c.rho[1][id] = accB;
            c.phi[id] = (accR - accB) / (accR + accB);
            if (PEER) {       // phi travels pp.gp planes deep, both ways
                if (z >= g.n2 - pp.gp && pp.up) pp.up[id - (int64_t)g.n2 * g.plane] = (accR - accB) / (accR + accB);
                if (z < pp.gp && pp.down) pp.down[id + (int64_t)g.n2 * g.plane] = (accR - accB) / (accR + accB);
            }
        }
#pragma unroll
        for (int q = 0; q < L::Q; ++q) cur[q] = nxt[q];
        mcur = mnxt; fcur = fnxt;
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled collision pass of the PERTURBATION-operator model for D3Q19 (what the reference's RKtwophasesetup3D.ini selects).
// Same column-marching layout as cg_collide_tiled_d3q19, but this model needs no normals and no curvature: the colour gradient
// of a node comes straight from a rolling window of phi planes (halo 1, four slots, 8-byte cp.async one plane step ahead), so
// a plane step has ONE barrier and the only shared-memory traffic is the 18-point phi stencil.  Solid neighbours show
// SolidColorDiff; which neighbours are solid is read off the node's pull mask (bit opp(q) = "x + e_q is fluid").  The node
// arithmetic is cg_fast_ops.cuh::cgp_collide_factored, the gradient that of cgp_gradient_at operation for operation.
// ------------------------------------------------------------------------------------------------
// PF (LBM_PERT_PREFETCH): the 21 per-node inputs of plane z + 1 (19 pulled populations, rhoR, rhoB) are requested during step z with
// 8-byte cp.async into thread-private shared-memory slots -- no registers, no barrier -- and picked up at the top of step z + 1.  The
// same idea lost on the CSF kernel (its ~95 shared-memory instructions per node already saturate the MIO queue); this kernel issues
// ~20, and its profile shows the plane's loads landing on the first use of the populations (long-scoreboard at `rho = rB + rR`).
template <bool SOLIDS, int TX, int TY, bool PEER = false, bool PF = false>
__global__ void __launch_bounds__(TX* TY, 512 / (TX * TY) > 0 ? 512 / (TX * TY) : 1)
cgp_collide_tiled_d3q19(const CGFields c, const FastFields s, const FastFields o, const int zchunk, const int z_lo, const int z_hi,
                        const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    using L = D3Q19;
    constexpr int NT = TX * TY, NW = TX + 2, NH = TY + 2;
    LBM_DYN_SMEM(smem_dyn);
    double (*sphi)[NH][NW] = reinterpret_cast<double (*)[NH][NW]>(smem_dyn);      // [4] planes
    double* spop = smem_dyn + 4 * NH * NW;                                        // PF: [21][NT] thread-private slots
    const Grid& g = c.g;
    const int64_t V = g.vol;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = z_lo + blockIdx.z * zchunk;
    const int z_end = min(z_begin + zchunk, z_hi);
    const int x = x0 + tx, y = y0 + ty;
    auto wrapx = [&](int v) { return v < 0 ? v + g.n0 : (v >= g.n0 ? v - g.n0 : v); };
    auto wrapy = [&](int v) { return v < 0 ? v + g.n1 : (v >= g.n1 ? v - g.n1 : v); };
    auto load_phi_plane = [&](int zp) {
        const int slot = (zp + 8) & 3;
        const double* src = c.phi + (int64_t)(zp + NG) * g.plane;
        for (int e = tid; e < NH * NW; e += NT) {
            const int ly = e / NW, lx = e - ly * NW;
            __pipeline_memcpy_async(&sphi[slot][ly][lx], src + (int64_t)wrapy(y0 + ly - 1) * g.n0 + wrapx(x0 + lx - 1), 8);
        }
    };
    const int64_t xo[3] = {(int64_t)wrapx(x - 1), (int64_t)x, (int64_t)wrapx(x + 1)};
    const int64_t yo[3] = {(int64_t)wrapy(y - 1) * g.n0, (int64_t)y * g.n0, (int64_t)wrapy(y + 1) * g.n0};
    for (int zp = z_begin - 1; zp <= z_begin + 1; ++zp) load_phi_plane(zp);
    uint32_t pm_next = 0xFFFFFFFFu, pm_next2 = 0xFFFFFFFFu;      // PF: the masks run two planes ahead (they steer the requests of z + 1)
    if (SOLIDS) {
        pm_next = c.pull[(int64_t)(z_begin + NG) * g.plane + yo[1] + xo[1]];
        if (PF && z_begin + 1 < z_end) pm_next2 = c.pull[(int64_t)(z_begin + 1 + NG) * g.plane + yo[1] + xo[1]];
    }
    // PF: request the per-node inputs of plane zp (pull mask pmz) into this thread's slots
    auto prefetch_inputs = [&](int zp, uint32_t pmz) {
        if (!(pmz & 1u)) return;
        const int64_t pid = (int64_t)(zp + NG) * g.plane + yo[1] + xo[1];
        __pipeline_memcpy_async(&spop[tid], s.gT + pid, 8);
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = (int64_t)(zp - L::d2(q) + NG) * g.plane + yo[1 - L::d1(q)] + xo[1 - L::d0(q)];
            int64_t addr = q * V + src;
            if (SOLIDS && !(pmz & (1u << q))) addr = L::opp(q) * V + pid;        // half-way bounce back
            __pipeline_memcpy_async(&spop[q * NT + tid], s.gT + addr, 8);
        }
        __pipeline_memcpy_async(&spop[19 * NT + tid], c.rho[0] + pid, 8);
        __pipeline_memcpy_async(&spop[20 * NT + tid], c.rho[1] + pid, 8);
    };
    if (PF) prefetch_inputs(z_begin, pm_next);
    __pipeline_commit();
    for (int z = z_begin; z < z_end; ++z) {
        const int64_t id = (int64_t)(z + NG) * g.plane + yo[1] + xo[1];
        uint32_t pm = 0xFFFFFFFFu;
        if (SOLIDS) {
            pm = pm_next;
            if (PF) {
                pm_next = pm_next2;
                if (z + 2 < z_end) pm_next2 = c.pull[id + 2 * g.plane];
            } else if (z + 1 < z_end) pm_next = c.pull[id + g.plane];
        }
        const bool fluid = pm & 1u;
        double fT[L::Q];
        double rR = 1.0, rB = 1.0;
        if (!PF && fluid) {
            fT[0] = __ldcs(s.gT + id);
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const int64_t src = (int64_t)(z - L::d2(q) + NG) * g.plane + yo[1 - L::d1(q)] + xo[1 - L::d0(q)];
                int64_t addr = q * V + src;
                if (SOLIDS && !(pm & (1u << q))) addr = L::opp(q) * V + id;      // half-way bounce back
                fT[q] = __ldcs(s.gT + addr);
            }
            rR = c.rho[0][id]; rB = c.rho[1][id];
        }
        __pipeline_wait_prior(0);           // phi plane z + 1 (and, PF, this thread's inputs of plane z), requested during the last step
        if (PF && fluid) {
#pragma unroll
            for (int q = 0; q < L::Q; ++q) fT[q] = spop[q * NT + tid];
            rR = spop[19 * NT + tid]; rB = spop[20 * NT + tid];
#ifndef LBM_HOSTCHECK
            // the values are in registers before the slots are handed to the next copies
            asm volatile("" : "+d"(fT[0]), "+d"(fT[1]), "+d"(fT[2]), "+d"(fT[3]), "+d"(fT[4]), "+d"(fT[5]), "+d"(fT[6]),
                              "+d"(fT[7]), "+d"(fT[8]), "+d"(fT[9]), "+d"(fT[10]), "+d"(fT[11]), "+d"(fT[12]), "+d"(fT[13]),
                              "+d"(fT[14]), "+d"(fT[15]), "+d"(fT[16]), "+d"(fT[17]), "+d"(fT[18]), "+d"(rR), "+d"(rB));
#endif
        }
        __syncthreads();
        // the slot of plane z - 2 is free now: every thread has left step z - 1, the last reader of that plane
        if (z + 1 < z_end) load_phi_plane(z + 2);
        if (PF && z + 1 < z_end) prefetch_inputs(z + 1, pm_next);
        __pipeline_commit();
        if (!fluid) continue;
        double G[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            double pk = sphi[(z + L::d2(q) + 8) & 3][ty + 1 + L::d1(q)][tx + 1 + L::d0(q)];
            if (SOLIDS && !(pm & (1u << L::opp(q)))) pk = c.p.solid_phi;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (L::c(q, k) != 0) G[k] = add_rn(G[k], mul_rn(3.0 * L::w(q) * L::c(q, k), pk));
        }
        const double phi0 = sphi[(z + 8) & 3][ty + 1][tx + 1];
        const double rho = rB + rR;
        double mom[3] = {0.0, 0.0, 0.0}, u[3], kR, a[3];
#pragma unroll
        for (int q = 1; q < L::Q; ++q)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (L::c(q, k) != 0) mom[k] += L::c(q, k) * fT[q];
#pragma unroll
        for (int k = 0; k < 3; ++k) u[k] = mom[k] / rho;
        cgp_collide_factored<L>(fT, rR, rB, phi0, u, G, c.p, &kR, a);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) __stcs(o.gT + q * V + id, fT[q]);
        o.kR[id] = kR;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o.a[k * V + id] = a[k]; c.G[k * V + id] = G[k]; }
        if (PEER) {         // the slab's first / last plane also goes into the neighbour slabs' ghost planes (see cg_collide_tiled_d3q19)
            if (z == g.n2 - 1 && pp.up) {
                const int64_t gid = id - (int64_t)g.n2 * g.plane;
#pragma unroll
                for (int q = 1; q < L::Q; ++q)
                    if (L::d2(q) == 1) pp.up[q * V + gid] = fT[q];
                pp.up[L::Q * V + gid] = kR;
#pragma unroll
                for (int k = 0; k < 3; ++k) pp.up[(L::Q + 1 + k) * V + gid] = a[k];
            }
            if (z == 0 && pp.down) {
                const int64_t gid = id + (int64_t)g.n2 * g.plane;
#pragma unroll
                for (int q = 1; q < L::Q; ++q)
                    if (L::d2(q) == -1) pp.down[q * V + gid] = fT[q];
                pp.down[L::Q * V + gid] = kR;
#pragma unroll
                for (int k = 0; k < 3; ++k) pp.down[(L::Q + 1 + k) * V + gid] = a[k];
            }
        }
    }
}

constexpr int TILE_X = 32;
// rows of a tile (tuning knobs, read once): the collision pass runs best with 32 x 4 tiles (4 CTAs of 128 threads
// per SM: their barriers interleave; 512^3: 10.1 ms vs 10.7 ms with 32 x 8 and 12.4 ms with 32 x 16), the density
// pass with 32 x 8 (4.8 vs 5.0 ms).  LBM_TILE_Y_COLLIDE / LBM_TILE_Y_DENSITY = 4 | 8 | 16, LBM_ZCHUNK = planes per CTA.
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int tile_y_collide() {
    static const int ty = [] { const int v = env_int("LBM_TILE_Y_COLLIDE", 4); return (v != 4 && v != 8 && v != 16) ? 4 : v; }();
    return ty;
}
static int tile_y_density() {
    static const int ty = [] { const int v = env_int("LBM_TILE_Y_DENSITY", 8); return (v != 4 && v != 8 && v != 16) ? 8 : v; }();
    return ty;
}
static int z_chunk(int n2) {
    static const int zc = [] { const int v = env_int("LBM_ZCHUNK", 32); return v < 4 ? 32 : v; }();
    return n2 >= 2 * zc ? zc : n2;
}
static bool tiled_ok(const lbm_handle* h) {
    return h->Q == 19 && h->g.n0 % TILE_X == 0 && h->g.n1 % tile_y_collide() == 0 && h->g.n1 % tile_y_density() == 0 && !(h->cfg.flags & 2u);
}

bool cg_tiled_possible(const lbm_handle* h) { return h->cfg.model == LBM_MODEL_CG && tiled_ok(h); }

template <bool SOLIDS, int TILE_Y, bool TMA, bool PEER, int AHEAD, bool PF = false>
static void launch_tiled_a(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o, int z_lo, int z_hi, const PeerPtrs pp) {
    const Grid& g = h->g;
    if (z_hi < 0) z_hi = g.n2;
    if (z_hi <= z_lo) return;
    const int zchunk = z_chunk(g.n2);
    dim3 grid(g.n0 / TILE_X, g.n1 / TILE_Y, (z_hi - z_lo + zchunk - 1) / zchunk), block(TILE_X, TILE_Y);
    constexpr size_t smem = sizeof(double) * (5 * (TILE_Y + 4) * (TILE_X + 4) + (AHEAD + 2) * 4 * (TILE_Y + 2) * (TILE_X + 2) + 8 +
                                              (SOLIDS ? 2 * 3 * (TILE_Y + 2) * (TILE_X + 2) : 0) + (PF ? 24 * TILE_X * TILE_Y : 0));
#ifdef LBM_HOSTCHECK
    cta_emu::launch(grid, block, smem, [&] { cg_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER, AHEAD, PF>(c, s, o, zchunk, z_lo, z_hi, pp); });
#else
    static std::atomic<bool> configured[64];  // per device: the attribute belongs to the function on ONE device (zero-initialised)
    if (!configured[h->cfg.device & 63]) {
        LBM_CUDA_CHECK(cudaFuncSetAttribute(cg_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER, AHEAD, PF>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device & 63] = true;
    }
    if (g_prof.on) g_prof.begin(SOLIDS ? "cg_collide_tiled_d3q19<solids>" : "cg_collide_tiled_d3q19<all-fluid>", h->stream);
    cg_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER, AHEAD, PF><<<grid, block, smem, h->stream>>>(c, s, o, zchunk, z_lo, z_hi, pp);
    if (g_prof.on) g_prof.end(h->stream);
    LBM_CUDA_CHECK(cudaGetLastError());
#endif
    ++g_launch_counter;
}

// normals one plane ahead (two barriers per plane step) or two planes ahead (one barrier): LBM_COLLIDE_AHEAD = 1 | 2
static int collide_ahead() {
    static const int a = [] { const int v = env_int("LBM_COLLIDE_AHEAD", 1); return v == 2 ? 2 : 1; }();
    return a;
}
template <bool SOLIDS, int TILE_Y, bool TMA, bool PEER = false>
static void launch_tiled_t(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o, int z_lo, int z_hi,
                           const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    static const bool pf = env_int("LBM_POP_PREFETCH", 0) != 0;
    if (!SOLIDS && TILE_Y == 4 && pf && collide_ahead() == 1) { launch_tiled_a<false, 4, TMA, PEER, 1, true>(h, c, s, o, z_lo, z_hi, pp); return; }
    if (collide_ahead() == 1) launch_tiled_a<SOLIDS, TILE_Y, TMA, PEER, 1>(h, c, s, o, z_lo, z_hi, pp);
    else launch_tiled_a<SOLIDS, TILE_Y, TMA, PEER, 2>(h, c, s, o, z_lo, z_hi, pp);
}

template <bool SOLIDS>
static void launch_tiled(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o, int z_lo = 0, int z_hi = -1) {
#ifdef LBM_HOSTCHECK
    static const bool tma = false;                               // host emulation: the copies are plain memcpy
#else
    static const bool tma = env_int("LBM_PHI_TMA", 1) != 0;      // 0: 8-byte cp.async copies instead of TMA bulk copies
#endif
    switch (tile_y_collide()) {
        case 4: if (tma) launch_tiled_t<SOLIDS, 4, true>(h, c, s, o, z_lo, z_hi); else launch_tiled_t<SOLIDS, 4, false>(h, c, s, o, z_lo, z_hi); break;
        case 16: launch_tiled_t<SOLIDS, 16, false>(h, c, s, o, z_lo, z_hi); break;
        default: if (tma) launch_tiled_t<SOLIDS, 8, true>(h, c, s, o, z_lo, z_hi); else launch_tiled_t<SOLIDS, 8, false>(h, c, s, o, z_lo, z_hi);
    }
}

template <bool SOLIDS, bool PEER = false>
static void launch_perturb_tiled(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o,
                                 const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    const Grid& g = h->g;
    constexpr int TILE_Y = 4;
    const int zchunk = z_chunk(g.n2);
    dim3 grid(g.n0 / TILE_X, g.n1 / TILE_Y, (g.n2 + zchunk - 1) / zchunk), block(TILE_X, TILE_Y);
    // measured on the B200 (profiles/r02b_ini3d_*): 1.567 -> 1.432 ms per 256^3 launch, 7 174 -> 7 614 MLUPS for the reference's 3-D ini
    // configuration (8.0 -> 8.6 GLUPS = 0.82 of the roofline at 256 x 256 x 512); LBM_PERT_PREFETCH=0 selects the plain loads
    static const bool pf = env_int("LBM_PERT_PREFETCH", 1) != 0;
    const size_t smem = sizeof(double) * (4 * (TILE_Y + 2) * (TILE_X + 2) + (pf ? 21 * TILE_X * TILE_Y : 0));
#ifdef LBM_HOSTCHECK
    if (pf) cta_emu::launch(grid, block, smem, [&] { cgp_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, PEER, true>(c, s, o, zchunk, 0, g.n2, pp); });
    else cta_emu::launch(grid, block, smem, [&] { cgp_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, PEER>(c, s, o, zchunk, 0, g.n2, pp); });
#else
    if (g_prof.on) g_prof.begin(SOLIDS ? "cgp_collide_tiled_d3q19<solids>" : "cgp_collide_tiled_d3q19<all-fluid>", h->stream);
    if (pf) cgp_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, PEER, true><<<grid, block, smem, h->stream>>>(c, s, o, zchunk, 0, g.n2, pp);
    else cgp_collide_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, PEER><<<grid, block, smem, h->stream>>>(c, s, o, zchunk, 0, g.n2, pp);
    if (g_prof.on) g_prof.end(h->stream);
    LBM_CUDA_CHECK(cudaGetLastError());
#endif
    ++g_launch_counter;
}

template <bool SOLIDS, int TILE_Y, bool TMA, bool PEER = false>
static void launch_density_tiled_t(lbm_handle* h, const CGFields& c, const FastFields& s, int z_lo, int z_hi,
                                   const PeerPtrs pp = PeerPtrs{nullptr, nullptr, 0}) {
    const Grid& g = h->g;
    if (z_hi < 0) z_hi = g.n2;
    if (z_hi <= z_lo) return;
    const int zchunk = z_chunk(g.n2);
    dim3 grid(g.n0 / TILE_X, g.n1 / TILE_Y, (z_hi - z_lo + zchunk - 1) / zchunk), block(TILE_X, TILE_Y);
    constexpr size_t smem = sizeof(double) * (5 * 4 * (TILE_Y + 2) * (TILE_X + (TMA ? 4 : 2)) + 8);
#ifdef LBM_HOSTCHECK
    cta_emu::launch(grid, block, smem, [&] { cg_density_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER>(c, s, zchunk, z_lo, z_hi, pp); });
#else
    static std::atomic<bool> configured[64];
    if (!configured[h->cfg.device & 63]) {
        LBM_CUDA_CHECK(cudaFuncSetAttribute(cg_density_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[h->cfg.device & 63] = true;
    }
    if (g_prof.on) g_prof.begin(SOLIDS ? "cg_density_tiled_d3q19<solids>" : "cg_density_tiled_d3q19<all-fluid>", h->stream);
    cg_density_tiled_d3q19<SOLIDS, TILE_X, TILE_Y, TMA, PEER><<<grid, block, smem, h->stream>>>(c, s, zchunk, z_lo, z_hi, pp);
    if (g_prof.on) g_prof.end(h->stream);
    LBM_CUDA_CHECK(cudaGetLastError());
#endif
    ++g_launch_counter;
}
// the variants that also store into the neighbour slabs exist for the default tile shapes only
static bool peer_tiles_default() { return tile_y_collide() == 4 && tile_y_density() == 8; }
template <bool SOLIDS>
static void launch_tiled_peer(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o, const PeerPtrs pp) {
#ifdef LBM_HOSTCHECK
    launch_tiled_t<SOLIDS, 4, false, true>(h, c, s, o, 0, -1, pp);
#else
    static const bool tma = env_int("LBM_PHI_TMA", 1) != 0;
    if (tma) launch_tiled_t<SOLIDS, 4, true, true>(h, c, s, o, 0, -1, pp); else launch_tiled_t<SOLIDS, 4, false, true>(h, c, s, o, 0, -1, pp);
#endif
}
template <bool SOLIDS>
static void launch_density_tiled_peer(lbm_handle* h, const CGFields& c, const FastFields& s, const PeerPtrs pp) {
    launch_density_tiled_t<SOLIDS, 8, false, true>(h, c, s, 0, -1, pp);
}

template <bool SOLIDS>
static void launch_density_tiled(lbm_handle* h, const CGFields& c, const FastFields& s, int z_lo = 0, int z_hi = -1) {
    // TMA bulk copies for the scalar planes exist (LBM_SCALAR_TMA=1) but measured 7.2 ms vs 4.7 ms per 512^3 launch:
    // 40 row copies per plane step issued by four threads sit on the critical path of this short loop; cp.async stays
#ifdef LBM_HOSTCHECK
    static const bool tma = false;
#else
    static const bool tma = env_int("LBM_SCALAR_TMA", 0) != 0;
#endif
    switch (tile_y_density()) {
        case 4: launch_density_tiled_t<SOLIDS, 4, false>(h, c, s, z_lo, z_hi); break;
        case 16: launch_density_tiled_t<SOLIDS, 16, false>(h, c, s, z_lo, z_hi); break;
        default: if (tma) launch_density_tiled_t<SOLIDS, 8, true>(h, c, s, z_lo, z_hi); else launch_density_tiled_t<SOLIDS, 8, false>(h, c, s, z_lo, z_hi);
    }
}

// ------------------------------------------------------------------------------------------------
// D2Q9: the two passes on TX x TY tiles, one thread per node (no marching: the 2-D configurations are a few hundred thousand
// nodes and live in the L2, so a step is bounded by how many dependent round trips to the L2 a CTA makes and by how many CTAs
// fit on an SM, not by bandwidth).  Everything a tile needs is requested up front -- the phi window (halo 2) and the recolouring
// scalars (halo 1) with cp.async into shared memory, the node's nine pulled populations into registers, with solids BOTH
// candidates of every direction (the upstream node's and the node's own opposite one; the pull mask picks afterwards) -- so a
// tile makes ONE round trip, then works out of shared memory and registers.  The collision pass derives the colour gradient
// and the unit normals of the tile + 1 ring in shared memory (the arithmetic of GradientOp, operation for operation: the
// results are bit-equal to the three one-thread-per-node launches this replaces) and never writes G / n to memory.
// Rows wrap by index arithmetic on a single slab (Grid::wrap2) or reach into the ghost rows of a slab decomposition.
// ------------------------------------------------------------------------------------------------
struct OpenRows {
    int n = 0;
    int mat_lo[2], mat_hi[2];     // planes whose streamed populations the row operators read
    int mod_lo[2], mod_hi[2];     // planes they modify
};
static OpenRows open_rows(const CGFields& c) {
    OpenRows r;
    if (c.outlet != LBM_BC_PERIODIC && c.z_out >= 0) {
        const bool conv = c.outlet == LBM_OUTLET_CONVECTIVE;      // rows 2 <- 3, 1 <- 2, 0 <- 1 | row 1 treated, 0 <- 1
        r.mat_lo[r.n] = 0; r.mat_hi[r.n] = conv ? 4 : 2; r.mod_lo[r.n] = 0; r.mod_hi[r.n] = conv ? 3 : 2; ++r.n;
    }
    if (c.inlet != LBM_BC_PERIODIC && c.z_in >= 0) {
        r.mat_lo[r.n] = r.mod_lo[r.n] = c.z_in; r.mat_hi[r.n] = r.mod_hi[r.n] = c.z_in_ghost + 1; ++r.n;
    }
    return r;
}
constexpr int T2X = 32, T2Y = 8;
template <int TX, int TY>
struct Tile2D {
    const Grid& g; int x0, z0;
    __device__ __forceinline__ int wx(int v) const { return v < 0 ? v + g.n0 : (v >= g.n0 ? v - g.n0 : v); }
    __device__ __forceinline__ int wz(int v) const { return g.wrap2 ? (v < 0 ? v + g.n2 : (v >= g.n2 ? v - g.n2 : v)) : v; }
    // flat id of tile-relative node (lx, lz), any halo depth
    __device__ __forceinline__ int64_t id(int lx, int lz) const { return (int64_t)(wz(z0 + lz) + NG) * g.plane + wx(x0 + lx); }
};

template <bool SOLIDS, int TX, int TY>
__global__ void __launch_bounds__(TX* TY)
cg_density_tile_d2q9(const CGFields c, const FastFields s, const OpenRows skip) {
    using L = D2Q9;
    LBM_PDL_PROLOGUE();
    constexpr int NT = TX * TY, NW = TX + 2, NH = TY + 2;
    LBM_DYN_SMEM(smem_dyn);
    double (*ss)[NH][NW] = reinterpret_cast<double (*)[NH][NW]>(smem_dyn);     // [3]: kR, ax, ay of the tile + 1 ring
    const Grid& g = c.g;
    const int64_t V = g.vol;
    const int tx = threadIdx.x, tz = threadIdx.y, tid = tz * TX + tx;
    const Tile2D<TX, TY> t{g, (int)blockIdx.x * TX, (int)blockIdx.y * TY};
    for (int e = tid; e < NH * NW; e += NT) {
        const int lz = e / NW, lx = e - lz * NW;
        const int64_t src = t.id(lx - 1, lz - 1);
        __pipeline_memcpy_async(&ss[0][lz][lx], s.kR + src, 8);
        __pipeline_memcpy_async(&ss[1][lz][lx], s.a + src, 8);
        __pipeline_memcpy_async(&ss[2][lz][lx], s.a + V + src, 8);
    }
    __pipeline_commit();
    const int64_t id = t.id(tx, tz);
    const uint32_t pm = SOLIDS ? c.pull[id] : 0xFFFFFFFFu;
    double up[L::Q], own[L::Q];
    up[0] = s.gT[id];
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        up[q] = s.gT[q * V + t.id(tx - L::d0(q), tz - L::d2(q))];
        if (SOLIDS) own[q] = s.gT[L::opp(q) * V + id];
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    if (!(pm & 1u)) return;
    {   // forked open-row chain (fast_one_step): the planes it materialises get their densities and phi from that chain
        const int zrow = (int)blockIdx.y * TY + tz;
        if ((skip.n > 0 && zrow >= skip.mat_lo[0] && zrow < skip.mat_hi[0]) || (skip.n > 1 && zrow >= skip.mat_lo[1] && zrow < skip.mat_hi[1])) return;
    }
    const bool pert = c.p.st_type == LBM_ST_PERTURBATION;
    const double kR0 = ss[0][tz + 1][tx + 1];
    const double a0[3] = {ss[1][tz + 1][tx + 1], ss[2][tz + 1][tx + 1], 0.0};
    double accR = kR0 * up[0], accB = up[0] - accR;
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        double gt, fr;
        if (!SOLIDS || (pm & (1u << q))) {
            const int lz = tz + 1 - L::d2(q), lx = tx + 1 - L::d0(q);
            gt = up[q];
            const double an[3] = {ss[1][lz][lx], ss[2][lz][lx], 0.0};
            fr = cg_red_part<L>(q, gt, ss[0][lz][lx], an, pert);
        } else {
            gt = own[q];
            fr = cg_red_part<L>(L::opp(q), gt, kR0, a0, pert);
        }
        accR += fr; accB += gt - fr;
    }
    c.rho[0][id] = accR; c.rho[1][id] = accB;
    c.phi[id] = (accR - accB) / (accR + accB);
}

template <bool SOLIDS, int TX, int TY>
__global__ void __launch_bounds__(TX* TY)
cg_collide_tile_d2q9(const CGFields c, const FastFields s, const FastFields o, const OpenRows rows) {
    using L = D2Q9;
    LBM_PDL_PROLOGUE();
    constexpr int NT = TX * TY, PW = TX + 4, PH = TY + 4, NW = TX + 2, NH = TY + 2;
    constexpr int NE = (NH * NW + NT - 1) / NT;
    LBM_DYN_SMEM(smem_dyn);
    double (*sphi)[PW] = reinterpret_cast<double (*)[PW]>(smem_dyn);                                  // phi, halo 2
    double (*sg)[NH][NW] = reinterpret_cast<double (*)[NH][NW]>(smem_dyn + PH * PW);                  // [4]: Gx, Gy, nx, ny, halo 1
    const Grid& g = c.g;
    const int64_t V = g.vol;
    const int tx = threadIdx.x, tz = threadIdx.y, tid = tz * TX + tx;
    const Tile2D<TX, TY> t{g, (int)blockIdx.x * TX, (int)blockIdx.y * TY};
    for (int e = tid; e < PH * PW; e += NT) {
        const int lz = e / PW, lx = e - lz * PW;
        __pipeline_memcpy_async(&sphi[lz][lx], c.phi + t.id(lx - 2, lz - 2), 8);
    }
    __pipeline_commit();
    // ---- everything else this thread will need, requested before anything is waited for ----
    const int64_t id = t.id(tx, tz);
    const uint32_t pm = SOLIDS ? c.pull[id] : 0xFFFFFFFFu;
    double fT[L::Q], own[L::Q];
    // open-boundary rows: their streamed populations were materialised and treated by the row operators (FastOpenPreOp), the
    // node collides those instead of pulling -- what used to be two launches behind this one (gradient + collision of the rows)
    const int zrow = (int)blockIdx.y * TY + tz;
    const bool treated = (rows.n > 0 && zrow >= rows.mod_lo[0] && zrow < rows.mod_hi[0]) ||
                         (rows.n > 1 && zrow >= rows.mod_lo[1] && zrow < rows.mod_hi[1]);
    if (treated) {
#pragma unroll
        for (int q = 0; q < L::Q; ++q) { fT[q] = c.fS[0][q * V + id] + c.fS[1][q * V + id]; own[q] = 0.0; }
    } else {
        fT[0] = s.gT[id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            fT[q] = s.gT[q * V + t.id(tx - L::d0(q), tz - L::d2(q))];
            if (SOLIDS) own[q] = s.gT[L::opp(q) * V + id];
        }
    }
    const double rR = c.rho[0][id], rB = c.rho[1][id];
    const double Fl[2] = {c.F[id], c.F[V + id]};
    uint8_t cl[NE];
    double nsx[NE], nsy[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = tid + k * NT, lz = e / NW, lx = e - lz * NW;
        cl[k] = CLS_FLUID; nsx[k] = 0.0; nsy[k] = 0.0;
        if (SOLIDS && e < NH * NW) {
            const int64_t nid = t.id(lx - 1, lz - 1);
            cl[k] = c.cls[nid]; nsx[k] = c.ns[nid]; nsy[k] = c.ns[V + nid];
        }
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    // ---- colour gradient and unit normal of the tile + 1 ring: GradientOp's arithmetic from the phi window ----
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = tid + k * NT;
        if (e >= NH * NW) break;
        const int lz = e / NW, lx = e - lz * NW;
        double G[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0};
        if (cl[k] & CLS_FLUID) {
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double v = mul_rn(L::w(q), sphi[lz + 1 + L::d2(q)][lx + 1 + L::d0(q)]);
#pragma unroll
                for (int a = 0; a < L::D; ++a)
                    if (L::c(q, a) != 0) G[a] = add_rn(G[a], L::c(q, a) > 0 ? v : -v);
            }
            G[0] *= 3.0; G[1] *= 3.0;
            if (SOLIDS && (cl[k] & CLS_NEAR)) {
                double ns[3] = {nsx[k], nsy[k], 0.0};
                cg_wetting<2>(G, ns, c.p.cosT, c.p.sinT, c.p.wetting_type);
            }
            cg_unit_normal<2>(G, c.p.wetting_type, n);
        }
        sg[0][lz][lx] = G[0]; sg[1][lz][lx] = G[1]; sg[2][lz][lx] = n[0]; sg[3][lz][lx] = n[1];
    }
    __syncthreads();
    if (!(pm & 1u)) return;
    if (SOLIDS && !treated) {
#pragma unroll
        for (int q = 1; q < L::Q; ++q)
            if (!(pm & (1u << q))) fT[q] = own[q];          // half-way bounce back
    }
    // ---- PullCollideOp from here on, with the normals in shared memory ----
    const double rho = rB + rR;
    double mom[3] = {0.0, 0.0, 0.0}, u[3] = {0, 0, 0}, G[3] = {0, 0, 0}, n[3] = {0, 0, 0}, F[3] = {0, 0, 0};
#pragma unroll
    for (int q = 1; q < L::Q; ++q)
#pragma unroll
        for (int d = 0; d < L::D; ++d)
            if (L::c(q, d) != 0) mom[d] += L::c(q, d) * fT[q];
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
        u[d] = (mom[d] + 0.5 * Fl[d]) / rho;                // force of the previous step
        G[d] = sg[d][tz + 1][tx + 1]; n[d] = sg[2 + d][tz + 1][tx + 1];
    }
    double dn[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};   // cg_force_at
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        const int lz = tz + 1 + L::d2(q), lx = tx + 1 + L::d0(q);
        const double nk[2] = {sg[2][lz][lx], sg[3][lz][lx]};
#pragma unroll
        for (int a = 0; a < L::D; ++a)
            if (L::c(q, a) != 0)
#pragma unroll
                for (int b = 0; b < L::D; ++b) dn[a][b] = add_rn(dn[a][b], mul_rn(3.0 * L::w(q) * L::c(q, a), nk[b]));
    }
    double K = 0.0, nn = 0.0, div = 0.0;
#pragma unroll
    for (int a = 0; a < L::D; ++a) {
        nn += n[a] * n[a]; div += dn[a][a];
#pragma unroll
        for (int b = 0; b < L::D; ++b) K += n[a] * n[b] * dn[a][b];
    }
    K -= nn * div;
    const double sgn = c.p.wetting_type == 1 ? 0.5 : -0.5;
#pragma unroll
    for (int a = 0; a < L::D; ++a) { F[a] = sgn * c.p.sigma * K * G[a]; c.F[a * V + id] = F[a]; }
    const double tau = cg_tau(sphi[tz + 2][tx + 2], rR, rB, c.p);
    cg_collide<L>(fT, rho, u, F, tau, c.p.relax);
    double kR, a[3];
    cg_recolour_coeffs<L>(rR, rB, G, c.p.beta, &kR, a);
#pragma unroll
    for (int q = 0; q < L::Q; ++q) o.gT[q * V + id] = fT[q];
    o.kR[id] = kR;
#pragma unroll
    for (int d = 0; d < 3; ++d) o.a[d * V + id] = a[d];
}

static bool tiled2d_ok(const lbm_handle* h) {
    static const bool off = env_int("LBM_TILE_2D", 1) == 0;
    return h->Q == 9 && !off && h->g.n0 % T2X == 0 && h->g.n2 % T2Y == 0 && !(h->cfg.flags & 2u) && !h->tracer && (!h->has_solid || h->pull);
}
template <bool SOLIDS>
static void launch_density_tile2d(lbm_handle* h, const CGFields& c, const FastFields& s, const OpenRows& skip = OpenRows()) {
    const Grid& g = h->g;
    dim3 grid(g.n0 / T2X, g.n2 / T2Y), block(T2X, T2Y);
    constexpr size_t smem = sizeof(double) * 3 * (T2Y + 2) * (T2X + 2);
#ifdef LBM_HOSTCHECK
    cta_emu::launch(grid, block, smem, [&] { cg_density_tile_d2q9<SOLIDS, T2X, T2Y>(c, s, skip); });
#else
    if (g_prof.on) g_prof.begin(SOLIDS ? "cg_density_tile_d2q9<solids>" : "cg_density_tile_d2q9<all-fluid>", h->stream);
    launch_kernel(cg_density_tile_d2q9<SOLIDS, T2X, T2Y>, grid, block, smem, h->stream, c, s, skip);
    if (g_prof.on) g_prof.end(h->stream);
    LBM_CUDA_CHECK(cudaGetLastError());
#endif
    ++g_launch_counter;
}
template <bool SOLIDS>
static void launch_collide_tile2d(lbm_handle* h, const CGFields& c, const FastFields& s, const FastFields& o, const OpenRows& rows) {
    const Grid& g = h->g;
    dim3 grid(g.n0 / T2X, g.n2 / T2Y), block(T2X, T2Y);
    constexpr size_t smem = sizeof(double) * ((T2Y + 4) * (T2X + 4) + 4 * (T2Y + 2) * (T2X + 2));
#ifdef LBM_HOSTCHECK
    cta_emu::launch(grid, block, smem, [&] { cg_collide_tile_d2q9<SOLIDS, T2X, T2Y>(c, s, o, rows); });
#else
    if (g_prof.on) g_prof.begin(SOLIDS ? "cg_collide_tile_d2q9<solids>" : "cg_collide_tile_d2q9<all-fluid>", h->stream);
    launch_kernel(cg_collide_tile_d2q9<SOLIDS, T2X, T2Y>, grid, block, smem, h->stream, c, s, o, rows);
    if (g_prof.on) g_prof.end(h->stream);
    LBM_CUDA_CHECK(cudaGetLastError());
#endif
    ++g_launch_counter;
}

// which ghost planes of the factored state are ever read: population q is pulled from z - c_z(q), so it has
// to travel upwards (c_z = +1) or downwards (c_z = -1) only, in-plane directions never cross a slab face
template <class L>
static const int8_t* factored_dirs() {
    struct Table {
        int8_t d[L::Q + 4];
        Table() {
            for (int q = 0; q < L::Q; ++q) d[q] = L::d2(q) == 0 ? 2 : (int8_t)L::d2(q);
            for (int k = 0; k < 4; ++k) d[L::Q + k] = 0;
        }
    };
    static const Table t;        // built once (several handles may step from several host threads)
    return t.d;
}

// ---- open boundaries on the fast path -------------------------------------------------------------------------
// The inlet / outlet treatment of the reference (RKD2Q9.py:1299-1352) rewrites the STREAMED populations of a few
// planes at the top of every iteration, which the factored state cannot express.  Those planes are therefore patched:
// after the density pass their streamed populations are materialised from the factored state, the reference's row
// operators run on them, and velocity + phi are re-evaluated there; after the collision pass the same planes are
// collided again from the treated populations (general-path arithmetic) and overwrite the fast pass's result.
// Everything else of the lattice stays on the two fused passes.
// planes [lo[k], hi[k]) (k < n <= 2) of an operator whose item 0 is the first node of plane -ext, in one launch
template <class Op>
static void launch_plane_ranges(lbm_handle* h, const Op& op, int ext, int n, const int* lo, const int* hi) {
    const int64_t plane = h->g.plane;
    if (n == 1) {
        launch(PlaneRangeOp<Op>{op, (int64_t)(lo[0] + ext) * plane}, (int64_t)(hi[0] - lo[0]) * plane, h->stream);
    } else if (n == 2) {
        const int64_t cnt0 = (int64_t)(hi[0] - lo[0]) * plane, cnt1 = (int64_t)(hi[1] - lo[1]) * plane;
        launch(TwoRangeOp<Op>{op, (int64_t)(lo[0] + ext) * plane, cnt0, (int64_t)(lo[1] + ext) * plane}, cnt0 + cnt1, h->stream);
    }
}
// One thread per column and side: materialise the streamed populations of the column's open rows from the factored state
// (read-only neighbours), run the row operators on them, re-evaluate velocity (lagged force) and phi there.  Every step of
// that chain only reads what the same thread wrote, so the three launches it used to be are one.
template <class L>
struct FastOpenPreOp {
    CGFields c; FastFields s; OpenRows rows;
    LBM_HD void operator()(int64_t i) const {
        const int64_t plane = c.g.plane;
        const bool outlet_side = i < plane;
        const int64_t r = outlet_side ? i : i - plane;
        const bool has_out = c.outlet != LBM_BC_PERIODIC && c.z_out >= 0, has_in = c.inlet != LBM_BC_PERIODIC && c.z_in >= 0;
        if (outlet_side ? !has_out : !has_in) return;
        const int k = outlet_side ? 0 : (has_out ? 1 : 0);
        for (int z = rows.mat_lo[k]; z < rows.mat_hi[k]; ++z) PullMaterialiseOp<L>{c, s}((int64_t)z * plane + r);
        OpenRowsOp<L>{c}(i);
        for (int z = rows.mod_lo[k]; z < rows.mod_hi[k]; ++z) HeadOp<L>{c}((int64_t)z * plane + r);
    }
};
template <class L>
static void fast_open_rows_pre(lbm_handle* h, const CGFields& c, const FastFields& s, bool head_on_materialised = false) {
    const OpenRows r = open_rows(c);
    if (!r.n) return;
    if (head_on_materialised) {     // forked chain: nobody else writes phi of the materialised planes
        launch_plane_ranges(h, PullMaterialiseOp<L>{c, s}, 0, r.n, r.mat_lo, r.mat_hi);
        launch(OpenRowsOp<L>{c}, 2 * h->g.plane, h->stream);
        launch_plane_ranges(h, HeadOp<L>{c}, 0, r.n, r.mat_lo, r.mat_hi);
        return;
    }
    // Three fully parallel launches (materialise the planes | row operators | velocity and phi of the treated planes) beat ONE launch
    // whose threads walk the rows of their column: the chain of one column is a few thousand dependent instructions and ~8 round
    // trips to the L2, and that latency -- not the launch count -- is what a replayed graph pays for.  BASELINE config 2: 64 -> 47 us
    // per step (4 101 -> 5 541 MLUPS, profiles/r02b_cfg2_split{0,1}.json); the end slabs of config 5 on 8 GPUs: 0.31 ms.
    // LBM_OPEN_PRE_SPLIT=0 selects the single launch (FastOpenPreOp).
    static const int split = env_int("LBM_OPEN_PRE_SPLIT", 1);
    if (split != 0) {
        launch_plane_ranges(h, PullMaterialiseOp<L>{c, s}, 0, r.n, r.mat_lo, r.mat_hi);
        launch(OpenRowsOp<L>{c}, 2 * h->g.plane, h->stream);
        launch_plane_ranges(h, HeadOp<L>{c}, 0, r.n, r.mod_lo, r.mod_hi);
        return;
    }
    launch(FastOpenPreOp<L>{c, s, r}, 2 * h->g.plane, h->stream);
}
template <class L>
static void fast_open_rows_post(lbm_handle* h, const CGFields& c, const FastFields& o, bool need_gradient) {
    const OpenRows r = open_rows(c);
    if (!r.n) return;
    if (c.p.st_type == LBM_ST_PERTURBATION) {     // this model's collision evaluates its gradient itself
        launch_plane_ranges(h, PerturbCollideFactoredOp<L>{c, o}, 0, r.n, r.mod_lo, r.mod_hi);
        return;
    }
    if (need_gradient) {      // the tiled collision pass keeps G and the normals in shared memory: evaluate them around the patched planes
        int lo[2], hi[2];
        for (int k = 0; k < r.n; ++k) {
            lo[k] = r.mod_lo[k] - 1 < -1 ? -1 : r.mod_lo[k] - 1;
            hi[k] = r.mod_hi[k] + 1 > h->g.n2 + 1 ? h->g.n2 + 1 : r.mod_hi[k] + 1;
        }
        if (r.n == 2 && hi[0] > lo[1]) { hi[0] = hi[1]; launch_plane_ranges(h, GradientOp<L>{c}, 1, 1, lo, hi); }   // thin slab: the two ranges touch
        else launch_plane_ranges(h, GradientOp<L>{c}, 1, r.n, lo, hi);
    }
    launch_plane_ranges(h, CollideFactoredOp<L>{c, o}, 0, r.n, r.mod_lo, r.mod_hi);
}

template <class L>
static void fast_enter(lbm_handle* h) {
    cg_ensure_head(h);
    if (h->cfg.surface_tension_type == LBM_ST_PERTURBATION) {
        CGFields c = h->fields();
        FastState* f = (FastState*)h->fast;
        exchange_f64(h, c.phi, 0, 1, 1);
        f->pushed[f->cur] = 0;
        launch(PerturbCollideFactoredOp<L>{c, fast_fields(h, f->cur)}, h->g.count(0), h->stream);
        h->head_done = false;
        h->fast_pending_stream = true;
        return;
    }
    cg_generic_forces(h);
    CGFields c = h->fields();
    FastState* f = (FastState*)h->fast;
    tracer_phase(h);          // no-op without tracers (or when a download already ran it for this iteration)
    f->pushed[f->cur] = 0;
    launch(CollideFactoredOp<L>{c, fast_fields(h, f->cur)}, h->g.count(0), h->stream);
    tracer_iteration_finished(h);
    h->head_done = false;
    h->fast_pending_stream = true;
}

#ifndef LBM_HOSTCHECK
// Slab decomposition, all-fluid slab, tiled kernels (opt-in, LBM_FLAG_OVERLAP): the two ghost-plane exchanges of
// a step run on the communication stream while the main stream works on the planes that do not need them.
// Measured at 4 GPUs it is 4 % SLOWER than the serial schedule: the exchanges cost ~0.3 ms of a 3.9 ms step,
// about as much as the extra prologues of the six thin boundary launches, so it is off by default.
//   comm:  [factored state, 1 plane]            [phi, 2 planes]
//   main:  density pass planes 1..n-2 | 0, n-1   collision pass planes 2..n-3 | 0,1, n-2,n-1
static void fast_one_step_overlapped(lbm_handle* h) {
    using L = D3Q19;
    FastState* f = (FastState*)h->fast;
    const Grid& g = h->g;
    CGFields c = h->fields();
    const FastFields s = fast_fields(h, f->cur), o = fast_fields(h, 1 - f->cur);
    const int n = g.n2;
    LBM_CUDA_CHECK(cudaEventRecord(h->ev_main, h->stream));             // previous collision pass complete
    LBM_CUDA_CHECK(cudaStreamWaitEvent(h->comm_stream, h->ev_main, 0));
    h->xstream = h->comm_stream;
    exchange_f64(h, f->buf[f->cur], g.vol, L::Q + 4, 1, factored_dirs<L>());
    LBM_CUDA_CHECK(cudaEventRecord(h->ev_comm, h->comm_stream));
    launch_density_tiled<false>(h, c, s, 1, n - 1);                     // needs no ghost plane
    LBM_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
    launch_density_tiled<false>(h, c, s, 0, 1);
    launch_density_tiled<false>(h, c, s, n - 1, n);
    LBM_CUDA_CHECK(cudaEventRecord(h->ev_main, h->stream));             // phi complete on the owned planes
    LBM_CUDA_CHECK(cudaStreamWaitEvent(h->comm_stream, h->ev_main, 0));
    exchange_f64(h, c.phi, 0, 1, 2);
    LBM_CUDA_CHECK(cudaEventRecord(h->ev_comm, h->comm_stream));
    h->xstream = nullptr;
    launch_tiled<false>(h, c, s, o, 2, n - 2);                          // phi halo of 2 planes stays inside the slab
    LBM_CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
    launch_tiled<false>(h, c, s, o, 0, 2);
    launch_tiled<false>(h, c, s, o, n - 2, n);
    f->cur = 1 - f->cur;
}
#endif

// one direction of a one-sided exchange as a launch of its own: the planes an open-row patch rewrites after the pass that
// stores the others (the outlet planes of the first slab travel down, the inlet planes of the last slab up)
static void peer_push_one_way(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs, bool upwards) {
    PeerPushOp op;
    comm_peer_pointers(h, base, &op.up, &op.down);
    op.g = h->g; op.base = base; op.stride = stride; op.narr = narr; op.gp = gp;
    for (int a = 0; a < 48; ++a) {
        const int d = (a < narr && dirs) ? dirs[a] : 0;
        op.dirs[a] = a >= narr ? 2 : (upwards ? ((d == 0 || d == 1) ? 1 : 2) : ((d == 0 || d == -1) ? -1 : 2));
    }
    launch(op, op.items(), h->stream);
}

// the two per-step exchanges of the fast path: NCCL send / recv, or (opt-in) stores into the neighbours' memory + flags
static void fast_exchange(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs = nullptr) {
    if (h->nranks > 1 && h->peer_ok) comm_peer_exchange_f64(h, base, stride, narr, gp, dirs);
    else exchange_f64(h, base, stride, narr, gp, dirs);
}

template <class L>
static void fast_one_step(lbm_handle* h) {
    FastState* f = (FastState*)h->fast;
    const Grid& g = h->g;
#ifndef LBM_HOSTCHECK
    if (h->nranks > 1 && !h->has_solid && !open_box(h) && tiled_ok(h) && g.n2 >= 8 && (h->cfg.flags & 16u) && !(h->cfg.flags & 4u) &&
        h->cfg.surface_tension_type == LBM_ST_CSF) {
        fast_one_step_overlapped(h);
        return;
    }
#endif
    CGFields c = h->fields();
    const FastFields s = fast_fields(h, f->cur), o = fast_fields(h, 1 - f->cur);
    const bool open = open_box(h);
    // one-sided exchange with the stores fused into the tiled passes.  Open channels: the open-row patches rewrite the outlet
    // planes of the first slab and the inlet planes of the last slab AFTER the passes, so those two slabs leave that direction
    // to a one-way push behind the patch.
    const bool peer = h->nranks > 1 && h->peer_ok;
    const bool pert = h->cfg.surface_tension_type == LBM_ST_PERTURBATION;
    const bool pert_tiled = pert && tiled_ok(h) && h->g.n1 % 4 == 0 && (!h->has_solid || h->pull);
    const bool fused = peer && tiled_ok(h) && !(h->cfg.flags & 4u) && peer_tiles_default() && (!pert || pert_tiled);
    const bool late_down = fused && open && h->rank == 0, late_up = fused && open && h->rank == h->nranks - 1;
    if (peer && f->pushed[f->cur] == 1) comm_peer_signal_wait(h);
    else if (!(peer && f->pushed[f->cur] == 2)) fast_exchange(h, f->buf[f->cur], g.vol, L::Q + 4, 1, factored_dirs<L>());
    f->pushed[f->cur] = 0;
    bool dens_done = false, phi_pushed = false;
    if (tiled_ok(h) && !(h->cfg.flags & 4u)) {       // host test hook: the same kernels on host threads (cta_emu.h)
        if (fused) {
            PeerPtrs pp{nullptr, nullptr, h->has_solid ? NG : 2};
            comm_peer_pointers(h, c.phi, &pp.up, &pp.down);
            if (late_up) pp.up = nullptr;
            if (late_down) pp.down = nullptr;
            if (h->has_solid) launch_density_tiled_peer<true>(h, c, s, pp); else launch_density_tiled_peer<false>(h, c, s, pp);
            phi_pushed = true;
        } else if (h->has_solid) launch_density_tiled<true>(h, c, s); else launch_density_tiled<false>(h, c, s);
        dens_done = true;
    }
    const bool tile2d = L::Q == 9 && tiled2d_ok(h) && !pert;
    // D2Q9 tiles on one slab, open box: the open-row chain only reads the factored state, so it runs BESIDE the
    // density tile (a second stream / a parallel branch of the replayed graph) once the tile leaves the materialised planes alone
    // BASELINE config 2, same box, A B A B: 5 262 / 5 335 -> 5 844 / 5 937 MLUPS (profiles/r02b_cfg2_fork{0,1}.json); LBM_OPEN_FORK=0: serial order
    static const bool fork_wanted = env_int("LBM_OPEN_FORK", 1) != 0;
    const bool fork = fork_wanted && tile2d && open && !dens_done && h->nranks == 1;
    if (!dens_done) {
        if (tile2d) {
            const OpenRows skip = fork ? open_rows(c) : OpenRows();
            if (fork) side_stream_fork(h);
            if (h->has_solid) launch_density_tile2d<true>(h, c, s, skip); else launch_density_tile2d<false>(h, c, s, skip);
        }
        else if (h->has_solid) launch(PullDensityOp<L, true>{c, s}, g.count(0), h->stream);
        else launch(PullDensityOp<L, false>{c, s}, g.count(0), h->stream);
    }
    if (fork) {
        side_stream_swap(h);
        try { fast_open_rows_pre<L>(h, c, s, true); } catch (...) { side_stream_swap(h); throw; }
        side_stream_swap(h);
        side_stream_join(h);
    } else if (open) fast_open_rows_pre<L>(h, c, s, tile2d);     // D2Q9 tiles: phi of the materialised planes from the head operator in
                                                                 // EVERY schedule (forked or not, one slab or many), so that slabs stay
                                                                 // bit-equal to one GPU under nvcc's per-kernel multiply-add contraction
    if (phi_pushed) {
        if (late_up) peer_push_one_way(h, c.phi, 0, 1, h->has_solid ? NG : 2, nullptr, true);
        if (late_down) peer_push_one_way(h, c.phi, 0, 1, h->has_solid ? NG : 2, nullptr, false);
        comm_peer_signal_wait(h);
    } else fast_exchange(h, c.phi, 0, 1, h->has_solid ? NG : 2);
    if (h->has_solid && (tiled_ok(h) || tile2d) && !pert) {      // the tiled kernels stage phi, wetting solids included
        if (h->pull) launch(PhiSolidListOp<L>{c, h->wet_list}, h->n_wet_list, h->stream);
        else launch(PhiSolidOp<L>{c}, g.count(2), h->stream);
    }
    bool done = false;
    if (pert) {
        if (pert_tiled && fused) {
            PeerPtrs pp{nullptr, nullptr, 1};
            comm_peer_pointers(h, f->buf[1 - f->cur], &pp.up, &pp.down);
            if (late_up) pp.up = nullptr;
            if (late_down) pp.down = nullptr;
            if (h->has_solid) launch_perturb_tiled<true, true>(h, c, s, o, pp); else launch_perturb_tiled<false, true>(h, c, s, o, pp);
            f->pushed[1 - f->cur] = 1;
        } else if (pert_tiled) {
            if (h->has_solid) launch_perturb_tiled<true>(h, c, s, o); else launch_perturb_tiled<false>(h, c, s, o);
        } else if (h->has_solid) launch(PullPerturbCollideOp<L, true>{c, s, o}, g.count(0), h->stream);
        else launch(PullPerturbCollideOp<L, false>{c, s, o}, g.count(0), h->stream);
        done = true;
    }
    if (!done && tiled_ok(h)) {
        if (fused) {
            PeerPtrs pp{nullptr, nullptr, 1};
            comm_peer_pointers(h, f->buf[1 - f->cur], &pp.up, &pp.down);
            if (late_up) pp.up = nullptr;
            if (late_down) pp.down = nullptr;
            if (h->has_solid) launch_tiled_peer<true>(h, c, s, o, pp); else launch_tiled_peer<false>(h, c, s, o, pp);
            f->pushed[1 - f->cur] = 1;
        } else if (h->has_solid) launch_tiled<true>(h, c, s, o); else launch_tiled<false>(h, c, s, o);
        done = true;
    }
    bool rows_done = false;
    if (!done && tile2d) {
        // the tile kernel collides the treated open rows itself (LBM_TILE_2D_ROWS=0: the two patch launches behind it, as before)
        static const bool fold_rows = env_int("LBM_TILE_2D_ROWS", 1) != 0;
        const OpenRows rows = (open && fold_rows) ? open_rows(c) : OpenRows();
        if (h->has_solid) launch_collide_tile2d<true>(h, c, s, o, rows); else launch_collide_tile2d<false>(h, c, s, o, rows);
        done = true;
        rows_done = open && fold_rows;
    }
    if (!done) {
        launch(GradientOp<L>{c}, g.count(1), h->stream);
        if (h->has_solid) launch(PullCollideOp<L, true>{c, s, o}, g.count(0), h->stream);
        else launch(PullCollideOp<L, false>{c, s, o}, g.count(0), h->stream);
        if (h->tracer) { tracer_phase(h); tracer_iteration_finished(h); }
    }
    if (open && !rows_done) fast_open_rows_post<L>(h, c, o, done);
    if (late_up) peer_push_one_way(h, f->buf[1 - f->cur], g.vol, L::Q + 4, 1, factored_dirs<L>(), true);
    if (late_down) peer_push_one_way(h, f->buf[1 - f->cur], g.vol, L::Q + 4, 1, factored_dirs<L>(), false);
    f->cur = 1 - f->cur;
}

// ------------------------------------------------------------------------------------------------
// Persistent form of the one-thread-per-node fast path (LBM_FLAG_PERSISTENT, opt-in).  The 2-D configurations are a few
// hundred thousand nodes: their lattice lives in the L2 and a step is bounded by its launches, not by bandwidth.  Here every
// step of an lbm_step call runs inside ONE cooperative kernel -- as many CTAs as are co-resident, grid-stride loops over the
// nodes, and a grid-wide barrier where the launch boundaries used to be (the phases of a step read what other threads wrote in
// the phase before).  Same operators, same order: bit-equal to the launched form (tests/test_hostcheck_tiled.py runs it on
// host threads).  Needs one slab with index wrap (no ghost-plane copies between the phases).
// ------------------------------------------------------------------------------------------------
template <class L, bool SOLIDS>
__global__ void __launch_bounds__(256)
cg_fast_persistent(const CGFields c, const FastFields b0, const FastFields b1, const OpenRows rows, const int open,
                   const int nsteps, const int cur0) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const Grid& g = c.g;
    const int64_t n_owned = g.count(0), n_grad = g.count(1), n_rows = 2 * g.plane;
    for (int step = 0; step < nsteps; ++step) {
        const bool odd = (cur0 + step) & 1;
        const FastFields& s = odd ? b1 : b0;
        const FastFields& o = odd ? b0 : b1;
        for (int64_t i = tid; i < n_owned; i += nth) PullDensityOp<L, SOLIDS>{c, s}(i);
        LBM_GRID_SYNC();
        if (open) {
            for (int64_t i = tid; i < n_rows; i += nth) FastOpenPreOp<L>{c, s, rows}(i);
            LBM_GRID_SYNC();
        }
        for (int64_t i = tid; i < n_grad; i += nth) GradientOp<L>{c}(i);
        LBM_GRID_SYNC();
        for (int64_t i = tid; i < n_owned; i += nth) PullCollideOp<L, SOLIDS>{c, s, o}(i);
        if (open) {
            LBM_GRID_SYNC();
            for (int k = 0; k < rows.n; ++k) {
                const int64_t off = (int64_t)rows.mod_lo[k] * g.plane, cnt = (int64_t)(rows.mod_hi[k] - rows.mod_lo[k]) * g.plane;
                for (int64_t i = tid; i < cnt; i += nth) CollideFactoredOp<L>{c, o}(i + off);
            }
        }
        LBM_GRID_SYNC();
    }
}

static bool persistent_ok(const lbm_handle* h) {
    return (h->cfg.flags & LBM_FLAG_PERSISTENT) && h->nranks == 1 && h->g.wrap2 && !tiled_ok(h) && !g_prof_active() && !h->tracer &&
           h->cfg.surface_tension_type == LBM_ST_CSF;
}

template <class L, bool SOLIDS>
static void launch_persistent_t(lbm_handle* h, int nsteps) {
    FastState* f = (FastState*)h->fast;
    CGFields c = h->fields();
    FastFields b0 = fast_fields(h, 0), b1 = fast_fields(h, 1);
    OpenRows rows = open_rows(c);
    int open = open_box(h) ? 1 : 0, cur0 = f->cur;
#ifdef LBM_HOSTCHECK
    cta_emu::launch_cooperative(dim3(3), dim3(32), [&] { cg_fast_persistent<L, SOLIDS>(c, b0, b1, rows, open, nsteps, cur0); });
#else
    static std::atomic<int> grid_for_device[64];
    int grid = grid_for_device[h->cfg.device & 63];
    if (!grid) {
        int per_sm = 0, sms = 0;
        LBM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_fast_persistent<L, SOLIDS>, 256, 0));
        LBM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if (per_sm < 1) throw BackendError{"the persistent kernel does not fit on an SM"};
        grid = per_sm * sms;
        grid_for_device[h->cfg.device & 63] = grid;
    }
    void* args[] = {&c, &b0, &b1, &rows, &open, &nsteps, &cur0};
    LBM_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)cg_fast_persistent<L, SOLIDS>, dim3(grid), dim3(256), args, 0, h->stream));
#endif
    ++g_launch_counter;
    f->cur = (f->cur + nsteps) & 1;
}
static void launch_persistent(lbm_handle* h, int nsteps) {
    if (h->Q == 9) { if (h->has_solid) launch_persistent_t<D2Q9, true>(h, nsteps); else launch_persistent_t<D2Q9, false>(h, nsteps); }
    else { if (h->has_solid) launch_persistent_t<D3Q19, true>(h, nsteps); else launch_persistent_t<D3Q19, false>(h, nsteps); }
}

void cg_fast_step(lbm_handle* h, int nsteps) {
    if (nsteps <= 0) return;
    fast_alloc(h);
    int left = nsteps;
    if (!h->fast_pending_stream) {
        if (h->Q == 9) fast_enter<D2Q9>(h); else fast_enter<D3Q19>(h);
        --left;
    }
    if (left > 0 && persistent_ok(h)) { launch_persistent(h, left); return; }
    auto one = [&] { if (h->Q == 9) fast_one_step<D2Q9>(h); else fast_one_step<D3Q19>(h); };
    if (left > 0) { one(); --left; }          // outside the graph: configures the tiled kernels on first use
    // the double buffer flips every step, so the replayed unit is a pair of steps
    replay(left / 2, h->graph_ok(), &h->graph, h->stream, [&] { one(); one(); });
    if (left & 1) one();
    // One-sided exchange with the stores fused into the passes: the handshake of the last collision pass is not left to a
    // next step that may never come.  When this call's work has drained, every store of the neighbours into my ghost planes
    // has landed (their signal is ordered behind their stores), so nothing is in flight between lbm_step calls and the
    // buffers may be freed / re-initialised / downloaded without talking to anybody.
    FastState* f = (FastState*)h->fast;
    if (h->nranks > 1 && h->peer_ok && f->pushed[f->cur] == 1) { comm_peer_signal_wait(h); f->pushed[f->cur] = 2; }
}

void cg_fast_materialise(lbm_handle* h) {
    if (!h->fast || !h->fast_pending_stream) return;
    FastState* f = (FastState*)h->fast;
    const Grid& g = h->g;
    CGFields c = h->fields();
    exchange_f64(h, f->buf[f->cur], g.vol, h->Q + 4, 1);
    if (h->Q == 9) launch(PullMaterialiseOp<D2Q9>{c, fast_fields(h, f->cur)}, g.count(0), h->stream);
    else launch(PullMaterialiseOp<D3Q19>{c, fast_fields(h, f->cur)}, g.count(0), h->stream);
    h->fast_pending_stream = false;
    h->head_done = false;
}

}  // namespace lbm
