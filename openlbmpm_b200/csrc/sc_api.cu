// sc_api.cu -- state and step loops of the Shan-Chen models behind the C ABI
// (original Shan-Chen: ShanChenD2Q9.py:1492-1629; explicit forcing SRT/MRT: ShanChenD2Q9.py:1714-2087).
// D2Q9 follows the reference; D3Q19 (`ShanChenD3Q19.runOriginalSC3DGPU / runEFS4LBM3DGPU`, named by main.py:73-77
// but absent upstream) runs the same lattice-generic operators, open boundaries included (oracle/sc_dense.py).
#include <atomic>
#include "coop.h"
#include "internal.h"
#include "sc_fast.cuh"

namespace lbm {

// launch a lattice-generic operator for the handle's lattice
#define SC_LAUNCH(n, OP, ...)                                                      \
    do {                                                                           \
        if (h->Q == 9) launch(OP<D2Q9>{__VA_ARGS__}, n, h->stream);                \
        else launch(OP<D3Q19>{__VA_ARGS__}, n, h->stream);                         \
    } while (0)

struct SCState {
    double *fS = nullptr, *fC = nullptr, *rho = nullptr, *F = nullptr, *ueq = nullptr, *uph = nullptr, *fold = nullptr;
    bool efs_prepared = false;    // EFS pre-loop (force, u_eq, f <- f - fF/2, boundary rows) done
    bool head_done = false;       // SC: inlet treatment of the current iteration already applied
    bool rho_is_sum = false;      // SC: rho == sum_q f_q on every node (the streaming leaves it so; a fresh state may not)
    bool uph_valid = true;        // SC: the physical velocity of the last finished iteration has been evaluated
};

// planes [z_lo, z_hi) of a lattice-generic operator whose item 0 is the first node of plane 0
#define SC_LAUNCH_PLANES(z_lo, z_hi, OP, ...)                                                                      \
    do {                                                                                                           \
        const int64_t off_ = (int64_t)(z_lo) * h->g.plane, cnt_ = (int64_t)((z_hi) - (z_lo)) * h->g.plane;         \
        if (h->Q == 9) launch(PlaneRangeOp<OP<D2Q9>>{OP<D2Q9>{__VA_ARGS__}, off_}, cnt_, h->stream);              \
        else launch(PlaneRangeOp<OP<D3Q19>>{OP<D3Q19>{__VA_ARGS__}, off_}, cnt_, h->stream);                     \
    } while (0)

static SCFields sc_fields(const lbm_handle* h) {
    const SCState* s = (const SCState*)h->sc;
    SCFields c;
    memset(&c, 0, sizeof(c));
    c.g = h->g; c.Q = h->Q; c.D = h->D;
    const lbm_config& cfg = h->cfg;
    c.p.nc = cfg.n_components; c.p.relax = cfg.relax; c.p.inlet = cfg.inlet; c.p.outlet = cfg.outlet;
    for (int k = 0; k < SC_MAXC; ++k) {
        c.p.tau[k] = cfg.sc_tau[k]; c.p.Gs[k] = cfg.sc_Gsolid[k]; c.p.vin[k] = cfg.sc_inlet_velocity[k];
        c.p.rho_out[k] = cfg.sc_rho_out[k];
        for (int j = 0; j < SC_MAXC; ++j) c.p.G[k * SC_MAXC + j] = cfg.sc_G[k * SC_MAXC + j];
    }
    c.fS = s->fS; c.fC = s->fC; c.rho = s->rho; c.F = s->F; c.ueq = s->ueq; c.uph = s->uph; c.fold = s->fold;
    c.cls = h->cls;
    c.p.scheme = cfg.sc_isotropy == 8 || cfg.sc_isotropy == 10 ? cfg.sc_isotropy : 4;
    // isotropy 8 moves the treated rows one node inwards and keeps two ghost rows
    // (constantVelocityZouHeBoundaryHigher8, constantPressureZouHeBoundaryLower8: OptimizedD2Q9GPU.py:590-620, 868-890)
    const int deep = c.p.scheme == 8 && cfg.model == LBM_MODEL_EFS ? 1 : 0;
    c.z_in = h->g.n2 - 2 - deep; c.z_in_ghost = h->g.n2 - 1; c.z_out = 1 + deep;
    return c;
}

void sc_free(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    if (!s) return;
    double* arrs[] = {s->fS, s->fC, s->rho, s->F, s->ueq, s->uph, s->fold};
    for (double* p : arrs) dev_free(p);
    delete s;
    h->sc = nullptr;
}

static void sc_alloc(lbm_handle* h) {
    if (h->sc) return;
    SCState* s = new SCState();
    h->sc = s;
    const int nc = h->cfg.n_components;
    const size_t V = (size_t)h->g.vol * sizeof(double);
    auto alloc0 = [&](size_t bytes) { double* p = (double*)dev_alloc(bytes); dev_zero(p, bytes, h->stream); return p; };
    const int Q = h->Q, D = h->D;
    s->fS = alloc0(nc * Q * V); s->fC = alloc0(nc * Q * V); s->rho = alloc0(nc * V); s->F = alloc0(nc * D * V);
    s->ueq = alloc0(D * V); s->uph = alloc0(D * V);
    s->fold = alloc0((size_t)nc * Q * 3 * h->g.plane * sizeof(double));
}

int sc_init_equilibrium(lbm_handle* h, const double* const* rho, int32_t n_comp) {
    if (n_comp != h->cfg.n_components || !rho) { h->err = "one density array per component expected"; return LBM_EINVAL; }
    for (int k = 0; k < n_comp; ++k) if (!rho[k]) { h->err = "NULL density array"; return LBM_EINVAL; }
    sc_alloc(h);
    const int64_t owned = h->g.plane * h->g.n2;
    double* tmp = (double*)dev_alloc((size_t)n_comp * owned * 8);
    try {
        for (int k = 0; k < n_comp; ++k) dev_h2d(tmp + k * owned, rho[k], owned * 8, h->stream);
        SC_LAUNCH(owned, ScInitOp, sc_fields(h), tmp);
        dev_sync(h->stream);
    } catch (...) { dev_free(tmp); throw; }
    dev_free(tmp);
    SCState* s = (SCState*)h->sc;
    s->efs_prepared = false; s->head_done = false; s->rho_is_sum = false; s->uph_valid = true;
    h->has_state = true;
    return LBM_OK;
}

int sc_upload_state(lbm_handle* h, const double* const* pdf, const double* const* rho, int32_t n_comp) {
    if (n_comp != h->cfg.n_components || !pdf) { h->err = "one population array per component expected"; return LBM_EINVAL; }
    sc_alloc(h);
    const int64_t owned = h->g.plane * h->g.n2;
    const int Q = h->Q;
    double* tmp = (double*)dev_alloc((size_t)owned * (Q + 1) * 8);
    try {
        SCFields c = sc_fields(h);
        for (int k = 0; k < n_comp; ++k) {
            if (!pdf[k]) throw BackendError{"NULL population array"};
            dev_h2d(tmp, pdf[k], (size_t)owned * Q * 8, h->stream);
            const double* rin = nullptr;
            if (rho && rho[k]) { dev_h2d(tmp + owned * Q, rho[k], owned * 8, h->stream); rin = tmp + owned * Q; }
            SC_LAUNCH(owned, ScUploadOp, c, k, tmp, rin);
            dev_sync(h->stream);
        }
        SCState* s = (SCState*)h->sc;
        dev_zero(s->F, (size_t)n_comp * h->D * h->g.vol * 8, h->stream);
        s->efs_prepared = false; s->head_done = false; s->rho_is_sum = false; s->uph_valid = true;
    } catch (...) { dev_free(tmp); throw; }
    dev_free(tmp);
    h->has_state = true;
    return LBM_OK;
}

// inlet rows: Zou-He velocity per component + ghost row (OptimizedD2Q9GPU.py:839-861, 710-736)
// slab decomposition: the inlet rows live on the last slab, the outlet rows on the first one
static bool owns_inlet(const lbm_handle* h) { return h->rank == h->nranks - 1; }
static bool owns_outlet(const lbm_handle* h) { return h->rank == 0; }

static void sc_inlet(lbm_handle* h, const SCFields& c) {
    if (c.p.inlet == LBM_INLET_VELOCITY && owns_inlet(h)) {
        SC_LAUNCH(h->g.plane, ScInletVelocityOp, c);
        for (int zr = c.z_in; zr < c.z_in_ghost; ++zr)         // ghost rows, one after the other (ghostPointsConstantVelocity8/82)
            SC_LAUNCH(h->g.plane, ScRowCopyOp, c, zr + 1, zr);
    }
}
static void sc_outlet_pressure(lbm_handle* h, const SCFields& c) {
    if (!owns_outlet(h)) return;
    SC_LAUNCH(h->g.plane, ScOutletPressureOp, c);
    for (int zr = c.z_out; zr > 0; --zr)                        // ghostPointsConstantPressureOutlet8/82
        SC_LAUNCH(h->g.plane, ScRowCopyOp, c, zr - 1, zr);
}

static void sc_ensure_head(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    if (h->cfg.model != LBM_MODEL_SC || s->head_done) return;
    sc_inlet(h, sc_fields(h));
    s->head_done = true;
}

// one iteration of runOptimizedLBM's loop (ShanChenD2Q9.py:1492-1629), in two halves: what precedes the streaming ...
static void sc_head_collide(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    const Grid& g = h->g;
    SCFields c = sc_fields(h);
    // Inlet rows, then calFluidRhoGPU (psi = rho).  The streaming of the previous iteration left rho = sum_q f_q (same order
    // of summation) on every node and so do the copied rows; only the Zou-He plane of the inlet carries another value, and
    // the thread that treated the column re-sums it: no pass over the lattice.  A fresh state is summed once.
    const bool inlet_here = c.p.inlet == LBM_INLET_VELOCITY && owns_inlet(h);
    if (!s->rho_is_sum) {
        sc_ensure_head(h);
        SC_LAUNCH(g.count(0), ScRhoOp, c); s->rho_is_sum = true;
    } else if (inlet_here) {
        if (s->head_done) SC_LAUNCH_PLANES(c.z_in, c.z_in + 1, ScRhoOp, c);      // a download already treated the inlet rows
        else SC_LAUNCH(2 * g.plane, ScOpenRowsOp, c, 0, 1, 1, 0, 1);
        s->head_done = true;
    }
    exchange_f64(h, c.rho, g.vol, c.p.nc, 1);
    SC_LAUNCH(g.count(0), ScCollideOp, c);                  // interactionCollisionProcess
}
// ... and the streaming with what follows it
static void sc_tail(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    const Grid& g = h->g;
    SCFields c = sc_fields(h);
    exchange_f64(h, c.fC, g.vol, c.p.nc * h->Q, 1);
    SC_LAUNCH(g.count(0), ScStreamOp, c);                   // calStreaming1GPU/2GPU (+ densities)
    if (c.p.outlet == LBM_OUTLET_CONVECTIVE && owns_outlet(h))      // convectiveOutletGPU / Ghost2 / Ghost3
        SC_LAUNCH(2 * g.plane, ScOpenRowsOp, c, 0, 2, 0, 1, 0);
    // calPhysicalVelocity: an output, nothing in the loop reads it -> evaluated when somebody asks (sc_ensure_velocity)
    s->uph_valid = false;
    s->head_done = false;
}
static void sc_iteration(lbm_handle* h) {
    sc_head_collide(h);
    sc_tail(h);
}

// the physical velocity of the last finished iteration of the original Shan-Chen loop (ShanChenD2Q9.py:1561-1573); must
// run BEFORE the inlet treatment of the next iteration touches the inlet rows
static void sc_ensure_velocity(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    if (h->cfg.model != LBM_MODEL_SC || s->uph_valid) return;
    SCFields c = sc_fields(h);
    SC_LAUNCH(h->g.count(0), ScPhysicalVelocityOp, c);
    s->uph_valid = true;
}

// pre-loop of runOptimizedEFLBM (ShanChenD2Q9.py:1714-1849)
static void efs_prepare(lbm_handle* h) {
    SCState* s = (SCState*)h->sc;
    if (s->efs_prepared) return;
    const Grid& g = h->g;
    SCFields c = sc_fields(h);
    exchange_f64(h, c.rho, g.vol, c.p.nc, c.p.scheme == 4 ? 1 : NG);
    SC_LAUNCH(g.count(0), EfsForceOp, c);
    SC_LAUNCH(g.count(0), EfsTransformOp, c);
    // boundary rows: the populations are treated, the densities the reference sets here are never read
    // before calFluidRhoGPU overwrites them, so the state keeps the densities f_eq was built from
    if (c.p.inlet == LBM_INLET_VELOCITY || c.p.outlet == LBM_OUTLET_PRESSURE) {
        double* keep = (double*)dev_alloc((size_t)c.p.nc * g.vol * 8);
        dev_d2d(keep, c.rho, (size_t)c.p.nc * g.vol * 8, h->stream);
        sc_inlet(h, c);
        if (c.p.outlet == LBM_OUTLET_PRESSURE) sc_outlet_pressure(h, c);
        dev_d2d(c.rho, keep, (size_t)c.p.nc * g.vol * 8, h->stream);
        dev_sync(h->stream);
        dev_free(keep);
    }
    s->efs_prepared = true;
}

// one iteration of runOptimizedEFLBM's loop (ShanChenD2Q9.py:1852-2087), in two halves: the collision ...
static void efs_collide(lbm_handle* h) {
    const Grid& g = h->g;
    SCFields c = sc_fields(h);
    const bool convective = c.p.outlet == LBM_OUTLET_CONVECTIVE;
    if (convective && owns_outlet(h)) SC_LAUNCH(3 * g.plane, ScSaveRowsOp, c);   // savePDFLastStep
    SC_LAUNCH(g.count(0), EfsCollideOp, c);
}
// ... and the streaming, the boundary rows and the force of the next collision
static void efs_tail(lbm_handle* h) {
    const Grid& g = h->g;
    SCFields c = sc_fields(h);
    exchange_f64(h, c.fC, g.vol, c.p.nc * h->Q, 1);
    SC_LAUNCH(g.count(0), ScStreamOp, c);
    // boundary rows (convective-each | pressure outlet, velocity inlet) and calFluidRhoGPU after them: the streaming and the
    // copied rows already hold rho = sum_q f_q, only the two Zou-He planes carry the imposed values and are re-summed by the
    // thread that treated the column (ShanChenD2Q9.py:1933-2014)
    {
        const int do_out = (c.p.outlet != LBM_BC_PERIODIC && owns_outlet(h)) ? 1 : 0;
        const int do_in = (c.p.inlet == LBM_INLET_VELOCITY && owns_inlet(h)) ? 1 : 0;
        if (do_out || do_in) SC_LAUNCH(2 * g.plane, ScOpenRowsOp, c, 1, 0, do_in, do_out, 1);
    }
    exchange_f64(h, c.rho, g.vol, c.p.nc, c.p.scheme == 4 ? 1 : NG);
    SC_LAUNCH(g.count(0), EfsForceOp, c);               // also the physical velocity of the output point (:2016-2027)
}
static void efs_iteration(lbm_handle* h) {
    efs_collide(h);
    efs_tail(h);
}

// ---- two-pass form (sc_fast.cuh) ---------------------------------------------------------------------------------------
// Two components, isotropy 4, no convective outlet under explicit forcing (its rows read the populations of the previous
// iteration and the physical velocity of plane 3, which the two-pass form does not keep); everything else -- solids, velocity
// inlet, pressure outlet (explicit forcing) / convective outlet (original Shan-Chen), D2Q9 and D3Q19, slabs -- runs on it.
static bool sc_fast_eligible(const lbm_handle* h) {
    const lbm_config& cfg = h->cfg;
    if (cfg.flags & LBM_FLAG_GENERIC_KERNELS) return false;
    if (cfg.n_components != 2) return false;
    if (cfg.model == LBM_MODEL_EFS) {
        if (cfg.sc_isotropy == 8 || cfg.sc_isotropy == 10) return false;
        if (cfg.outlet == LBM_OUTLET_CONVECTIVE) return false;
    }
    return !h->has_solid || h->pull != nullptr;
}

// `m` >= 2 iterations; the state is the reference-ordered one (streamed populations, densities, force) at entry and at exit:
// the first collision and the last streaming run on the reference-ordered operators, the m - 1 streaming + collision pairs in
// between on the two passes.  The two population buffers of the state alternate as source and destination.
// resident 256-thread CTAs per SM the collision pass is compiled for: D2Q9 holds 2 x 9 populations and 2 x 8 neighbour densities
// (120 / 148 registers uncapped; 2 CTAs = a 128-register cap cost the explicit-forcing operator 64 bytes of spills), D3Q19 holds
// 2 x 19 + 2 x 18 values and spills hundreds of bytes under any cap.  LBM_SC_OCC = 1 | 2 | 3 overrides (measurement aid).
static int sc_occupancy(int Q) {
    static const int forced = [] { const char* e = getenv("LBM_SC_OCC"); return e ? atoi(e) : 0; }();
    if (forced >= 1 && forced <= 3) return forced;
    return Q == 9 ? 2 : 1;
}
template <class Op>
static void sc_launch_collide(lbm_handle* h, const Op& op) {
    const int64_t n = h->g.count(0);
    switch (sc_occupancy(h->Q)) {
        case 3: launch_occ<3>(op, n, h->stream); break;
        case 2: launch_occ<2>(op, n, h->stream); break;
        default: launch(op, n, h->stream);
    }
}
template <class L>
static void sc_fast_iterations(lbm_handle* h, int m) {
    SCState* s = (SCState*)h->sc;
    const Grid& g = h->g;
    const bool efs = h->cfg.model == LBM_MODEL_EFS;
    if (efs) efs_collide(h); else sc_head_collide(h);
    const SCFields c = sc_fields(h);
    const int do_in = (c.p.inlet == LBM_INLET_VELOCITY && owns_inlet(h)) ? 1 : 0;
    const int do_out = ((efs ? c.p.outlet == LBM_OUTLET_PRESSURE : c.p.outlet == LBM_OUTLET_CONVECTIVE) && owns_outlet(h)) ? 1 : 0;
    ScFast f;
    f.pull = h->has_solid ? h->pull : nullptr;
    f.mat_lo = do_out ? (efs ? c.z_out + 1 : 4) : 0;       // pressure outlet: planes 0 .. z_out; convective copies: 3 -> 2 -> 1 -> 0
    f.mat_hi = do_in ? g.n2 - c.z_in : 0;                  // velocity inlet: planes z_in .. n2 - 1
    // One slab with open ends: the planes next to them are pulled and treated on a second stream (a parallel branch of the replayed
    // graph) BESIDE the density pass of the other planes -- both only read the source buffer and write disjoint planes -- so the
    // serial chain of the row operators (one thread per column, ~10 us) is off the critical path.  Measured on BASELINE config 3:
    // 4 808 (forked) vs 4 871 MLUPS (serial) -- the chain is 7 % of that step and a second density launch costs what it hides, so the
    // serial order stays the default here (the colour-gradient D2Q9 tiles gained 11 % from the same move); LBM_SC_FORK=1 selects it.
    static const bool fork_wanted = [] { const char* e = getenv("LBM_SC_FORK"); return e ? atoi(e) != 0 : false; }();
    const bool fork = fork_wanted && h->nranks == 1 && (do_in || do_out) && g.n2 > f.mat_lo + f.mat_hi;
    auto fused = [&](double* src, double* dst) {
        exchange_f64(h, src, g.vol, c.p.nc * h->Q, 1);
        f.src = src; f.dst = dst;
        SCFields r = c;
        r.fS = dst;
        const ScOpenRowsOp<L> rows{r, efs ? 1 : 0, efs ? 0 : 3, do_in, do_out, 1};
        if (fork) {
            const ScPullDensityOp<L, 2> dens{c, f};
            side_stream_fork(h);
            launch(PlaneRangeOp<ScPullDensityOp<L, 2>>{dens, (int64_t)f.mat_lo * g.plane}, (int64_t)(g.n2 - f.mat_lo - f.mat_hi) * g.plane, h->stream);
            side_stream_swap(h);
            try {
                const int64_t cnt0 = (int64_t)f.mat_lo * g.plane, cnt1 = (int64_t)f.mat_hi * g.plane;
                launch(TwoRangeOp<ScPullDensityOp<L, 2>>{dens, 0, cnt0, (int64_t)(g.n2 - f.mat_hi) * g.plane}, cnt0 + cnt1, h->stream);
                launch(rows, 2 * g.plane, h->stream);
            } catch (...) { side_stream_swap(h); throw; }
            side_stream_swap(h);
            side_stream_join(h);
        } else {
            launch(ScPullDensityOp<L, 2>{c, f}, g.count(0), h->stream);
            if (do_in || do_out) launch(rows, 2 * g.plane, h->stream);
        }
        exchange_f64(h, c.rho, g.vol, c.p.nc, 1);
        if (efs) sc_launch_collide(h, EfsPullCollideOp<L, 2>{c, f});
        else sc_launch_collide(h, ScPullCollideOp<L, 2>{c, f});
    };
    double *A = s->fC, *B = s->fS;
    const int n = m - 1;
    replay(n / 2, h->graph_ok(), &h->graph, h->stream, [&] { fused(A, B); fused(B, A); });
    if (n & 1) {
        fused(A, B);
        s->fC = B; s->fS = A;
    }
    if (efs) efs_tail(h); else sc_tail(h);
}

// Persistent form of both loops (LBM_FLAG_PERSISTENT, opt-in; see cg_fast.cu::cg_fast_persistent): every iteration of an
// lbm_step call inside ONE cooperative kernel, a grid-wide barrier where the launch boundaries of sc_iteration /
// efs_iteration are.  128 x 128 nodes (BASELINE configuration 1) are sixteen thousand threads: two launches per step cost
// more than the step.  Same operators, same order.
template <class L>
__global__ void __launch_bounds__(256)
sc_persistent(const SCFields c, const int efs, const int nsteps, const int do_in, const int do_out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const Grid& g = c.g;
    const int64_t n_owned = g.count(0), n_rows = 2 * g.plane;
    const bool convective = c.p.outlet == LBM_OUTLET_CONVECTIVE && do_out;
    for (int step = 0; step < nsteps; ++step) {
        if (!efs) {
            if (do_in) {
                for (int64_t i = tid; i < n_rows; i += nth) ScOpenRowsOp<L>{c, 0, 1, 1, 0, 1}(i);
                LBM_GRID_SYNC();
            }
            for (int64_t i = tid; i < n_owned; i += nth) ScCollideOp<L>{c}(i);
            LBM_GRID_SYNC();
            for (int64_t i = tid; i < n_owned; i += nth) ScStreamOp<L>{c}(i);
            if (convective) {
                LBM_GRID_SYNC();
                for (int64_t i = tid; i < n_rows; i += nth) ScOpenRowsOp<L>{c, 0, 2, 0, 1, 0}(i);
            }
            LBM_GRID_SYNC();
        } else {
            if (convective)
                for (int64_t i = tid; i < 3 * g.plane; i += nth) ScSaveRowsOp<L>{c}(i);      // reads what the collision reads
            for (int64_t i = tid; i < n_owned; i += nth) EfsCollideOp<L>{c}(i);
            LBM_GRID_SYNC();
            for (int64_t i = tid; i < n_owned; i += nth) ScStreamOp<L>{c}(i);
            LBM_GRID_SYNC();
            if (do_in || do_out) {
                for (int64_t i = tid; i < n_rows; i += nth) ScOpenRowsOp<L>{c, 1, 0, do_in, do_out, 1}(i);
                LBM_GRID_SYNC();
            }
            for (int64_t i = tid; i < n_owned; i += nth) EfsForceOp<L>{c}(i);
            LBM_GRID_SYNC();
        }
    }
}

template <class L>
static void sc_launch_persistent(lbm_handle* h, int nsteps) {
    SCState* s = (SCState*)h->sc;
    SCFields c = sc_fields(h);
    int efs = h->cfg.model == LBM_MODEL_EFS ? 1 : 0;
    int do_in = (c.p.inlet == LBM_INLET_VELOCITY && owns_inlet(h)) ? 1 : 0;
    int do_out = (c.p.outlet != LBM_BC_PERIODIC && owns_outlet(h)) ? 1 : 0;
#ifdef LBM_HOSTCHECK
    cta_emu::launch_cooperative(dim3(3), dim3(32), [&] { sc_persistent<L>(c, efs, nsteps, do_in, do_out); });
#else
    static std::atomic<int> grid_for_device[64];
    int grid = grid_for_device[h->cfg.device & 63];
    if (!grid) {
        int per_sm = 0, sms = 0;
        LBM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sc_persistent<L>, 256, 0));
        LBM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if (per_sm < 1) throw BackendError{"the persistent kernel does not fit on an SM"};
        grid = per_sm * sms;
        grid_for_device[h->cfg.device & 63] = grid;
    }
    void* args[] = {&c, &efs, &nsteps, &do_in, &do_out};
    LBM_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)sc_persistent<L>, dim3(grid), dim3(256), args, 0, h->stream));
#endif
    ++g_launch_counter;
    s->head_done = false; s->uph_valid = efs ? s->uph_valid : false;
}

void sc_step(lbm_handle* h, int nsteps) {
    if (h->cfg.model == LBM_MODEL_EFS) efs_prepare(h);
    auto one = [&] { if (h->cfg.model == LBM_MODEL_SC) sc_iteration(h); else efs_iteration(h); };
    if (nsteps <= 0) return;
    one();      // outside the graph: whether the inlet treatment of this iteration is still due depends on the host state
    if (nsteps > 1 && (h->cfg.flags & LBM_FLAG_PERSISTENT) && h->nranks == 1 && h->g.wrap2 && !g_prof_active()) {
        if (h->Q == 9) sc_launch_persistent<D2Q9>(h, nsteps - 1); else sc_launch_persistent<D3Q19>(h, nsteps - 1);
        return;
    }
    if (nsteps > 2 && sc_fast_eligible(h)) {
        if (h->Q == 9) sc_fast_iterations<D2Q9>(h, nsteps - 1); else sc_fast_iterations<D3Q19>(h, nsteps - 1);
        return;
    }
    replay(nsteps - 1, h->graph_ok(), &h->graph, h->stream, one);
}

int sc_download_macros(lbm_handle* h, double* const* rho, int32_t n_comp, double* const* u) {
    if (rho && n_comp != h->cfg.n_components) { h->err = "one density array per component expected"; return LBM_EINVAL; }
    if (h->cfg.model == LBM_MODEL_EFS) efs_prepare(h);
    sc_ensure_velocity(h);
    sc_ensure_head(h);
    SCFields c = sc_fields(h);
    const Grid& g = h->g;
    const int64_t owned = g.plane * g.n2, off = NG * g.plane;
    if (rho)
        for (int k = 0; k < n_comp; ++k)
            if (rho[k]) dev_d2h(rho[k], c.rho + k * g.vol + off, owned * 8, h->stream);
    if (u)
        for (int a = 0; a < h->D; ++a)
            if (u[a]) dev_d2h(u[a], c.uph + a * g.vol + off, owned * 8, h->stream);
    dev_sync(h->stream);
    return LBM_OK;
}

void sc_output_pointers(lbm_handle* h, const double** rho, const double** u) {
    if (h->cfg.model == LBM_MODEL_EFS) efs_prepare(h);
    sc_ensure_velocity(h);
    sc_ensure_head(h);
    SCFields c = sc_fields(h);
    *rho = c.rho; *u = c.uph;
}

int sc_download_pdfs(lbm_handle* h, double* const* pdf, int32_t n_comp) {
    if (!pdf || n_comp != h->cfg.n_components) { h->err = "one population array per component expected"; return LBM_EINVAL; }
    if (h->cfg.model == LBM_MODEL_EFS) efs_prepare(h);
    sc_ensure_velocity(h);
    sc_ensure_head(h);
    SCFields c = sc_fields(h);
    const int64_t owned = h->g.plane * h->g.n2;
    double* tmp = (double*)dev_alloc((size_t)owned * h->Q * 8);
    try {
        for (int k = 0; k < n_comp; ++k) {
            if (!pdf[k]) continue;
            SC_LAUNCH(owned, ScDownloadOp, c, k, tmp);
            dev_d2h(pdf[k], tmp, (size_t)owned * h->Q * 8, h->stream);
        }
    } catch (...) { dev_free(tmp); throw; }
    dev_free(tmp);
    return LBM_OK;
}

int sc_total_mass(lbm_handle* h, double* mass, int32_t n_comp) {
    if (n_comp != h->cfg.n_components) { h->err = "one value per component expected"; return LBM_EINVAL; }
    SCFields c = sc_fields(h);
    const Grid& g = h->g;
    const int64_t owned = g.plane * g.n2;
    std::vector<double> buf((size_t)owned);
    for (int k = 0; k < n_comp; ++k) {
        dev_d2h(buf.data(), c.rho + k * g.vol + NG * g.plane, owned * 8, h->stream);
        long double s = 0.0L;
        for (int64_t i = 0; i < owned; ++i) s += buf[i];
        mass[k] = (double)s;
    }
    return LBM_OK;
}

}  // namespace lbm
