// tr_api.cu -- solute tracers behind the C ABI (lbm_tracer_*): state, the tracer phase of one iteration, downloads.
#include "internal.h"
#include "tr_ops.cuh"

namespace lbm {

struct TracerState {
    TracerParams p;
    double *g = nullptr, *gC = nullptr, *conc = nullptr;
    bool inlet = false, outlet = false;      // 5-velocity branch: Inamuro inlet row / free-flow outlet row requested
    bool has_state = false;
    bool phase_done = false;      // the tracer phase of the current iteration has already run (a download asked for it)
};

static TracerFields tracer_fields(const lbm_handle* h) {
    const TracerState* s = (const TracerState*)h->tracer;
    TracerFields t;
    t.p = s->p; t.g = s->g; t.gC = s->gC; t.conc = s->conc;
    // the rows of the tracer boundaries live on the first / last slab (rank / nranks are known once the geometry is set)
    t.p.inlet_row = s->inlet && h->rank == h->nranks - 1 ? h->g.n2 - 1 : -1;
    t.p.outlet_row = s->outlet && h->rank == 0 ? 0 : -1;
    return t;
}

void tracer_free(lbm_handle* h) {
    TracerState* s = (TracerState*)h->tracer;
    if (!s) return;
    dev_free(s->g); dev_free(s->gC); dev_free(s->conc);
    delete s;
    h->tracer = nullptr;
}

// between the colour gradient and the flow collision of one iteration (Transport2DRK.py:1341-1425)
void tracer_phase(lbm_handle* h) {
    TracerState* s = (TracerState*)h->tracer;
    if (!s || !s->has_state || s->phase_done) return;
    const Grid& g = h->g;
    CGFields c = h->fields();
    TracerFields t = tracer_fields(h);
    if (s->p.schemes == 5) {
        launch(TracerCollideQ5Op{c, t}, g.count(0), h->stream);
        if (t.p.outlet_row >= 0) launch(TracerFreeflowOp{c, t}, g.plane, h->stream);
        exchange_f64(h, s->gC, g.vol, s->p.nt * 5, 1);
        launch(TracerStreamOp<D2Q5>{c, t}, g.count(0), h->stream);
    } else {
        if (h->Q == 9) launch(TracerCollideOp<D2Q9>{c, t}, g.count(0), h->stream);
        else launch(TracerCollideOp<D3Q19>{c, t}, g.count(0), h->stream);
        exchange_f64(h, s->gC, g.vol, s->p.nt * h->Q, 1);
        if (h->Q == 9) launch(TracerStreamOp<D2Q9>{c, t}, g.count(0), h->stream);
        else launch(TracerStreamOp<D3Q19>{c, t}, g.count(0), h->stream);
    }
    s->phase_done = true;
}
void tracer_iteration_finished(lbm_handle* h) {
    if (h->tracer) ((TracerState*)h->tracer)->phase_done = false;
}

}  // namespace lbm

using namespace lbm;

#ifndef LBM_HOSTCHECK
#define TR_SET_DEVICE(h) LBM_CUDA_CHECK(cudaSetDevice((h)->cfg.device))
#else
#define TR_SET_DEVICE(h) (void)0
#endif
#define TR_API_BEGIN(h)                \
    if (!(h)) return LBM_EINVAL;       \
    try {                              \
        TR_SET_DEVICE(h);
#define TR_API_END(h)                                                        \
    }                                                                        \
    catch (const BackendError& e) { (h)->err = e.msg; return e.oom ? LBM_ENOMEM : LBM_ECUDA; } \
    catch (const std::exception& e) { (h)->err = e.what(); return LBM_ECUDA; }                 \
    return LBM_OK;

extern "C" int lbm_tracer_setup(lbm_handle* h, const lbm_tracer_config* cfg) {
    TR_API_BEGIN(h)
    if (!cfg) { h->err = "null tracer configuration"; return LBM_EINVAL; }
    if (h->cfg.model != LBM_MODEL_CG || h->cfg.surface_tension_type != LBM_ST_CSF) {
        h->err = "tracers ride on the colour-gradient CSF flow (runTransport2DMPMCRKNew)"; return LBM_EINVAL;
    }
    const int schemes = cfg->n_schemes == 0 ? 9 : cfg->n_schemes;
    if (schemes != 9 && schemes != 5) { h->err = "n_schemes must be 9 or 5"; return LBM_EINVAL; }
    if (schemes == 9 && (h->cfg.inlet != LBM_BC_PERIODIC || h->cfg.outlet != LBM_BC_PERIODIC)) {
        h->err = "9-velocity tracers: closed boxes only (that branch of the reference has no inlet / outlet treatment)"; return LBM_EINVAL;
    }
    if (cfg->n_tracers < 1 || cfg->n_tracers > TR_MAX) { h->err = "n_tracers must be 1..4"; return LBM_EINVAL; }
    if (cfg->relax != LBM_RELAX_SRT && cfg->relax != LBM_RELAX_MRT) { h->err = "tracer relax must be SRT or MRT"; return LBM_EINVAL; }
    if (cfg->relax == LBM_RELAX_MRT && h->Q != 9) { h->err = "the tracer MRT is defined for D2Q9 (Transport2DRK.py:367-391); use SRT on D3Q19"; return LBM_EINVAL; }
    if (schemes == 5) {
        if (h->Q != 9) { h->err = "the 5-velocity tracer lattice is two-dimensional (D2Q9 flow)"; return LBM_EINVAL; }
        if (cfg->relax != LBM_RELAX_MRT) { h->err = "the 5-velocity branch of the reference collides with MRT only (Transport2DRK.py:1345-1351)"; return LBM_EINVAL; }
        if (cfg->reaction && cfg->n_tracers != 3) { h->err = "the reaction A + B -> C needs 3 tracers (calReactionTracersGPU)"; return LBM_EINVAL; }
        if (cfg->inlet_type != LBM_TR_NONE && cfg->inlet_type != LBM_TR_INLET_DIRICHLET) { h->err = "unknown tracer inlet type"; return LBM_EINVAL; }
        if (cfg->outlet_type != LBM_TR_NONE && cfg->outlet_type != LBM_TR_OUTLET_FREEFLOW) { h->err = "unknown tracer outlet type"; return LBM_EINVAL; }
    } else if (cfg->reaction || cfg->inlet_type || cfg->outlet_type) {
        h->err = "reactions and tracer inlet / outlet rows belong to the 5-velocity branch (n_schemes = 5)"; return LBM_EINVAL;
    }
    if (h->has_state) { h->err = "lbm_tracer_setup must precede lbm_init_equilibrium / lbm_upload_state"; return LBM_ESTATE; }
    tracer_free(h);
    TracerState* s = new TracerState();
    h->tracer = s;
    s->p.nt = cfg->n_tracers; s->p.relax = cfg->relax; s->p.criterion = cfg->criterion;
    s->p.schemes = schemes; s->p.reaction = cfg->reaction ? 1 : 0; s->p.rate = cfg->reaction_rate;
    s->p.inlet_row = s->p.outlet_row = -1;
    s->inlet = cfg->inlet_type == LBM_TR_INLET_DIRICHLET; s->outlet = cfg->outlet_type == LBM_TR_OUTLET_FREEFLOW;
    for (int k = 0; k < TR_MAX; ++k) {
        s->p.j0[k] = cfg->diff_j[k]; s->p.inlet_conc[k] = cfg->inlet_conc[k];
        s->p.tau[k] = cfg->tau[k]; s->p.beta[k] = cfg->beta[k];
        s->p.sa[k] = 0.5 + 3.0 * cfg->dxx[k]; s->p.sd[k] = 0.5 + 3.0 * cfg->dyy[k];
        s->p.sb[k] = 3.0 * cfg->dxy[k]; s->p.sc[k] = 3.0 * cfg->dyx[k];
        if (k < cfg->n_tracers) {
            if (cfg->relax == LBM_RELAX_SRT && !(cfg->tau[k] > 0.5)) { h->err = "tracer tau must exceed 1/2"; tracer_free(h); return LBM_EINVAL; }
            if (cfg->relax == LBM_RELAX_MRT && !(s->p.sa[k] * s->p.sd[k] - s->p.sb[k] * s->p.sc[k] > 0.0)) {
                h->err = "tracer diffusion tensor is not positive"; tracer_free(h); return LBM_EINVAL;
            }
        }
    }
    TR_API_END(h)
}

extern "C" int lbm_tracer_init(lbm_handle* h, const double* const* conc, int32_t n) {
    TR_API_BEGIN(h)
    TracerState* s = (TracerState*)h->tracer;
    if (!s) { h->err = "lbm_tracer_setup has not been called"; return LBM_ESTATE; }
    if (!h->has_state) { h->err = "set the flow state first"; return LBM_ESTATE; }
    if (!conc || n != s->p.nt) { h->err = "one concentration array per tracer expected"; return LBM_EINVAL; }
    for (int k = 0; k < n; ++k) if (!conc[k]) { h->err = "NULL concentration array"; return LBM_EINVAL; }
    const Grid& g = h->g;
    const int64_t owned = g.plane * g.n2;
    const size_t bytes = (size_t)s->p.nt * (s->p.schemes == 5 ? 5 : h->Q) * g.vol * sizeof(double);
    if (!s->g) {
        s->g = (double*)dev_alloc(bytes); s->gC = (double*)dev_alloc(bytes);
        s->conc = (double*)dev_alloc((size_t)s->p.nt * g.vol * sizeof(double));
    }
    dev_zero(s->g, bytes, h->stream); dev_zero(s->gC, bytes, h->stream);
    dev_zero(s->conc, (size_t)s->p.nt * g.vol * sizeof(double), h->stream);
    double* tmp = (double*)dev_alloc((size_t)n * owned * 8);
    try {
        for (int k = 0; k < n; ++k) dev_h2d(tmp + k * owned, conc[k], owned * 8, h->stream);
        CGFields c = h->fields();
        TracerFields t = tracer_fields(h);
        if (s->p.schemes == 5) launch(TracerInitOp<D2Q5>{c, t, tmp}, owned, h->stream);
        else if (h->Q == 9) launch(TracerInitOp<D2Q9>{c, t, tmp}, owned, h->stream);
        else launch(TracerInitOp<D3Q19>{c, t, tmp}, owned, h->stream);
        dev_sync(h->stream);
    } catch (...) { dev_free(tmp); throw; }
    dev_free(tmp);
    s->has_state = true; s->phase_done = false;
    TR_API_END(h)
}

extern "C" int lbm_tracer_download(lbm_handle* h, double* const* conc, int32_t n) {
    TR_API_BEGIN(h)
    TracerState* s = (TracerState*)h->tracer;
    if (!s || !s->has_state) { h->err = "no tracer state"; return LBM_ESTATE; }
    if (!conc || n != s->p.nt) { h->err = "one concentration array per tracer expected"; return LBM_EINVAL; }
    // the output point of the reference: after the tracer phase of the current iteration (Transport2DRK.py:1427-1437)
    cg_fast_materialise(h);       // no-op unless the fast path left the streaming pending
    cg_ensure_head(h);
    cg_generic_forces(h);
    tracer_phase(h);
    const Grid& g = h->g;
    const int64_t owned = g.plane * g.n2, off = NG * g.plane;
    for (int k = 0; k < n; ++k)
        if (conc[k]) dev_d2h(conc[k], s->conc + (int64_t)k * g.vol + off, owned * 8, h->stream);
    dev_sync(h->stream);
    TR_API_END(h)
}
