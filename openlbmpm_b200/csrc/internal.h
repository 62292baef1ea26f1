// internal.h -- functions shared between the translation units of liblbmpm.so
#pragma once
#include "handle.h"

namespace lbm {

// ghost planes along the slab axis (periodic wrap on one GPU, NCCL send/recv between slabs)
void exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs = nullptr);
void exchange_u8(lbm_handle* h, uint8_t* base, int gp);
void comm_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs);
void comm_exchange_u8(lbm_handle* h, uint8_t* base, int gp);
// one-sided form (LBM_FLAG_PEER_EXCHANGE): boundary planes stored into the neighbours' ghost planes through peer pointers,
// then signal (release) / wait (acquire) on a flag pair.  `base` must be the start of a device allocation.
void comm_peer_exchange_f64(lbm_handle* h, double* base, int64_t stride, int narr, int gp, const int8_t* dirs);
// its two halves, for passes that store into the neighbours themselves: the neighbours' images of an allocation (mapped on
// first use), and the flag handshake that closes an exchange
void comm_peer_pointers(lbm_handle* h, double* base, double** up, double** down);
void comm_peer_signal_wait(lbm_handle* h);
// Allocations the neighbours have mapped are about to be freed (re-initialisation, new geometry, destruction).  Nothing is in
// flight between lbm_step calls (cg_fast_step closes its last handshake), so this only has to make sure nobody frees under a
// mapping: my mappings of the neighbours' arrays are dropped, "dropped" is signalled into their flag words, and theirs are
// awaited -- BOUNDED (LBM_PEER_CLOSE_TIMEOUT_MS, default 3 s): lbm_destroy must not hang on a rank that is gone.
// -> true: the neighbours confirmed, the exported arrays may be freed; false: they may still be mapped, do not free them.
bool comm_peer_release(lbm_handle* h);
bool comm_peer_probe(lbm_handle* h);     // collective: can every slab map its neighbours' memory (CUDA IPC over NVLink / PCIe P2P)?
void comm_peer_check(lbm_handle* h);     // throws if a bounded wait of the one-sided exchange timed out (call after a stream sync)
void comm_destroy(lbm_handle* h);
int comm_allreduce_max(lbm_handle* h, int v);

// A second stream beside the handle's for work that is independent of what the main stream runs next (captured into the same
// graph as a parallel branch when the step is being captured); created on first use, destroyed with the handle (comm_destroy).
// fork: the side stream waits for what the main stream has been given so far; swap: launches that go to h->stream go to the
// side stream (swap again to restore); join: the main stream waits for the side stream.  No-ops on the host test hook.
void side_stream_fork(lbm_handle* h);
void side_stream_swap(lbm_handle* h);
void side_stream_join(lbm_handle* h);

// general colour-gradient path (lbm_api.cu)
void cg_alloc_state(lbm_handle* h);
void cg_alloc_postcollision(lbm_handle* h);
void cg_ensure_head(lbm_handle* h);
void cg_apply_open_rows(lbm_handle* h);      // inlet / outlet treatment of the streamed populations (no-op on periodic boxes)
void cg_generic_body(lbm_handle* h);
void cg_generic_forces(lbm_handle* h);

// fused fast path for closed boxes (cg_fast.cu)
bool cg_fast_eligible(const lbm_handle* h);
bool cg_tiled_possible(const lbm_handle* h);     // lattice extents and flags admit the tiled kernels (they read ghost planes)
void cg_fast_step(lbm_handle* h, int nsteps);
void cg_fast_materialise(lbm_handle* h);
void cg_fast_free(lbm_handle* h);
void cg_fast_reset(lbm_handle* h);     // new state on the same geometry: keep the factored buffers, forget their contents

// Shan-Chen / explicit-forcing models (sc_api.cu)
int sc_init_equilibrium(lbm_handle* h, const double* const* rho, int32_t n_comp);
int sc_upload_state(lbm_handle* h, const double* const* pdf, const double* const* rho, int32_t n_comp);
void sc_step(lbm_handle* h, int nsteps);
int sc_download_macros(lbm_handle* h, double* const* rho, int32_t n_comp, double* const* u);
int sc_download_pdfs(lbm_handle* h, double* const* pdf, int32_t n_comp);
int sc_total_mass(lbm_handle* h, double* mass, int32_t n_comp);
void sc_output_pointers(lbm_handle* h, const double** rho, const double** u);   // device arrays [n_comp][vol], [D][vol] at the output point
void sc_free(lbm_handle* h);

// solute tracers (tr_api.cu)
void tracer_phase(lbm_handle* h);                 // no-op without tracers or when already run for this iteration
void tracer_iteration_finished(lbm_handle* h);
void tracer_free(lbm_handle* h);

}  // namespace lbm
