// handle.h -- the state behind an lbm_handle (device arrays are owned by the handle).
#pragma once
#include <string>
#include <vector>

#include "../../include/lbmpm.h"
#include "backend.h"
#include "cg_ops.cuh"

struct lbm_handle {
    lbm_config cfg;
    lbm::Grid g;
    int D = 0, Q = 0;
    lbm::stream_t stream = 0;
    int rank = 0, nranks = 1;
    std::string err;

    // geometry
    uint8_t* dom = nullptr;     // [vol] 1 = void
    uint8_t* cls = nullptr;     // [vol] node classes
    double* ns = nullptr;       // [3][vol]
    uint32_t* pull = nullptr;   // [vol] pull masks of the tiled kernels (grid.cuh::PullMaskOp; D3Q19 with solids only)
    int64_t* wet_list = nullptr;   // flat ids of the wetting solids of planes [-2, n2 + 2) (tiled kernels with solids only)
    int64_t n_wet_list = 0;
    int64_t n_fluid = 0, n_wet = 0, n_near = 0;
    bool has_geometry = false;
    bool has_solid = false;     // any solid node in the slab or its ghost planes

    // colour-gradient state (general path)
    double* fS = nullptr;       // [2][Q][vol] streamed populations
    double* fC = nullptr;       // [2][Q][vol] post-collision populations
    double* rho = nullptr;      // [2][vol]
    double* u = nullptr;        // [3][vol]
    double* phi = nullptr;
    double* G = nullptr;        // [3][vol]
    double* nrm = nullptr;      // [3][vol]
    double* F = nullptr;        // [3][vol]
    double* K = nullptr;
    bool has_state = false;
    bool head_done = false;     // boundary treatment + velocity + phi of the current iteration already applied

    // fast path (cg_fast.cu): state kept post-collision in factored form
    void* fast = nullptr;
    bool fast_pending_stream = false;

    // Shan-Chen / explicit forcing state (sc_ops)
    void* sc = nullptr;

    // solute tracers riding on the colour-gradient flow (tr_api.cu)
    void* tracer = nullptr;

    // launch-bound lattices replay a captured CUDA graph of the step (backend.h::replay)
    lbm::GraphKeep graph;
    bool graph_ok() const { return nranks == 1 && g.plane * (int64_t)g.n2 <= (int64_t)1 << 22 && !(cfg.flags & 8u); }

    // asynchronous output (lbm_download_macros_async): device staging buffer [n_comp + D][owned]
    double* out_stage = nullptr;
    size_t out_stage_bytes = 0;

    void* nccl = nullptr;       // ncclComm_t (host test hook: the in-process ring of host_stubs.cu)
    void* peer = nullptr;       // PeerState of comm.cu / host_stubs.cu: peer pointers and flags of the one-sided exchange
    uint64_t peer_epoch = 0;    // number of one-sided exchanges so far (identical on every rank)
    bool peer_unreachable = false;   // a neighbour slab did not confirm that it dropped its mappings of my arrays (see comm_peer_release)
    bool peer_ok = false;       // decided collectively in lbm_set_geometry: the fast path's two per-step exchanges are stores into
                                // the neighbours' memory + flags (default on slabs) instead of NCCL send / recv

    // measurement
    double last_ms = 0.0;
    int64_t last_launches = 0;
#ifndef LBM_HOSTCHECK
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // slab decomposition: ghost-plane exchanges run on their own stream, overlapped with the interior planes
    cudaStream_t comm_stream = nullptr, xstream = nullptr;   // xstream: where exchange_* currently enqueues (null = stream)
    cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
    cudaStream_t out_stream = nullptr;                       // device -> host copies of the asynchronous output
    cudaEvent_t ev_out_ready = nullptr, ev_out_done = nullptr;
    double *stage_send = nullptr, *stage_recv = nullptr;     // packed boundary planes (comm.cu)
    size_t stage_bytes = 0;
#endif

    lbm::CGFields fields() const;
};
