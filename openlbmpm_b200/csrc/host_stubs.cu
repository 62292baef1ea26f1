// host_stubs.cu -- compiled ONLY into the host test hook (tests/hostcheck): the pieces that exist
// solely as CUDA code (tiled collision kernel, NCCL halo exchange) are absent there.
#ifdef LBM_HOSTCHECK
#include "internal.h"
namespace lbm {
void comm_exchange_f64(lbm_handle*, double*, int64_t, int, int, const int8_t*) { throw BackendError{"multi-rank needs the CUDA build"}; }
void comm_exchange_u8(lbm_handle*, uint8_t*, int) { throw BackendError{"multi-rank needs the CUDA build"}; }
void comm_destroy(lbm_handle*) {}
int comm_allreduce_max(lbm_handle*, int v) { return v; }
}  // namespace lbm
// test hook: the two forms of the 3-D Akai wetting correction (reference-ordered / the tiled kernels' fast form)
extern "C" void hostcheck_wetting3(const double* G, const double* ns, double cosT, double sinT, int fast, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        double g[3] = {G[3 * i], G[3 * i + 1], G[3 * i + 2]};
        if (fast) lbm::cg_wetting_akai3_fast(g, ns + 3 * i, cosT, sinT);
        else lbm::cg_wetting<3>(g, ns + 3 * i, cosT, sinT, 2);
        for (int a = 0; a < 3; ++a) out[3 * i + a] = g[a];
    }
}
extern "C" int lbm_nccl_unique_id(uint8_t*) { return LBM_ENCCL; }
extern "C" int lbm_comm_init(lbm_handle*, int32_t, int32_t, const uint8_t*) { return LBM_ENCCL; }
#endif
