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
extern "C" int lbm_nccl_unique_id(uint8_t*) { return LBM_ENCCL; }
extern "C" int lbm_comm_init(lbm_handle*, int32_t, int32_t, const uint8_t*) { return LBM_ENCCL; }
#endif
