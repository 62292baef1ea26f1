// cg_fast.cu -- fused fast path (placeholder: not yet enabled)
#ifndef LBM_HOSTCHECK
#include "internal.h"
namespace lbm {
bool cg_fast_eligible(const lbm_handle*) { return false; }
void cg_fast_step(lbm_handle*, int) {}
void cg_fast_materialise(lbm_handle*) {}
void cg_fast_free(lbm_handle*) {}
}  // namespace lbm
#endif
