// sc_api.cu -- Shan-Chen / explicit-forcing models (placeholder until sc_ops.cuh lands)
#include "internal.h"
namespace lbm {
int sc_init_equilibrium(lbm_handle* h, const double* const*, int32_t) { h->err = "Shan-Chen models not built yet"; return LBM_EINVAL; }
int sc_upload_state(lbm_handle* h, const double* const*, const double* const*, int32_t) { h->err = "Shan-Chen models not built yet"; return LBM_EINVAL; }
void sc_step(lbm_handle*, int) {}
int sc_download_macros(lbm_handle*, double* const*, int32_t, double* const*) { return LBM_EINVAL; }
int sc_download_pdfs(lbm_handle*, double* const*, int32_t) { return LBM_EINVAL; }
int sc_total_mass(lbm_handle*, double*, int32_t) { return LBM_EINVAL; }
void sc_free(lbm_handle*) {}
}  // namespace lbm
