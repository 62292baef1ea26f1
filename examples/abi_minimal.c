/* abi_minimal.c -- liblbmpm.so used from plain C through include/lbmpm.h: a D3Q19 colour-gradient MRT box with a solid
 * sphere, 20 steps, mass check.  Build and run on a GPU box:
 *     gcc -std=c99 -Iinclude examples/abi_minimal.c -Lopenlbmpm_b200 -llbmpm -Wl,-rpath,$PWD/openlbmpm_b200 -lm -o abi_minimal
 *     ./abi_minimal
 * (the CPU test tier links the same file against the host test hook, tests/test_abi.py). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lbmpm.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc_ = (call);                                                                    \
        if (rc_ != LBM_OK) {                                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lbm_last_error(h));                \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(void) {
    const int nx = 32, ny = 16, nz = 24;
    const size_t n = (size_t)nx * ny * nz;
    lbm_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = LBM_ABI_VERSION;
    cfg.lattice = 19; cfg.model = LBM_MODEL_CG; cfg.relax = LBM_RELAX_MRT;
    cfg.nx = nx; cfg.ny = ny; cfg.nz = nz;
    cfg.tau_type = 2; cfg.wetting_type = 2;
    cfg.sigma = 0.1; cfg.contact_angle_deg = 60.0; cfg.beta = 0.7; cfg.delta = 0.98; cfg.tauR = 1.0; cfg.tauB = 1.0;

    lbm_handle* h = NULL;
    if (lbm_create(&cfg, &h) != LBM_OK) { fprintf(stderr, "lbm_create: %s\n", lbm_last_error(NULL)); return 1; }

    unsigned char* dom = malloc(n);
    double* rhoR = malloc(n * sizeof(double)); double* rhoB = malloc(n * sizeof(double));
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = ((size_t)z * ny + y) * nx + x;
                const double r2 = pow(x - 15.5, 2) + pow(y - 7.5, 2) + pow(z - 11.5, 2);
                dom[i] = r2 > 16.0;                          /* a solid sphere of radius 4 */
                rhoR[i] = dom[i] ? (z < nz / 2 ? 1.0 : 0.0) : 0.0;
                rhoB[i] = dom[i] ? (z < nz / 2 ? 0.0 : 1.0) : 0.0;
            }
    CHECK(lbm_set_geometry(h, dom));
    const double* rho_in[2] = {rhoR, rhoB};
    CHECK(lbm_init_equilibrium(h, rho_in, 2));
    double m0[2], m1[2];
    CHECK(lbm_total_mass(h, m0, 2));
    CHECK(lbm_step(h, 20));
    double* ux = malloc(n * sizeof(double));
    double* rho_out[2] = {rhoR, rhoB};
    double* u_out[3] = {ux, NULL, NULL};
    CHECK(lbm_download_macros(h, rho_out, 2, u_out));
    CHECK(lbm_total_mass(h, m1, 2));
    int64_t n_fluid = 0, n_wet = 0, n_near = 0;
    CHECK(lbm_index_sizes(h, &n_fluid, &n_wet, &n_near));
    printf("void nodes %lld, wetting solids %lld, mass R %.12f -> %.12f, mass B %.12f -> %.12f\n",
           (long long)n_fluid, (long long)n_wet, m0[0], m1[0], m0[1], m1[1]);
    const int ok = fabs(m1[0] - m0[0]) < 1e-9 * m0[0] && fabs(m1[1] - m0[1]) < 1e-9 * m0[1];
    lbm_destroy(h);
    free(dom); free(rhoR); free(rhoB); free(ux);
    puts(ok ? "abi_minimal ok" : "abi_minimal FAILED");
    return ok ? 0 : 1;
}
