/* CPU ORACLE (test infrastructure, NOT product code) -- C + OpenMP restatement of the reference's
 * colour-gradient CSF loop, one function per reference kernel, in the reference's launch order
 * (/root/reference/RKCG2D/RKD2Q9.py:1295-1490; kernels in RKCG2D/AcceleratedRKGPU2D.py, commit 3d84189).
 *
 * It keeps the reference's arithmetic: populations as array-of-structures [node][Q], one pass over the
 * lattice per kernel, dense Q x Q matrix-vector products with M and M^-1 for the MRT collision and its
 * forcing term, equilibrium as feq(rhoR) + feq(rhoB).  Only the node addressing differs: a dense periodic
 * [z][y][x] grid with a void mask instead of the compact node list + int64 neighbour table, so that the
 * same code serves D2Q9 and D3Q19 (the reference ships no 3-D code; the D3Q19 tables come from
 * oracle/cg_dense.py).  Lattice tables, node classes and solid normals are
 * passed in from oracle/cg_dense.py.
 *
 * PINNED: tests/test_oracle_c.py checks it against oracle/cg_dense.py, which itself is pinned to the
 * reference's golden vectors (tests/golden/cg2d_*.npz) in tests/test_oracle_dense.py.
 * Used as (a) a checker at sizes NumPy is too slow for and (b) bench.py's cpu_baseline / --impl reference arm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QMAX 19

typedef struct {
    int D, Q, nx, ny, nz;
    int e[QMAX][3], opp[QMAX];
    double w[QMAX], enorm[QMAX], M[QMAX][QMAX], Mi[QMAX][QMAX], Sfix[QMAX]; /* Sfix < 0: slot takes 1/tau */
    double sigma, cosT, sinT, beta, delta, tauR, tauB;
    int tautype, wetting, relax;  /* relax: 0 SRT, 1 MRT */
    int64_t N;
    uint8_t *dom, *wet, *near;
    double *ns;                   /* [3][N] */
    double *fR, *fB, *fT, *fNewR, *fNewB;   /* [N][Q] */
    double *rhoR, *rhoB, *phi, *phiExt, *K; /* [N] */
    double *u, *F, *G, *n;        /* [3][N] */
} cgc;

static inline int64_t nbr(const cgc* c, int x, int y, int z, const int* e, int sgn) {
    int xn = x + sgn * e[0], yn = y + sgn * e[1], zn = z + sgn * e[2];
    if (xn < 0) xn += c->nx; else if (xn >= c->nx) xn -= c->nx;
    if (yn < 0) yn += c->ny; else if (yn >= c->ny) yn -= c->ny;
    if (zn < 0) zn += c->nz; else if (zn >= c->nz) zn -= c->nz;
    return ((int64_t)zn * c->ny + yn) * c->nx + xn;
}
#define FOR_NODES(c)                                        \
    _Pragma("omp parallel for collapse(2) schedule(static)") \
    for (int z = 0; z < (c)->nz; ++z)                       \
        for (int y = 0; y < (c)->ny; ++y)                   \
            for (int x = 0; x < (c)->nx; ++x)
#define ID(c) (((int64_t)z * (c)->ny + y) * (c)->nx + x)

cgc* cgc_create(int D, int Q, int nx, int ny, int nz, const int64_t* e, const int64_t* opp, const double* w,
                const double* M, const double* Mi, const double* Sfix, const uint8_t* dom, const uint8_t* wet,
                const uint8_t* near, const double* ns, double sigma, double theta_deg, double beta, double delta,
                double tauR, double tauB, int tautype, int wetting, int relax, int threads) {
    cgc* c = (cgc*)calloc(1, sizeof(cgc));
    c->D = D; c->Q = Q; c->nx = nx; c->ny = ny; c->nz = nz; c->N = (int64_t)nx * ny * nz;
    for (int i = 0; i < Q; ++i) {
        for (int a = 0; a < 3; ++a) c->e[i][a] = (int)e[i * 3 + a];
        c->opp[i] = (int)opp[i]; c->w[i] = w[i]; c->Sfix[i] = Sfix[i];
        c->enorm[i] = sqrt((double)(c->e[i][0] * c->e[i][0] + c->e[i][1] * c->e[i][1] + c->e[i][2] * c->e[i][2]));
        for (int j = 0; j < Q; ++j) { c->M[i][j] = M[i * Q + j]; c->Mi[i][j] = Mi[i * Q + j]; }
    }
    c->sigma = sigma; c->cosT = cos(theta_deg / 180.0 * M_PI); c->sinT = sin(theta_deg / 180.0 * M_PI);
    c->beta = beta; c->delta = delta; c->tauR = tauR; c->tauB = tauB;
    c->tautype = tautype; c->wetting = wetting; c->relax = relax;
    const int64_t N = c->N;
    c->dom = (uint8_t*)malloc(N); c->wet = (uint8_t*)malloc(N); c->near = (uint8_t*)malloc(N);
    memcpy(c->dom, dom, N); memcpy(c->wet, wet, N); memcpy(c->near, near, N);
    c->ns = (double*)malloc(3 * N * 8); memcpy(c->ns, ns, 3 * N * 8);
    double** pq[] = {&c->fR, &c->fB, &c->fT, &c->fNewR, &c->fNewB};
    for (int k = 0; k < 5; ++k) *pq[k] = (double*)calloc((size_t)N * Q, 8);
    double** p1[] = {&c->rhoR, &c->rhoB, &c->phi, &c->phiExt, &c->K};
    for (int k = 0; k < 5; ++k) *p1[k] = (double*)calloc((size_t)N, 8);
    double** p3[] = {&c->u, &c->F, &c->G, &c->n};
    for (int k = 0; k < 4; ++k) *p3[k] = (double*)calloc((size_t)N * 3, 8);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    return c;
}

void cgc_destroy(cgc* c) {
    void* p[] = {c->dom, c->wet, c->near, c->ns, c->fR, c->fB, c->fT, c->fNewR, c->fNewB, c->rhoR, c->rhoB,
                 c->phi, c->phiExt, c->K, c->u, c->F, c->G, c->n};
    for (unsigned k = 0; k < sizeof p / sizeof p[0]; ++k) free(p[k]);
    free(c);
}

/* initial condition: f = w rho at rest (RKD2Q9.py:561-585), lagged force zero */
void cgc_set_densities(cgc* c, const double* rhoR, const double* rhoB) {
    const int Q = c->Q;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        const double r = c->dom[id] ? rhoR[id] : 0.0, b = c->dom[id] ? rhoB[id] : 0.0;
        c->rhoR[id] = r; c->rhoB[id] = b;
        for (int i = 0; i < Q; ++i) { c->fR[id * Q + i] = c->w[i] * r; c->fB[id * Q + i] = c->w[i] * b; }
        for (int a = 0; a < 3; ++a) { c->F[a * c->N + id] = 0.0; c->u[a * c->N + id] = 0.0; }
    }
}

/* calTotalFluidPDF (AcceleratedRKGPU2D.py:1413-1422) */
static void total_pdf(cgc* c) {
    const int Q = c->Q;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        for (int i = 0; i < Q; ++i) c->fT[id * Q + i] = c->fR[id * Q + i] + c->fB[id * Q + i];
    }
}
/* calPhysicalVelocityRKGPU2DNew1 (2632-2653), calPhaseFieldPhi (1347-1356) */
static void velocity_phi(cgc* c) {
    const int Q = c->Q; const int64_t N = c->N;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) { c->phi[id] = 0.0; continue; }
        const double rho = c->rhoB[id] + c->rhoR[id];
        for (int a = 0; a < c->D; ++a) {
            double m = 0.0;
            for (int i = 1; i < Q; ++i) if (c->e[i][a]) m += c->e[i][a] * c->fT[id * Q + i];
            c->u[a * N + id] = (m + 0.5 * c->F[a * N + id]) / rho;
        }
        c->phi[id] = (c->rhoR[id] - c->rhoB[id]) / (c->rhoR[id] + c->rhoB[id]);
    }
}
/* calColorValueOnSolid (1559-1580) */
static void phi_on_solid(cgc* c) {
    const int Q = c->Q;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        c->phiExt[id] = c->dom[id] ? c->phi[id] : 0.0;
        if (!c->wet[id]) continue;
        double num = 0.0, den = 0.0;
        for (int i = 1; i < Q; ++i) {
            const int64_t nb = nbr(c, x, y, z, c->e[i], 1);
            if (c->dom[nb]) { num += c->w[i] * c->phi[nb]; den += c->w[i]; }
        }
        c->phiExt[id] = den > 0.0 ? num / den : 0.0;
    }
}
/* updateColorGradientOnWetting (1637-1679) / ...New (2428-2492) */
static void wetting_correction(const cgc* c, double* G, const double* ns) {
    const int D = c->D; const double ct = c->cosT, st = c->sinT;
    double gn = 0.0; for (int a = 0; a < D; ++a) gn += G[a] * G[a]; gn = sqrt(gn);
    if (c->wetting == 1) {
        const double n1x = ns[0] * ct - ns[1] * st, n1y = ns[1] * ct + ns[0] * st;
        const double n2x = ns[0] * ct + ns[1] * st, n2y = ns[1] * ct - ns[0] * st;
        double ux = 0.0, uy = 0.0;
        if (gn > 1.0e-8) { ux = G[0] / gn; uy = G[1] / gn; }
        const double d1 = sqrt((ux - n1x) * (ux - n1x) + (uy - n1y) * (uy - n1y));
        const double d2 = sqrt((ux - n2x) * (ux - n2x) + (uy - n2y) * (uy - n2y));
        double mx = 0.0, my = 0.0;
        if (d1 < d2) { mx = n1x; my = n1y; } else if (d1 > d2) { mx = n2x; my = n2y; } else if (d1 == d2) { mx = ns[0]; my = ns[1]; }
        G[0] = gn * mx; G[1] = gn * my;
        return;
    }
    double un[3] = {0, 0, 0}, dot = 0.0;
    if (gn > 1.0e-8) for (int a = 0; a < D; ++a) un[a] = -G[a] / gn;
    for (int a = 0; a < D; ++a) dot += un[a] * ns[a];
    if (dot > 1.0) dot = 1.0; if (dot < -1.0) dot = -1.0;
    const double th = acos(dot), sth = sin(th), cth = cos(th);
    double c1 = 0.0, c2 = 0.0;
    if (fabs(sth) > 1.0e-9) { c1 = st * cth / sth; c2 = st / sth; }
    double n1[3], n2[3], d1 = 0.0, d2 = 0.0;
    for (int a = 0; a < D; ++a) {
        n1[a] = (ct - c1) * ns[a] + c2 * un[a]; n2[a] = (ct + c1) * ns[a] - c2 * un[a];
        d1 += (n1[a] - un[a]) * (n1[a] - un[a]); d2 += (n2[a] - un[a]) * (n2[a] - un[a]);
    }
    d1 = sqrt(d1); d2 = sqrt(d2);
    if (d1 < d2) for (int a = 0; a < D; ++a) G[a] = -gn * n1[a];
    else if (d1 > d2) for (int a = 0; a < D; ++a) G[a] = -gn * n2[a];
}
/* calRKInitialGradient (1582-1632) + wetting + the unit normal used by the force kernel */
static void gradient(cgc* c) {
    const int Q = c->Q, D = c->D; const int64_t N = c->N;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        double G[3] = {0, 0, 0}, nn[3] = {0, 0, 0};
        if (c->dom[id]) {
            for (int i = 1; i < Q; ++i) {
                const double v = c->w[i] * c->phiExt[nbr(c, x, y, z, c->e[i], 1)];
                for (int a = 0; a < D; ++a) if (c->e[i][a]) G[a] += v * c->e[i][a];
            }
            for (int a = 0; a < D; ++a) G[a] *= 3.0;
            if (c->near[id]) {
                const double ns[3] = {c->ns[id], c->ns[N + id], c->ns[2 * N + id]};
                wetting_correction(c, G, ns);
            }
            double gn = 0.0; for (int a = 0; a < D; ++a) gn += G[a] * G[a]; gn = sqrt(gn);
            const int big = c->wetting == 1 ? (gn > 0.0) : (gn > 1.0e-8);
            if (big) for (int a = 0; a < D; ++a) nn[a] = (c->wetting == 1 ? 1.0 : -1.0) * G[a] / gn;
        }
        for (int a = 0; a < 3; ++a) { c->G[a * N + id] = G[a]; c->n[a * N + id] = nn[a]; }
    }
}
/* calForceTermInColorGradient2D (1684-1735) / ...New2D (2497-2552) */
static void csf_force(cgc* c) {
    const int Q = c->Q, D = c->D; const int64_t N = c->N;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) continue;
        double dn[3][3] = {{0}}, n[3] = {c->n[id], c->n[N + id], c->n[2 * N + id]};
        for (int i = 1; i < Q; ++i) {
            const int64_t nb = nbr(c, x, y, z, c->e[i], 1);
            for (int a = 0; a < D; ++a) if (c->e[i][a])
                for (int b = 0; b < D; ++b) dn[a][b] += 3.0 * c->w[i] * c->e[i][a] * c->n[b * N + nb];
        }
        double K = 0.0, nn = 0.0, div = 0.0;
        for (int a = 0; a < D; ++a) {
            nn += n[a] * n[a]; div += dn[a][a];
            for (int b = 0; b < D; ++b) K += n[a] * n[b] * dn[a][b];
        }
        K -= nn * div;
        c->K[id] = K;
        const double sg = c->wetting == 1 ? 0.5 : -0.5;
        for (int a = 0; a < D; ++a) c->F[a * N + id] = sg * c->sigma * K * c->G[a * N + id];
    }
}
static double tau_of(const cgc* c, double phi, double rR, double rB) {
    double tau = 1.0;
    if (phi > c->delta) tau = c->tauR;
    else if (phi < -c->delta) tau = c->tauB;
    else if (fabs(phi) <= c->delta) {
        if (c->tautype == 1) tau = 0.5 + 1.0 / ((1.0 + phi) / (2.0 * (c->tauR - 0.5)) + (1.0 - phi) / (2.0 * (c->tauB - 0.5)));
        else {
            const double xR = rR / (rR + rB), xB = rB / (rR + rB);
            tau = 3.0 * (1.0 / (xR * (3.0 / (c->tauR - 0.5)) + xB * (3.0 / (c->tauB - 0.5)))) + 0.5;
        }
    }
    return tau;
}
/* calRKCollision1TotalGPU2D{SRT,MRT}M (1801-1849, 1934-2018) + calPerturbationFromForce2D[MRT] (1740-1796, 2023-2114) */
static void collide(cgc* c) {
    const int Q = c->Q, D = c->D; const int64_t N = c->N;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) continue;
        double* fT = c->fT + id * Q;
        const double rR = c->rhoR[id], rB = c->rhoB[id];
        const double tau = tau_of(c, c->phi[id], rR, rB);
        double u[3] = {c->u[id], c->u[N + id], c->u[2 * N + id]}, F[3] = {c->F[id], c->F[N + id], c->F[2 * N + id]};
        double uu = 0.0; for (int a = 0; a < D; ++a) uu += u[a] * u[a];
        double fe[QMAX], src[QMAX];
        for (int i = 0; i < Q; ++i) {
            double eu = 0.0; for (int a = 0; a < D; ++a) eu += c->e[i][a] * u[a];
            const double poly = 1.0 + (3.0 * eu + 4.5 * eu * eu - 1.5 * uu);
            fe[i] = rR * c->w[i] * poly + rB * c->w[i] * poly;
            if (c->relax == 0) {
                double t = 0.0;
                for (int a = 0; a < D; ++a) t += (3.0 * (c->e[i][a] - u[a]) + 9.0 * c->e[i][a] * eu) * F[a];
                src[i] = c->w[i] * t * (1.0 - 1.0 / (2.0 * tau));
            } else {
                double t = 0.0;
                for (int a = 0; a < D; ++a) {
                    t += 3.0 * c->e[i][a] * F[a];
                    for (int b = 0; b < D; ++b) t += 9.0 * (c->e[i][a] * c->e[i][b] - (a == b ? 1.0 / 3.0 : 0.0)) * u[a] * F[b];
                }
                src[i] = c->w[i] * t;
            }
        }
        if (c->relax == 0) {
            for (int i = 0; i < Q; ++i) fT[i] = -1.0 / tau * (fT[i] - fe[i]) + fT[i] + src[i];
        } else {
            double m[QMAX], ms[QMAX], S[QMAX];
            for (int k = 0; k < Q; ++k) {
                S[k] = c->Sfix[k] < 0.0 ? 1.0 / tau : c->Sfix[k];
                double a = 0.0, b = 0.0;
                for (int i = 0; i < Q; ++i) { a += c->M[k][i] * (fT[i] - fe[i]); b += c->M[k][i] * src[i]; }
                m[k] = S[k] * a; ms[k] = (1.0 - 0.5 * S[k]) * b;
            }
            for (int i = 0; i < Q; ++i) {
                double a = 0.0, b = 0.0;
                for (int k = 0; k < Q; ++k) { a += c->Mi[i][k] * m[k]; b += c->Mi[i][k] * ms[k]; }
                fT[i] = fT[i] - a + b;
            }
        }
    }
}
/* calRecoloringProcessM (1854-1900) */
static void recolour(cgc* c) {
    const int Q = c->Q, D = c->D; const int64_t N = c->N;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) continue;
        const double rR = c->rhoR[id], rB = c->rhoB[id], tot = rR + rB;
        double gn = 0.0; for (int a = 0; a < D; ++a) gn += c->G[a * N + id] * c->G[a * N + id]; gn = sqrt(gn);
        for (int i = 0; i < Q; ++i) {
            double cost = 0.0;
            if (gn > 1.0e-8 && c->enorm[i] > 1.0e-8) {
                double eg = 0.0; for (int a = 0; a < D; ++a) eg += c->e[i][a] * c->G[a * N + id];
                cost = eg / (c->enorm[i] * gn);
            }
            const double a_ = c->beta * rR * rB / tot * c->w[i] * cost * c->enorm[i];
            c->fR[id * Q + i] = rR / tot * c->fT[id * Q + i] + a_;
            c->fB[id * Q + i] = rB / tot * c->fT[id * Q + i] - a_;
        }
    }
}
/* calStreaming1GPU + calStreaming2GPU (338-417) for both colours, calMacroDensityRKGPU2D (101-118) */
static void stream_density(cgc* c) {
    const int Q = c->Q;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) continue;
        c->fNewR[id * Q] = c->fR[id * Q]; c->fNewB[id * Q] = c->fB[id * Q];
        for (int i = 1; i < Q; ++i) {
            const int64_t s = nbr(c, x, y, z, c->e[i], -1);
            if (c->dom[s]) { c->fNewR[id * Q + i] = c->fR[s * Q + i]; c->fNewB[id * Q + i] = c->fB[s * Q + i]; }
            else { c->fNewR[id * Q + i] = c->fR[id * Q + c->opp[i]]; c->fNewB[id * Q + i] = c->fB[id * Q + c->opp[i]]; }
        }
    }
    double* t = c->fR; c->fR = c->fNewR; c->fNewR = t;
    t = c->fB; c->fB = c->fNewB; c->fNewB = t;
    FOR_NODES(c) {
        const int64_t id = ID(c);
        if (!c->dom[id]) continue;
        double r = c->fR[id * Q], b = c->fB[id * Q];
        for (int i = 1; i < Q; ++i) { r += c->fR[id * Q + i]; b += c->fB[id * Q + i]; }
        c->rhoR[id] = r; c->rhoB[id] = b;
    }
}

void cgc_head(cgc* c) { total_pdf(c); velocity_phi(c); }
void cgc_body(cgc* c) { phi_on_solid(c); gradient(c); csf_force(c); collide(c); recolour(c); stream_density(c); }
void cgc_step(cgc* c, int n) {
    for (int s = 0; s < n; ++s) { cgc_head(c); cgc_body(c); }
}
/* dense copies out: rho [N], u [3][N], pdf [N][Q] */
void cgc_get(cgc* c, double* rhoR, double* rhoB, double* u, double* fR, double* fB) {
    if (rhoR) memcpy(rhoR, c->rhoR, c->N * 8);
    if (rhoB) memcpy(rhoB, c->rhoB, c->N * 8);
    if (u) memcpy(u, c->u, 3 * c->N * 8);
    if (fR) memcpy(fR, c->fR, (size_t)c->N * c->Q * 8);
    if (fB) memcpy(fB, c->fB, (size_t)c->N * c->Q * 8);
}
