// lattice.cuh -- lattice constants and per-node arithmetic of the colour-gradient step,
// shared by every kernel.  Everything is `__host__ __device__` so that the node math can be
// compiled for the host by tests/hostcheck (a test hook, never a product path).
//
// Conventions follow the reference: D2Q9 velocity order and MRT basis of
// RKCG2D/RKD2Q9.py:299-340; D3Q19 in d'Humieres' order (the reference ships no 3-D code).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LBM_HD __host__ __device__ __forceinline__
#else
#define LBM_HD inline
#endif

namespace lbm {

// Products inside the gradient / curvature stencils must be rounded before they are summed: the reference
// relies on EXACT cancellation of w_k phi(x+e_k) against w_k phi(x-e_k) in uniform regions (its type-1
// force kernel turns any non-zero |G|, even 1e-17 of rounding noise, into a unit normal).  An FMA would
// keep the product unrounded and break that cancellation.
LBM_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#elif defined(LBM_HOST_FMA_TEST)     // host build with -ffp-contract=fast (debug aid): block the contraction
    volatile double r = a * b;
    return r;
#else
    return a * b;
#endif
}
LBM_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

struct D2Q9 {
    static constexpr int Q = 9;
    static constexpr int D = 2;
    static constexpr int NMOM = 9;
    LBM_HD static constexpr int cx(int i) { constexpr int t[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1}; return t[i]; }
    LBM_HD static constexpr int cz(int) { return 0; }
    // offsets along the array axes (0 = fastest, 2 = slab axis): the 2-D lattice lives on axes 0 and 2
    LBM_HD static constexpr int d0(int i) { return cx(i); }
    LBM_HD static constexpr int d1(int) { return 0; }
    LBM_HD static constexpr int d2(int i) { return cy(i); }
    LBM_HD static constexpr int c(int i, int a) { return a == 0 ? cx(i) : (a == 1 ? cy(i) : 0); }
    LBM_HD static constexpr int opp(int i) { constexpr int t[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6}; return t[i]; }
    LBM_HD static constexpr double w(int i) { return i == 0 ? 4.0 / 9.0 : (i < 5 ? 1.0 / 9.0 : 1.0 / 36.0); }
    LBM_HD static constexpr double enorm(int i) { return i == 0 ? 0.0 : (i < 5 ? 1.0 : 1.4142135623730951); }

    // m = M f, rows (rho, e, eps, jx, qx, jy, qy, pxx, pxy)            RKD2Q9.py:309-336
    LBM_HD static void to_moments(const double* f, double* m) {
        const double s13 = f[1] + f[3], s24 = f[2] + f[4];
        const double sa = s13 + s24, sd = (f[5] + f[6]) + (f[7] + f[8]);
        const double dx = f[1] - f[3], dy = f[2] - f[4];
        const double ddx = (f[5] - f[6]) + (f[8] - f[7]);
        const double ddy = (f[5] + f[6]) - (f[7] + f[8]);
        m[0] = f[0] + sa + sd;
        m[1] = -4.0 * f[0] - sa + 2.0 * sd;
        m[2] = 4.0 * f[0] - 2.0 * sa + sd;
        m[3] = dx + ddx;
        m[4] = -2.0 * dx + ddx;
        m[5] = dy + ddy;
        m[6] = -2.0 * dy + ddy;
        m[7] = s13 - s24;
        m[8] = (f[5] - f[6]) + (f[7] - f[8]);
    }
    // f = M^-1 m = M^T diag(1/|row|^2) m
    LBM_HD static void from_moments(const double* m, double* f) {
        const double a0 = m[0] * (1.0 / 9.0), a1 = m[1] * (1.0 / 36.0), a2 = m[2] * (1.0 / 36.0);
        const double a3 = m[3] * (1.0 / 6.0), a4 = m[4] * (1.0 / 12.0);
        const double a5 = m[5] * (1.0 / 6.0), a6 = m[6] * (1.0 / 12.0);
        const double a7 = m[7] * 0.25, a8 = m[8] * 0.25;
        const double ba = a0 - a1 - 2.0 * a2, bd = a0 + 2.0 * a1 + a2;
        const double ox = a3 - 2.0 * a4, oy = a5 - 2.0 * a6;
        const double px = a3 + a4, py = a5 + a6;
        f[0] = a0 - 4.0 * a1 + 4.0 * a2;
        f[1] = ba + ox + a7;
        f[3] = ba - ox + a7;
        f[2] = ba + oy - a7;
        f[4] = ba - oy - a7;
        f[5] = bd + px + py + a8;
        f[6] = bd - px + py - a8;
        f[7] = bd - px - py + a8;
        f[8] = bd + px - py - a8;
    }
    // in place: m <- m - S (m - meq(rho,u)) + (1 - S/2) (M s)(u,F);  s_nu = 1/tau
    LBM_HD static void relax_moments(double* m, double rho, const double* u, const double* F, double s_nu) {
        const double ux = u[0], uy = u[1], Fx = F[0], Fy = F[1];
        const double uu = ux * ux + uy * uy, uF = ux * Fx + uy * Fy;
        constexpr double s1 = 1.64, s2 = 1.54, s4 = 1.9;            // RKD2Q9.py:338-340
        m[1] += -s1 * (m[1] - rho * (-2.0 + 3.0 * uu)) + (1.0 - 0.5 * s1) * (6.0 * uF);
        m[2] += -s2 * (m[2] - rho * (1.0 - 3.0 * uu)) + (1.0 - 0.5 * s2) * (-6.0 * uF);
        m[3] += Fx;
        m[4] += -s4 * (m[4] + rho * ux) + (1.0 - 0.5 * s4) * (-Fx);
        m[5] += Fy;
        m[6] += -s4 * (m[6] + rho * uy) + (1.0 - 0.5 * s4) * (-Fy);
        m[7] += -s_nu * (m[7] - rho * (ux * ux - uy * uy)) + (1.0 - 0.5 * s_nu) * (2.0 * (ux * Fx - uy * Fy));
        m[8] += -s_nu * (m[8] - rho * ux * uy) + (1.0 - 0.5 * s_nu) * (ux * Fy + uy * Fx);
    }
};

struct D3Q19 {
    static constexpr int Q = 19;
    static constexpr int D = 3;
    static constexpr int NMOM = 19;
    LBM_HD static constexpr int cx(int i) {
        constexpr int t[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
        return t[i];
    }
    LBM_HD static constexpr int cy(int i) {
        constexpr int t[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
        return t[i];
    }
    LBM_HD static constexpr int cz(int i) {
        constexpr int t[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
        return t[i];
    }
    LBM_HD static constexpr int d0(int i) { return cx(i); }
    LBM_HD static constexpr int d1(int i) { return cy(i); }
    LBM_HD static constexpr int d2(int i) { return cz(i); }
    LBM_HD static constexpr int c(int i, int a) { return a == 0 ? cx(i) : (a == 1 ? cy(i) : cz(i)); }
    LBM_HD static constexpr int opp(int i) {
        constexpr int t[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
        return t[i];
    }
    LBM_HD static constexpr double w(int i) { return i == 0 ? 1.0 / 3.0 : (i < 7 ? 1.0 / 18.0 : 1.0 / 36.0); }
    LBM_HD static constexpr double enorm(int i) { return i == 0 ? 0.0 : (i < 7 ? 1.0 : 1.4142135623730951); }

    // d'Humieres et al. 2002: (rho, e, eps, jx, qx, jy, qy, jz, qz, 3pxx, 3pixx, pww, piww, pxy, pyz, pxz, mx, my, mz)
    LBM_HD static void to_moments(const double* f, double* m) {
        const double sx = f[1] + f[2], sy = f[3] + f[4], sz = f[5] + f[6];
        const double a = f[7] + f[10], b = f[8] + f[9];        // (1,1,0)/(-1,-1,0) ; (-1,1,0)/(1,-1,0)
        const double c = f[11] + f[14], d = f[12] + f[13];     // (1,0,1)/(-1,0,-1) ; (-1,0,1)/(1,0,-1)
        const double e = f[15] + f[18], g = f[16] + f[17];     // (0,1,1)/(0,-1,-1) ; (0,-1,1)/(0,1,-1)
        const double xy = a + b, xz = c + d, yz = e + g;
        const double A = sx + sy + sz, Dg = xy + xz + yz;
        m[0] = f[0] + A + Dg;
        m[1] = -30.0 * f[0] - 11.0 * A + 8.0 * Dg;
        m[2] = 12.0 * f[0] - 4.0 * A + Dg;
        const double p1 = 2.0 * sx - sy - sz, p2 = xy + xz - 2.0 * yz;
        m[9] = p1 + p2;
        m[10] = -2.0 * p1 + p2;
        const double w1 = sy - sz, w2 = xy - xz;
        m[11] = w1 + w2;
        m[12] = -2.0 * w1 + w2;
        m[13] = a - b;
        m[14] = e - g;
        m[15] = c - d;
        const double dx = f[1] - f[2], dy = f[3] - f[4], dz = f[5] - f[6];
        const double da = f[7] - f[10], db = f[8] - f[9];
        const double dc = f[11] - f[14], dd = f[12] - f[13];
        const double de = f[15] - f[18], dg = f[16] - f[17];
        const double xA = da - db, xB = dc - dd;               // cx-weighted diagonal sums (xy, xz planes)
        const double yA = da + db, yB = de - dg;               // cy-weighted (xy, yz planes)
        const double zA = dc + dd, zB = de + dg;               // cz-weighted (xz, yz planes)
        m[3] = dx + xA + xB;
        m[4] = -4.0 * dx + xA + xB;
        m[5] = dy + yA + yB;
        m[6] = -4.0 * dy + yA + yB;
        m[7] = dz + zA + zB;
        m[8] = -4.0 * dz + zA + zB;
        m[16] = xA - xB;
        m[17] = yB - yA;
        m[18] = zA - zB;
    }
    LBM_HD static void from_moments(const double* m, double* f) {
        const double a0 = m[0] * (1.0 / 19.0), a1 = m[1] * (1.0 / 2394.0), a2 = m[2] * (1.0 / 252.0);
        const double a3 = m[3] * 0.1, a4 = m[4] * 0.025, a5 = m[5] * 0.1, a6 = m[6] * 0.025;
        const double a7 = m[7] * 0.1, a8 = m[8] * 0.025;
        const double a9 = m[9] * (1.0 / 36.0), a10 = m[10] * (1.0 / 72.0);
        const double a11 = m[11] * (1.0 / 12.0), a12 = m[12] * (1.0 / 24.0);
        const double a13 = m[13] * 0.25, a14 = m[14] * 0.25, a15 = m[15] * 0.25;
        const double a16 = m[16] * 0.125, a17 = m[17] * 0.125, a18 = m[18] * 0.125;
        f[0] = a0 - 30.0 * a1 + 12.0 * a2;
        const double ba = a0 - 11.0 * a1 - 4.0 * a2;
        const double ex_ = 2.0 * (a9 - 2.0 * a10);             // even part, x axis: 2 a9 - 4 a10
        const double eyz = -(a9 - 2.0 * a10);                  // even part shared by y and z axes
        const double ew = a11 - 2.0 * a12;
        const double ox = a3 - 4.0 * a4, oy = a5 - 4.0 * a6, oz = a7 - 4.0 * a8;
        f[1] = ba + ex_ + ox;
        f[2] = ba + ex_ - ox;
        f[3] = ba + eyz + ew + oy;
        f[4] = ba + eyz + ew - oy;
        f[5] = ba + eyz - ew + oz;
        f[6] = ba + eyz - ew - oz;
        const double bd = a0 + 8.0 * a1 + a2;
        const double p = a9 + a10, q = a11 + a12;
        const double jx = a3 + a4, jy = a5 + a6, jz = a7 + a8;
        // xy plane: 7 (1,1,0), 8 (-1,1,0), 9 (1,-1,0), 10 (-1,-1,0)
        const double bxy = bd + p + q;
        const double x1 = jx + a16, y1 = jy - a17;
        f[7] = bxy + a13 + x1 + y1;
        f[10] = bxy + a13 - x1 - y1;
        f[8] = bxy - a13 - x1 + y1;
        f[9] = bxy - a13 + x1 - y1;
        // xz plane: 11 (1,0,1), 12 (-1,0,1), 13 (1,0,-1), 14 (-1,0,-1)
        const double bxz = bd + p - q;
        const double x2 = jx - a16, z2 = jz + a18;
        f[11] = bxz + a15 + x2 + z2;
        f[14] = bxz + a15 - x2 - z2;
        f[12] = bxz - a15 - x2 + z2;
        f[13] = bxz - a15 + x2 - z2;
        // yz plane: 15 (0,1,1), 16 (0,-1,1), 17 (0,1,-1), 18 (0,-1,-1)
        const double byz = bd - 2.0 * p;
        const double y3 = jy + a17, z3 = jz - a18;
        f[15] = byz + a14 + y3 + z3;
        f[18] = byz + a14 - y3 - z3;
        f[16] = byz - a14 - y3 + z3;
        f[17] = byz - a14 + y3 - z3;
    }
    LBM_HD static void relax_moments(double* m, double rho, const double* u, const double* F, double s_nu) {
        const double ux = u[0], uy = u[1], uz = u[2], Fx = F[0], Fy = F[1], Fz = F[2];
        const double uu = ux * ux + uy * uy + uz * uz, uF = ux * Fx + uy * Fy + uz * Fz;
        constexpr double s_e = 1.19, s_eps = 1.4, s_q = 1.2, s_pi = 1.4, s_m = 1.98;   // DESIGN.md, D3Q19 spec
        constexpr double t3 = 2.0 / 3.0;
        m[1] += -s_e * (m[1] - rho * (-11.0 + 19.0 * uu)) + (1.0 - 0.5 * s_e) * (38.0 * uF);
        m[2] += -s_eps * (m[2] - rho * (3.0 - 5.5 * uu)) + (1.0 - 0.5 * s_eps) * (-11.0 * uF);
        m[3] += Fx;
        m[4] += -s_q * (m[4] + t3 * rho * ux) + (1.0 - 0.5 * s_q) * (-t3 * Fx);
        m[5] += Fy;
        m[6] += -s_q * (m[6] + t3 * rho * uy) + (1.0 - 0.5 * s_q) * (-t3 * Fy);
        m[7] += Fz;
        m[8] += -s_q * (m[8] + t3 * rho * uz) + (1.0 - 0.5 * s_q) * (-t3 * Fz);
        const double pe = rho * (2.0 * ux * ux - uy * uy - uz * uz), pf = 2.0 * (2.0 * ux * Fx - uy * Fy - uz * Fz);
        m[9] += -s_nu * (m[9] - pe) + (1.0 - 0.5 * s_nu) * pf;
        m[10] += -s_pi * (m[10] + 0.5 * pe) + (1.0 - 0.5 * s_pi) * (-0.5 * pf);
        const double we = rho * (uy * uy - uz * uz), wf = 2.0 * (uy * Fy - uz * Fz);
        m[11] += -s_nu * (m[11] - we) + (1.0 - 0.5 * s_nu) * wf;
        m[12] += -s_pi * (m[12] + 0.5 * we) + (1.0 - 0.5 * s_pi) * (-0.5 * wf);
        m[13] += -s_nu * (m[13] - rho * ux * uy) + (1.0 - 0.5 * s_nu) * (ux * Fy + uy * Fx);
        m[14] += -s_nu * (m[14] - rho * uy * uz) + (1.0 - 0.5 * s_nu) * (uy * Fz + uz * Fy);
        m[15] += -s_nu * (m[15] - rho * ux * uz) + (1.0 - 0.5 * s_nu) * (ux * Fz + uz * Fx);
        m[16] += -s_m * m[16];
        m[17] += -s_m * m[17];
        m[18] += -s_m * m[18];
    }
};

// Model constants of one run (kernel argument, lives in constant bank)
struct CGParams {
    double sigma, cosT, sinT, beta, delta, tauR, tauB;
    int tau_type, wetting_type, relax;
    int st_type;        // LBM_ST_*: 0 = continuum surface force, 1 = perturbation operator
    double Ak, solid_phi, bf[3];   // perturbation operator: (AkR + AkB) / 2, phi on solid neighbours, body force
    int exact_trig;     // tiled 3-D kernels: 1 = the reference-ordered wetting arithmetic (cg_wetting), 0 = cg_wetting_akai3_fast
};

// tau(phi) -- AcceleratedRKGPU2D.py:1820-1834 (identical in the four collision/forcing kernels)
LBM_HD double cg_tau(double phi, double rhoR, double rhoB, const CGParams& p) {
    double tau = 1.0;
    if (phi > p.delta) tau = p.tauR;
    else if (phi < -p.delta) tau = p.tauB;
    else if (fabs(phi) <= p.delta) {
        if (p.tau_type == 1) {
            tau = 0.5 + 1.0 / ((1.0 + phi) / (2.0 * (p.tauR - 0.5)) + (1.0 - phi) / (2.0 * (p.tauB - 0.5)));
        } else if (p.tau_type == 2) {
            const double inv = 1.0 / (rhoR + rhoB);
            const double miu = 1.0 / (rhoR * inv * (3.0 / (p.tauR - 0.5)) + rhoB * inv * (3.0 / (p.tauB - 0.5)));
            tau = 3.0 * miu + 0.5;
        }
    }
    return tau;
}

// Collision of the total population incl. forcing, in place.
//   SRT: calRKCollision1TotalGPU2DSRTM (1801-1849) + calPerturbationFromForce2D (1740-1796)
//   MRT: calRKCollision1TotalGPU2DMRTM (1934-2018) + calPerturbationFromForce2DMRT (2023-2114),
//        done once in moment space with the analytic M*feq and M*s (DESIGN.md "moment space").
template <class L>
LBM_HD void cg_collide(double* fT, double rho, const double* u, const double* F, double tau, int relax) {
    if (relax == 0) {
        const double uu = u[0] * u[0] + u[1] * u[1] + (L::D == 3 ? u[2] * u[2] : 0.0);
        const double it = 1.0 / tau, fc = 1.0 - 1.0 / (2.0 * tau);
#pragma unroll
        for (int i = 0; i < L::Q; ++i) {
            const double eu = L::cx(i) * u[0] + L::cy(i) * u[1] + (L::D == 3 ? L::cz(i) * u[2] : 0.0);
            const double feq = rho * L::w(i) * (1.0 + (3.0 * eu + 4.5 * eu * eu - 1.5 * uu));
            double src = (3.0 * (L::cx(i) - u[0]) + 9.0 * L::cx(i) * eu) * F[0] +
                         (3.0 * (L::cy(i) - u[1]) + 9.0 * L::cy(i) * eu) * F[1];
            if (L::D == 3) src += (3.0 * (L::cz(i) - u[2]) + 9.0 * L::cz(i) * eu) * F[2];
            fT[i] = fT[i] - it * (fT[i] - feq) + L::w(i) * src * fc;
        }
    } else {
        double m[L::NMOM];
        L::to_moments(fT, m);
        L::relax_moments(m, rho, u, F, 1.0 / tau);
        L::from_moments(m, fT);
    }
}

// Recolouring, calRecoloringProcessM (1854-1900): fR_i = rhoR/rho fT_i + beta rhoR rhoB/rho w_i cos(theta_i)|e_i|
template <class L>
LBM_HD void cg_recolour(const double* fT, double rhoR, double rhoB, const double* G, double beta,
                        double* fR, double* fB) {
    const double gn = sqrt(G[0] * G[0] + G[1] * G[1] + (L::D == 3 ? G[2] * G[2] : 0.0));
    const double inv = 1.0 / (rhoR + rhoB);
    const double kR = rhoR * inv, kB = rhoB * inv;
    const double ig = gn > 1.0e-8 ? 1.0 / gn : 0.0;
    const double amp = beta * rhoR * rhoB * inv * ig;
#pragma unroll
    for (int i = 0; i < L::Q; ++i) {
        // cos(theta_i) |e_i| = e_i . G / |G|   (both norms > 1e-8, else 0)
        const double eg = L::cx(i) * G[0] + L::cy(i) * G[1] + (L::D == 3 ? L::cz(i) * G[2] : 0.0);
        const double a = i == 0 ? 0.0 : amp * L::w(i) * eg;
        fR[i] = kR * fT[i] + a;
        fB[i] = kB * fT[i] - a;
    }
}

// Contact-angle correction of the colour gradient on a fluid node next to solid.
//   type 1: updateColorGradientOnWetting    (1637-1679, Xu et al. 2017; 2-D rotation)
//   type 2: updateColorGradientOnWettingNew (2428-2492, Akai et al. 2018; any D)
template <int D>
LBM_HD void cg_wetting(double* G, const double* ns, double cosT, double sinT, int type) {
    // Every product and sum below is rounded separately (mul_rn / add_rn, no FMA contraction): which of the
    // two candidate normals is "closer" is decided by comparing two distances that are EQUAL by symmetry at
    // corner and axis nodes; the reference's tie rule (d1 == d2) only reproduces with its own rounding.
    auto M = [](double a, double b) { return mul_rn(a, b); };
    auto A = [](double a, double b) { return add_rn(a, b); };
    auto S = [](double a, double b) { return add_rn(a, -b); };
    const double gn = sqrt(D == 3 ? A(A(M(G[0], G[0]), M(G[1], G[1])), M(G[2], G[2])) : A(M(G[0], G[0]), M(G[1], G[1])));
    if (type == 1) {
        const double n1x = S(M(ns[0], cosT), M(ns[1], sinT)), n1y = A(M(ns[1], cosT), M(ns[0], sinT));
        const double n2x = A(M(ns[0], cosT), M(ns[1], sinT)), n2y = S(M(ns[1], cosT), M(ns[0], sinT));
        double ux = 0.0, uy = 0.0;
        if (gn > 1.0e-8) { ux = G[0] / gn; uy = G[1] / gn; }
        const double d1 = sqrt(A(M(S(ux, n1x), S(ux, n1x)), M(S(uy, n1y), S(uy, n1y))));
        const double d2 = sqrt(A(M(S(ux, n2x), S(ux, n2x)), M(S(uy, n2y), S(uy, n2y))));
        double mx = 0.0, my = 0.0;
        if (d1 < d2) { mx = n1x; my = n1y; }
        else if (d1 > d2) { mx = n2x; my = n2y; }
        else if (d1 == d2) { mx = ns[0]; my = ns[1]; }
        G[0] = M(gn, mx); G[1] = M(gn, my);
        return;
    }
    double un[3] = {0.0, 0.0, 0.0};
    if (gn > 1.0e-8) {
        un[0] = -G[0] / gn; un[1] = -G[1] / gn;
        if (D == 3) un[2] = -G[2] / gn;
    }
    double dot = A(M(un[0], ns[0]), M(un[1], ns[1]));
    if (D == 3) dot = A(dot, M(un[2], ns[2]));
    // acos outside [-1,1] is NaN on the GPU and ends in "no update" (2451-2460); clamping gives
    // sin(theta') = 0 (or 1.2e-16) and therefore the same outcome without the NaN.
    dot = fmin(1.0, fmax(-1.0, dot));
    const double th = acos(dot);
    const double sth = sin(th), cth = cos(th);
    double c1 = 0.0, c2 = 0.0;
    if (fabs(sth) > 1.0e-9) { c1 = M(sinT, cth) / sth; c2 = sinT / sth; }
    double d1 = 0.0, d2 = 0.0, n1[3], n2[3];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        n1[a] = A(M(S(cosT, c1), ns[a]), M(c2, un[a]));
        n2[a] = S(M(A(cosT, c1), ns[a]), M(c2, un[a]));
        d1 = A(d1, M(S(n1[a], un[a]), S(n1[a], un[a])));
        d2 = A(d2, M(S(n2[a], un[a]), S(n2[a], un[a])));
    }
    d1 = sqrt(d1); d2 = sqrt(d2);
    if (d1 < d2) {
#pragma unroll
        for (int a = 0; a < D; ++a) G[a] = M(-gn, n1[a]);
    } else if (d1 > d2) {
#pragma unroll
        for (int a = 0; a < D; ++a) G[a] = M(-gn, n2[a]);
    }
}

// The same Akai-2018 correction for the tiled 3-D kernels, whose parity target is the oracle at 1e-10 (no golden
// vector hangs on its last bit): one rsqrt instead of three divisions, cos(theta') = dot and sin(theta') =
// sqrt(1 - dot^2) instead of acos / sin / cos, one reciprocal, and the two candidate distances compared squared.
// ~80 FP64 instructions instead of ~450 on the critical path of every tile that touches a solid surface.  The
// guards are those of cg_wetting: |G| <= 1e-8 gives u = 0 and two identical candidates (no update), sin(theta')
// <= 1e-9 gives c1 = c2 = 0 and two identical candidates (no update).
LBM_HD void cg_wetting_akai3_fast(double* G, const double* ns, double cosT, double sinT) {
    const double g2 = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
    if (!(g2 > 1.0e-16)) return;
#ifdef __CUDA_ARCH__
    const double inv = rsqrt(g2);
#else
    const double inv = 1.0 / sqrt(g2);
#endif
    const double gn = g2 * inv;
    const double un[3] = {-G[0] * inv, -G[1] * inv, -G[2] * inv};
    double dot = un[0] * ns[0] + un[1] * ns[1] + un[2] * ns[2];
    dot = fmin(1.0, fmax(-1.0, dot));
    const double sth = sqrt(fmax(0.0, 1.0 - dot * dot));
    if (!(sth > 1.0e-9)) return;
    const double r = 1.0 / sth, c1 = sinT * dot * r, c2 = sinT * r;
    double d1 = 0.0, d2 = 0.0, n1[3], n2[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        n1[a] = (cosT - c1) * ns[a] + c2 * un[a];
        n2[a] = (cosT + c1) * ns[a] - c2 * un[a];
        d1 += (n1[a] - un[a]) * (n1[a] - un[a]);
        d2 += (n2[a] - un[a]) * (n2[a] - un[a]);
    }
    if (d1 < d2) {
#pragma unroll
        for (int a = 0; a < 3; ++a) G[a] = -gn * n1[a];
    } else if (d1 > d2) {
#pragma unroll
        for (int a = 0; a < 3; ++a) G[a] = -gn * n2[a];
    }
}

// Interface normal used by the curvature: type 1: +G/|G| (|G| > 0); type 2: -G/|G| (|G| > 1e-8), else 0.
template <int D>
LBM_HD void cg_unit_normal(const double* G, int wetting_type, double* n) {
    const double gn = sqrt(G[0] * G[0] + G[1] * G[1] + (D == 3 ? G[2] * G[2] : 0.0));
    const bool big = wetting_type == 1 ? (gn > 0.0) : (gn > 1.0e-8);
    const double s = big ? (wetting_type == 1 ? 1.0 : -1.0) / gn : 0.0;
    n[0] = s * G[0]; n[1] = s * G[1];
    if (D == 3) n[2] = s * G[2];
}

}  // namespace lbm
