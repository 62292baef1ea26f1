// tr_ops.cuh -- solute tracers riding on the colour-gradient CSF flow, one thread per node,
// any lattice.  The tracer part of the reference's runTransport2DMPMCRKNew with NumberSchemes = 9
// (RKCG2D/Transport2DRK.py:1341-1425; kernels in RKCG2D/AccelerateTransport2DRK.py):
//   calValueTransportDomain (957-971)   indicator = -1 where rho_R <= criterion, else 0
//   calCollisionQ9 (704-730)            SRT towards the linear equilibrium C w_j (1 + 3 e_j.u)
//   calCollisionTransportLinearEqlMRTGPUD2Q9 (1053-1105) with the constants of Transport2DRK.py:367-391:
//                                       g <- g - M^-1 S^-1 M (g - g_eq); S = 1 except the flux moments
//   calTransportWithInterfaceD2Q9 (1019-1047)   g_j += beta indicator w_j C cos(angle(e_j, -G)),  j > 0, |G| > 1e-8
//   calStreaming1GPU / calStreaming2GPU (736-835), calConcentrationGPU (78-90)
// The reference runs these as five kernels over [tracer, node, 9] AoS arrays; here collision + interface term are one
// operator, streaming (pull form, half-way bounce back) + concentration the other, on SoA arrays [tracer][Q][vol].
// Next to a wetting solid the reference's streaming tests `neighbour != -1` on a table whose wetting solids are <= -2 and
// writes through a negative index; here the population bounces back.
#pragma once
#include "cg_ops.cuh"

namespace lbm {

constexpr int TR_MAX = 4;

struct TracerParams {
    int nt, relax;
    double tau[TR_MAX], beta[TR_MAX];
    double sa[TR_MAX], sb[TR_MAX], sc[TR_MAX], sd[TR_MAX];   // S block of the flux moments: [[a, b], [c, d]] on (j_x, j_y); a also on q_x, d on q_y
    double criterion;
    // 5-velocity branch (NumberSchemes = 5)
    int schemes;                 // 9 (the flow lattice) | 5
    int reaction;                // A + B -> C on tracers 0, 1, 2
    double rate, j0[TR_MAX];     // reaction rate; J_0 of each tracer (rest share of the source; (1 - J_0)/4 on the others)
    int inlet_row, outlet_row;   // local plane of the Inamuro inlet / of the free-flow outlet on this slab, -1: none here
    double inlet_conc[TR_MAX];
};

// The reference's 5-velocity tracer lattice (Transport2DRK.py:60-61, 313-322): rest, +x, -x, +y, -y; the 2-D lattice lives on
// array axes 0 and 2 like D2Q9.
struct D2Q5 {
    static constexpr int Q = 5;
    static constexpr int D = 2;
    LBM_HD static constexpr int cx(int i) { constexpr int t[5] = {0, 1, -1, 0, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[5] = {0, 0, 0, 1, -1}; return t[i]; }
    LBM_HD static constexpr int d0(int i) { return cx(i); }
    LBM_HD static constexpr int d1(int) { return 0; }
    LBM_HD static constexpr int d2(int i) { return cy(i); }
    LBM_HD static constexpr int c(int i, int a) { return a == 0 ? cx(i) : (a == 1 ? cy(i) : 0); }
    LBM_HD static constexpr int opp(int i) { constexpr int t[5] = {0, 2, 1, 4, 3}; return t[i]; }
    LBM_HD static constexpr double w(int i) { return i == 0 ? 1.0 / 3.0 : 1.0 / 6.0; }
};

struct TracerFields {
    TracerParams p;
    double* g;      // [nt][Q][vol] streamed tracer populations
    double* gC;     // [nt][Q][vol] post-collision
    double* conc;   // [nt][vol]
};

// collision + interface term: g -> gC
template <class L>
struct TracerCollideOp {
    CGFields c; TracerFields t;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        const int64_t id = (int64_t)NG * g.plane + i, V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double u[3] = {0.0, 0.0, 0.0}, ug[3] = {0.0, 0.0, 0.0};
        double g2 = 0.0;
#pragma unroll
        for (int a = 0; a < L::D; ++a) { u[a] = c.u[a * V + id]; const double G = c.G[a * V + id]; ug[a] = G; g2 += G * G; }
        const double gn = sqrt(g2);
        double un = 0.0;
        if (gn > 1.0e-8) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a) { ug[a] = -ug[a] / gn; s += ug[a] * ug[a]; }
            un = sqrt(s);
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) ug[a] = 0.0;
        }
        const double value = c.rho[0][id] > t.p.criterion ? -0.0 : -1.0;
        for (int k = 0; k < t.p.nt; ++k) {
            const double C = t.conc[(int64_t)k * V + id];
            double f[L::Q], d[L::Q];
#pragma unroll
            for (int q = 0; q < L::Q; ++q) {
                f[q] = t.g[((int64_t)k * L::Q + q) * V + id];
                double eu = 0.0;
#pragma unroll
                for (int a = 0; a < L::D; ++a)
                    if (L::c(q, a) != 0) eu += L::c(q, a) * u[a];
                d[q] = f[q] - C * L::w(q) * (1.0 + 3.0 * eu);
            }
            if (t.p.relax == 0) {
#pragma unroll
                for (int q = 0; q < L::Q; ++q) f[q] = -d[q] / t.p.tau[k] + f[q];
            } else {          // D2Q9 only (checked by lbm_tracer_setup): moments (rho, e, eps, jx, qx, jy, qy, pxx, pxy)
                double m[L::NMOM], back[L::Q];
                L::to_moments(d, m);
                const double a = t.p.sa[k], b = t.p.sb[k], cc = t.p.sc[k], dd = t.p.sd[k];
                const double det = a * dd - b * cc;
                const double m3 = (dd * m[3] - b * m[5]) / det, m5 = (a * m[5] - cc * m[3]) / det;
                m[3] = m3; m[5] = m5; m[4] = m[4] / a; m[6] = m[6] / dd;
                L::from_moments(m, back);
#pragma unroll
                for (int q = 0; q < L::Q; ++q) f[q] = f[q] - back[q];
            }
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                double eg = 0.0;
#pragma unroll
                for (int a = 0; a < L::D; ++a)
                    if (L::c(q, a) != 0) eg += L::c(q, a) * ug[a];
                const double cosT = un > 1.0e-8 ? eg / (L::enorm(q) * un) : 0.0;
                f[q] = f[q] + t.p.beta[k] * value * (L::w(q) * C) * cosT;
            }
#pragma unroll
            for (int q = 0; q < L::Q; ++q) t.gC[((int64_t)k * L::Q + q) * V + id] = f[q];
        }
    }
};

// 5-velocity branch: MRT collision (calCollisionTransportLinearEqlMRTGPU :535-590 with M, S of Transport2DRK.py:313-347),
// interface term (calTransportWithInterfaceD2Q5 :976-1013), reaction source (calReactionTracersGPU :95-112): g -> gC.
// M rows: (1,1,1,1,1), j_x = (0,1,-1,0,0), j_y = (0,0,0,1,-1), (4,-1,-1,-1,-1), (0,1,1,-1,-1); squared norms 5, 2, 2, 20, 4.
struct TracerCollideQ5Op {
    CGFields c; TracerFields t;
    LBM_HD void operator()(int64_t i) const {
        using L = D2Q5;
        const Grid& g = c.g;
        const int64_t id = (int64_t)NG * g.plane + i, V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const double u[2] = {c.u[id], c.u[V + id]};
        double ug[2] = {c.G[id], c.G[V + id]};
        const double gn = sqrt(ug[0] * ug[0] + ug[1] * ug[1]);
        double un = 0.0;
        if (gn > 1.0e-8) {
            ug[0] = -ug[0] / gn; ug[1] = -ug[1] / gn;
            un = sqrt(ug[0] * ug[0] + ug[1] * ug[1]);
        } else {
            ug[0] = 0.0; ug[1] = 0.0;
        }
        const double value = c.rho[0][id] > t.p.criterion ? -0.0 : -1.0;
        double src = 0.0;
        if (t.p.reaction) src = t.p.rate * t.conc[id] * t.conc[V + id];
        for (int k = 0; k < t.p.nt; ++k) {
            const double C = t.conc[(int64_t)k * V + id];
            double f[5], d[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                f[q] = t.g[((int64_t)k * 5 + q) * V + id];
                d[q] = f[q] - C * L::w(q) * (1.0 + 3.0 * (L::cx(q) * u[0] + L::cy(q) * u[1]));
            }
            double m0 = (((d[0] + d[1]) + d[2]) + d[3]) + d[4];
            const double mx = d[1] - d[2], my = d[3] - d[4];
            double m3 = 4.0 * d[0] - (((d[1] + d[2]) + d[3]) + d[4]);
            double m4 = (d[1] + d[2]) - (d[3] + d[4]);
            const double a = t.p.sa[k], b = t.p.sb[k], cc = t.p.sc[k], dd = t.p.sd[k];
            const double det = a * dd - b * cc;
            const double rx = (dd * mx - b * my) / det * 0.5, ry = (a * my - cc * mx) / det * 0.5;
            m0 *= 0.2; m3 *= 0.05; m4 *= 0.25;
            f[0] -= m0 + 4.0 * m3;
            f[1] -= m0 + rx - m3 + m4;
            f[2] -= m0 - rx - m3 + m4;
            f[3] -= m0 + ry - m3 - m4;
            f[4] -= m0 - ry - m3 - m4;
#pragma unroll
            for (int q = 1; q < 5; ++q) {
                const double eg = L::cx(q) * ug[0] + L::cy(q) * ug[1];
                const double cosT = un > 1.0e-8 ? eg / un : 0.0;
                f[q] = f[q] + t.p.beta[k] * value * (L::w(q) * C) * cosT;
            }
            if (t.p.reaction) {
                const double sk = k == 2 ? src : -src, jr = (1.0 - t.p.j0[k]) / 4.0;
                f[0] = f[0] + t.p.j0[k] * sk;
#pragma unroll
                for (int q = 1; q < 5; ++q) f[q] = f[q] + jr * sk;
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) t.gC[((int64_t)k * 5 + q) * V + id] = f[q];
        }
    }
};

// free-flow outlet (calFreeConcBoundary3 :461-474): the outlet row takes the post-collision populations of the row above it
// (where that node is solid the reference reads through index -1; here the node keeps its own)
struct TracerFreeflowOp {
    CGFields c; TracerFields t;
    LBM_HD void operator()(int64_t i) const {       // i: node of the outlet plane
        const Grid& g = c.g;
        const int64_t id = (int64_t)(NG + t.p.outlet_row) * g.plane + i, up = id + g.plane, V = g.vol;
        if (!(c.cls[id] & CLS_FLUID) || !(c.cls[up] & CLS_FLUID)) return;
        for (int k = 0; k < t.p.nt; ++k)
            for (int q = 0; q < 5; ++q) t.gC[((int64_t)k * 5 + q) * V + id] = t.gC[((int64_t)k * 5 + q) * V + up];
    }
};

// pull streaming with half-way bounce back + concentration: gC -> g, conc
template <class L>
struct TracerStreamOp {
    CGFields c; TracerFields t;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        int64_t src[L::Q]; bool fl[L::Q];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            src[q] = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            fl[q] = c.cls[src[q]] & CLS_FLUID;
        }
        for (int k = 0; k < t.p.nt; ++k) {
            const double* gC = t.gC + (int64_t)k * L::Q * V; double* gS = t.g + (int64_t)k * L::Q * V;
            double acc = 0.0;
            const double v0 = gC[id];
            gS[id] = v0; acc += v0;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                double v = fl[q] ? gC[q * V + src[q]] : gC[L::opp(q) * V + id];
                if (L::Q == 5 && q == 4 && z == t.p.inlet_row) v = L::w(4) * ((t.p.inlet_conc[k] - acc) / L::w(4));      // Inamuro (calInamuroConstConcBoundary :682-698)
                gS[q * V + id] = v;
                acc += v;
            }
            t.conc[(int64_t)k * V + id] = acc;
        }
    }
};

// g = w C at rest (Transport2DRK.py:461-470)
template <class L>
struct TracerInitOp {
    CGFields c; TracerFields t; const double* conc_in;     // [nt][owned]
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g; const int64_t id = (int64_t)NG * g.plane + i, owned = g.plane * g.n2, V = g.vol;
        const bool fl = c.cls[id] & CLS_FLUID;
        for (int k = 0; k < t.p.nt; ++k) {
            const double C = fl ? conc_in[k * owned + i] : 0.0;
            t.conc[(int64_t)k * V + id] = C;
            for (int q = 0; q < L::Q; ++q) t.g[((int64_t)k * L::Q + q) * V + id] = L::w(q) * C;
        }
    }
};

}  // namespace lbm
