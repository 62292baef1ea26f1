// sc_fast.cuh -- two-pass form of the Shan-Chen loops (original Shan-Chen: ShanChenD2Q9.py:1492-1629; explicit forcing
// SRT / MRT with isotropy 4: ShanChenD2Q9.py:1852-2087), the counterpart of the colour-gradient fast path.
//
// The reference-ordered loop (sc_api.cu::sc_iteration / efs_iteration) keeps the STREAMED populations and moves, per
// iteration and node, ~112 doubles in four launches (explicit forcing: collision 44, streaming 38, force 30).  Here the state
// between iterations is the POST-COLLISION populations only:
//   pass 1  ScPullDensityOp      rho_k(x) = sum_q f*_k,q(x - e_q)   (half-way bounce back through the pull mask)
//                                reads 2Q, writes 2; on the open-boundary planes it also stores the streamed populations
//                                into the destination buffer, where the reference's own row operators (Zou-He inlet,
//                                pressure / convective outlet, ghost rows: sc_ops.cuh::ScOpenRowsOp, unchanged) treat them
//   pass 2  ScPullCollideOp /    pulls the same populations again (row planes: the treated ones), evaluates the
//           EfsPullCollideOp     interaction force from the densities of the neighbours, the common velocity and the
//                                collision in registers and writes the post-collision populations into the other buffer:
//                                reads 2Q + 2 (+ cached neighbours), writes 2Q + 2D
// i.e. ~62 doubles per node in two (closed box) or three launches.  The arithmetic of every node is that of ScStreamOp,
// ScCollideOp, EfsForceOp and EfsCollideOp, operation for operation; only where intermediate values live differs.
#pragma once
#include "sc_ops.cuh"

namespace lbm {

struct ScFast {
    const double* src;       // [nc][Q][vol] post-collision populations of the previous iteration
    double* dst;             // [nc][Q][vol] this iteration's post-collision populations; between the passes its planes
                             //              [0, mat_lo) and [n2 - mat_hi, n2) hold the streamed, boundary-treated populations
    const uint32_t* pull;    // grid.cuh::PullMaskOp (nullptr: no solid node anywhere, every population is pulled)
    int mat_lo, mat_hi;
    LBM_HD bool materialised(const Grid& g, int z) const { return z < mat_lo || z >= g.n2 - mat_hi; }
};

// streaming + densities, nothing else written (calStreaming1GPU/2GPU + calFluidRhoGPU, OptimizedD2Q9GPU.py:450-548, 84-93)
template <class L, int NC>
struct ScPullDensityOp {
    SCFields c; ScFast s;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        const uint32_t m = s.pull ? s.pull[id] : 0xFFFFFFFFu;
        if (!(m & 1u)) return;
        const bool mat = s.materialised(g, z);
        const NbTable nb(g, x, y, z);
        int64_t from[L::Q];     // where population q comes from: the upstream node, or the node's own opposite direction
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t pulled = (int64_t)q * V + nb(-L::d0(q), -L::d1(q), -L::d2(q)), own = (int64_t)L::opp(q) * V + id;
            from[q] = (m >> q & 1u) ? pulled : own;
        }
        double v[NC][L::Q];     // both components requested before either is summed
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double* fk = s.src + (int64_t)k * L::Q * V;
            v[k][0] = fk[id];
#pragma unroll
            for (int q = 1; q < L::Q; ++q) v[k][q] = fk[from[q]];
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            double acc = v[k][0];
#pragma unroll
            for (int q = 1; q < L::Q; ++q) acc += v[k][q];
            c.rho[k * V + id] = acc;
            if (mat) {
                double* dk = s.dst + (int64_t)k * L::Q * V;
#pragma unroll
                for (int q = 0; q < L::Q; ++q) dk[(int64_t)q * V + id] = v[k][q];
            }
        }
    }
};

// the streamed populations of a node, all components: pulled from the post-collision buffer, or -- on the open-boundary
// planes -- what the row operators left in the destination buffer
template <class L, int NC>
LBM_HD void sc_fast_gather(const SCFields& c, const ScFast& s, const NbTable& nb, int z, int64_t id, uint32_t m, double (*f)[L::Q]) {
    const Grid& g = c.g;
    if (s.materialised(g, z)) {
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int q = 0; q < L::Q; ++q) f[k][q] = s.dst[((int64_t)k * L::Q + q) * g.vol + id];
        return;
    }
    int64_t from[L::Q];
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        const int64_t pulled = (int64_t)q * g.vol + nb(-L::d0(q), -L::d1(q), -L::d2(q)), own = (int64_t)L::opp(q) * g.vol + id;
        from[q] = (m >> q & 1u) ? pulled : own;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const double* fk = s.src + (int64_t)k * L::Q * g.vol;
        f[k][0] = fk[id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) f[k][q] = fk[from[q]];
    }
}

// original Shan-Chen: streaming + interactionCollisionProcess (OptimizedD2Q9GPU.py:1274-1446) of one node
template <class L, int NC>
struct ScPullCollideOp {
    SCFields c; ScFast s;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        const uint32_t m = s.pull ? s.pull[id] : 0xFFFFFFFFu;
        if (!(m & 1u)) return;
        // densities of the node and of its neighbours; x + e_q is fluid <=> the upstream node of direction opp(q) is
        const NbTable nb(g, x, y, z);
        double rho[NC], rn[NC][L::Q];
        bool fl[L::Q];
#pragma unroll
        for (int k = 0; k < NC; ++k) rho[k] = c.rho[k * V + id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            fl[q] = m >> L::opp(q) & 1u;
            const int64_t nq = nb(L::d0(q), L::d1(q), L::d2(q));
#pragma unroll
            for (int k = 0; k < NC; ++k) rn[k][q] = c.rho[k * V + nq];      // only used where fl[q] (a solid node holds 0)
        }
        double f[NC][L::Q];
        sc_fast_gather<L, NC>(c, s, nb, z, id, m, f);
        double vt[3] = {0.0, 0.0, 0.0}, rt = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            double mm[3];
            sc_momentum<L>(f[k], mm);
#pragma unroll
            for (int a = 0; a < L::D; ++a) vt[a] += mm[a] / c.p.tau[k];
            rt += rho[k] / c.p.tau[k];
        }
        double up[3];
#pragma unroll
        for (int a = 0; a < L::D; ++a) up[a] = vt[a] / rt;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double psi = rho[k];
            double F[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double wI = L::w(q);
                if (fl[q]) {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const double t = -wI * c.p.G[k * SC_MAXC + j] * psi * rn[j][q];
#pragma unroll
                        for (int a = 0; a < L::D; ++a)
                            if (L::c(q, a) != 0) F[a] += t * L::c(q, a);
                    }
                } else {
                    const double t = -wI * c.p.Gs[k] * psi;
#pragma unroll
                    for (int a = 0; a < L::D; ++a)
                        if (L::c(q, a) != 0) F[a] += t * L::c(q, a);
                }
            }
            const double tau = c.p.tau[k];
            double u[3];
#pragma unroll
            for (int a = 0; a < L::D; ++a) {
                c.Fc(k, a, id) = F[a];
                u[a] = up[a] + tau * F[a] / psi;
            }
            double uu = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a) uu += u[a] * u[a];
            double* dk = s.dst + (int64_t)k * L::Q * V;
#pragma unroll
            for (int q = 0; q < L::Q; ++q) {
                double eu = 0.0;
#pragma unroll
                for (int a = 0; a < L::D; ++a)
                    if (L::c(q, a) != 0) eu += L::c(q, a) * u[a];
                dk[(int64_t)q * V + id] = (1.0 - 1.0 / tau) * f[k][q] + L::w(q) * psi / tau * (1.0 + 3.0 * eu + 4.5 * (eu * eu) - 1.5 * uu);
            }
        }
    }
};

// explicit forcing, isotropy 4: streaming + calExplicit4thOrderScheme + calEquilibriumVEFGPU / transformEquilibriumVelocity +
// calCollisionEXGPU / calAfterCollisionMRT (ExplicitD2Q9GPU.py:51-217, 340-363, 1426-1449, 294-304, 1379-1469) of one node.
// The physical velocity (an output) is left to the reference-ordered tail of the call, which reads the force written here.
template <class L, int NC>
struct EfsPullCollideOp {
    SCFields c; ScFast s;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        const uint32_t m = s.pull ? s.pull[id] : 0xFFFFFFFFu;
        if (!(m & 1u)) return;
        const NbTable nb(g, x, y, z);
        double rho[NC], rn[NC][L::Q];
        bool fl[L::Q];
#pragma unroll
        for (int k = 0; k < NC; ++k) rho[k] = c.rho[k * V + id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            fl[q] = m >> L::opp(q) & 1u;
            const int64_t nq = nb(L::d0(q), L::d1(q), L::d2(q));
#pragma unroll
            for (int k = 0; k < NC; ++k) rn[k][q] = c.rho[k * V + nq];      // only used where fl[q] (a solid node holds 0)
        }
        double f[NC][L::Q];
        sc_fast_gather<L, NC>(c, s, nb, z, id, m, f);
        double F[NC][3];
        double mt[3] = {0.0, 0.0, 0.0}, rt = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double psi = rho[k];
            double gr[3] = {0.0, 0.0, 0.0}, sl[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double wI = L::D == 2 ? (q < 5 ? 1.0 / 3.0 : 1.0 / 12.0) : (q < 7 ? 1.0 / 6.0 : 1.0 / 12.0);
                if (fl[q]) {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const double t = wI * (rn[j][q] - rho[j]);
#pragma unroll
                        for (int a = 0; a < L::D; ++a)
                            if (L::c(q, a) != 0) gr[a] += t * L::c(q, a) * c.p.G[k * SC_MAXC + j];
                    }
                } else {
                    const double t = -wI * c.p.Gs[k] * psi;
#pragma unroll
                    for (int a = 0; a < L::D; ++a)
                        if (L::c(q, a) != 0) sl[a] += t * L::c(q, a);
                }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) F[k][a] = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a) F[k][a] = -6.0 * psi * gr[a] + sl[a];
            double e[3];
            sc_momentum<L>(f[k], e);
            const double wgt = c.p.relax == 0 ? 1.0 / c.p.tau[k] : 1.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a) {
                c.Fc(k, a, id) = F[k][a];
                const double ea = e[a] + 1.0 / 2.0 * F[k][a];
                if (c.p.relax == 0) mt[a] += ea / c.p.tau[k]; else mt[a] += ea * wgt;
            }
            if (c.p.relax == 0) rt = rt + psi / c.p.tau[k]; else rt += psi * wgt;
        }
        double ueq[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < L::D; ++a) ueq[a] = mt[a] / rt;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            double feq[L::Q], ff[L::Q], d[L::Q];
            efs_feq_ff_at<L>(rho[k], ueq, F[k], feq, ff);
#pragma unroll
            for (int q = 0; q < L::Q; ++q) d[q] = feq[q] - f[k][q] - 1.0 / 2.0 * ff[q];
            double* dk = s.dst + (int64_t)k * L::Q * V;
            if (c.p.relax == 0) {
#pragma unroll
                for (int q = 0; q < L::Q; ++q) dk[(int64_t)q * V + id] = f[k][q] + 1.0 / c.p.tau[k] * d[q] + 1.0 * ff[q];
            } else {
                double mo[L::Q], cd[L::Q];
                L::to_moments(d, mo);
                efs_scale_moments<L>(mo, 1.0 / c.p.tau[k], k < 2);
                L::from_moments(mo, cd);
#pragma unroll
                for (int q = 0; q < L::Q; ++q) dk[(int64_t)q * V + id] = f[k][q] + cd[q] + 1.0 * ff[q];
            }
        }
    }
};

}  // namespace lbm
