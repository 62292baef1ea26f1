// cg_fast_ops.cuh -- the colour-gradient step in FACTORED post-collision form (closed boxes: periodic
// and/or bounce-back, no open boundaries).
//
// The reference stores both colours' populations (2 x Q doubles per node) and re-reads them in every one of
// its ~17 kernels.  After recolouring, however, the two colours are not independent:
//     fR*_i = kR fT*_i + w_i e_i . a,    fB*_i = fT*_i - fR*_i,
//     kR = rhoR / rho,   a = beta (rhoR rhoB / rho) G / |G|        (calRecoloringProcessM, 1854-1900)
// so the post-collision state of a node is the TOTAL population fT* (Q doubles) plus 4 scalars (kR, a).
// The fast path keeps exactly that (Q + 4 doubles per node instead of 2 Q) and advances it with two
// passes per time step:
//   pass 1 (density):  pull fT*, kR, a from the neighbours -> rhoR, rhoB, phi of the new time level
//   pass 2 (collide):  pull fT* again -> u (lagged force), G, n, K, F from phi -> MRT/SRT collision in
//                      moment space -> new fT*, kR, a, F
// Half-way bounce back is applied while pulling (own opposite post-collision population when the upstream
// node is solid), the wetting boundary condition enters through phi on the wetting solids and the
// contact-angle correction of G, exactly as in the general path (cg_ops.cuh).
//
// This header holds the straightforward one-thread-per-node form of both passes (also compiled for the
// host by tests/hostcheck); cg_fast.cu holds the tiled sm_100a kernel for pass 2.
#pragma once
#include "cg_ops.cuh"

namespace lbm {

struct FastFields {
    double* gT;      // [Q][vol] post-collision total population
    double* kR;      // [vol]
    double* a;       // [3][vol]
};

// recolouring coefficients of a node (see header comment)
template <class L>
LBM_HD void cg_recolour_coeffs(double rR, double rB, const double* G, double beta, double* kR, double* a) {
    const double gn = sqrt(G[0] * G[0] + G[1] * G[1] + (L::D == 3 ? G[2] * G[2] : 0.0));
    const double inv = 1.0 / (rR + rB);
    *kR = rR * inv;
    const double amp = gn > 1.0e-8 ? beta * rR * rB * inv / gn : 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) a[d] = d < L::D ? amp * G[d] : 0.0;
}

// red part of post-collision population q of node `id`.  pert: the recolouring of the perturbation-operator model weighs the
// direction cosine, fR_i = kR fT_i + beta rhoR rhoB / rho^2 w_i (e_i . G) / (|e_i| |G|) (calRKCollision23GPUNew,
// AcceleratedRKGPU2D.py:1169-1266), so its vector a = beta rhoR rhoB / rho^2 G / |G| enters with w_i / |e_i|
template <class L>
LBM_HD double cg_red_weight(int q, bool pert) { return pert ? L::w(q) / L::enorm(q) : L::w(q); }
template <class L>
LBM_HD double cg_red_part(int q, double gT, double kR, const double* a, bool pert = false) {
    double ea = 0.0;
#pragma unroll
    for (int d = 0; d < L::D; ++d)
        if (L::c(q, d) != 0) ea += L::c(q, d) * a[d];
    return q == 0 ? kR * gT : kR * gT + cg_red_weight<L>(q, pert) * ea;
}

// general-path collision writing the factored state (entry into the fast path from a streamed state):
// same arithmetic as CollideOp up to the recolouring, which is stored as (kR, a)
template <class L>
struct CollideFactoredOp {
    CGFields c; FastFields o;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double G[3] = {0, 0, 0}, n[3] = {0, 0, 0}, u[3] = {0, 0, 0}, F[3] = {0, 0, 0}, K;
#pragma unroll
        for (int d = 0; d < L::D; ++d) { G[d] = c.G[d * V + id]; n[d] = c.nrm[d * V + id]; u[d] = c.u[d * V + id]; }
        cg_force_at<L>(c, x, y, z, id, G, n, F, &K);
#pragma unroll
        for (int d = 0; d < L::D; ++d) c.F[d * V + id] = F[d];
        c.K[id] = K;
        double fT[L::Q];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) fT[q] = c.fS[0][q * V + id] + c.fS[1][q * V + id];
        const double rR = c.rho[0][id], rB = c.rho[1][id];
        const double tau = cg_tau(c.phi[id], rR, rB, c.p);
        cg_collide<L>(fT, rR + rB, u, F, tau, c.p.relax);
        double kR, a[3];
        cg_recolour_coeffs<L>(rR, rB, G, c.p.beta, &kR, a);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) o.gT[q * V + id] = fT[q];
        o.kR[id] = kR;
#pragma unroll
        for (int d = 0; d < 3; ++d) o.a[d * V + id] = a[d];
    }
};

// pass 1: densities and colour field of the new time level from the factored state
// (streaming + half-way bounce back: calStreaming1GPU/2GPU 338-417; calMacroDensityRKGPU2D 101-118;
//  calPhaseFieldPhi 1347-1356)
template <class L, bool SOLIDS>
struct PullDensityOp {
    CGFields c; FastFields s;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (SOLIDS && !(c.cls[id] & CLS_FLUID)) return;
        const double kR0 = s.kR[id];
        const double a0[3] = {s.a[id], s.a[V + id], s.a[2 * V + id]};
        const double g0 = s.gT[id];
        const bool pert = c.p.st_type == LBM_ST_PERTURBATION;
        double accR = kR0 * g0, accB = g0 - accR;
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            double gt, fr;
            if (!SOLIDS || (c.cls[src] & CLS_FLUID)) {
                gt = s.gT[q * V + src];
                const double an[3] = {s.a[src], s.a[V + src], s.a[2 * V + src]};
                fr = cg_red_part<L>(q, gt, s.kR[src], an, pert);
            } else {
                gt = s.gT[L::opp(q) * V + id];
                fr = cg_red_part<L>(L::opp(q), gt, kR0, a0, pert);
            }
            accR += fr; accB += gt - fr;
        }
        c.rho[0][id] = accR; c.rho[1][id] = accB;
        c.phi[id] = (accR - accB) / (accR + accB);
    }
};

// leave the fast path: materialise the streamed populations of both colours and their densities
template <class L>
struct PullMaterialiseOp {
    CGFields c; FastFields s;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const double kR0 = s.kR[id];
        const double a0[3] = {s.a[id], s.a[V + id], s.a[2 * V + id]};
        const double g0 = s.gT[id];
        double accR = kR0 * g0, accB = g0 - accR;
        c.fS[0][id] = accR; c.fS[1][id] = accB;
        const bool pert = c.p.st_type == LBM_ST_PERTURBATION;
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            double gt, fr;
            if (c.cls[src] & CLS_FLUID) {
                gt = s.gT[q * V + src];
                const double an[3] = {s.a[src], s.a[V + src], s.a[2 * V + src]};
                fr = cg_red_part<L>(q, gt, s.kR[src], an, pert);
            } else {
                gt = s.gT[L::opp(q) * V + id];
                fr = cg_red_part<L>(L::opp(q), gt, kR0, a0, pert);
            }
            c.fS[0][q * V + id] = fr; c.fS[1][q * V + id] = gt - fr;
            accR += fr; accB += gt - fr;
        }
        c.rho[0][id] = accR; c.rho[1][id] = accB;
    }
};

// pass 2, straightforward form: pull the total population, velocity with the lagged force, force from the
// unit normals in c.nrm (GradientOp must have run on the new phi), collision, new factored state.
template <class L, bool SOLIDS>
struct PullCollideOp {
    CGFields c; FastFields s, o;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (SOLIDS && !(c.cls[id] & CLS_FLUID)) return;
        double fT[L::Q];
        fT[0] = s.gT[id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            fT[q] = (!SOLIDS || (c.cls[src] & CLS_FLUID)) ? s.gT[q * V + src] : s.gT[L::opp(q) * V + id];
        }
        const double rR = c.rho[0][id], rB = c.rho[1][id], rho = rB + rR;
        double mom[3] = {0.0, 0.0, 0.0}, u[3] = {0, 0, 0}, G[3] = {0, 0, 0}, n[3] = {0, 0, 0}, F[3] = {0, 0, 0}, K;
#pragma unroll
        for (int q = 1; q < L::Q; ++q)
#pragma unroll
            for (int d = 0; d < L::D; ++d)
                if (L::c(q, d) != 0) mom[d] += L::c(q, d) * fT[q];
#pragma unroll
        for (int d = 0; d < L::D; ++d) {
            u[d] = (mom[d] + 0.5 * c.F[d * V + id]) / rho;      // force of the previous step
            G[d] = c.G[d * V + id]; n[d] = c.nrm[d * V + id];
        }
        cg_force_at<L>(c, x, y, z, id, G, n, F, &K);
#pragma unroll
        for (int d = 0; d < L::D; ++d) c.F[d * V + id] = F[d];
        if (c.store_u) {          // solute tracers ride on this step: their collision (after this pass) reads the velocity
#pragma unroll
            for (int d = 0; d < L::D; ++d) c.u[d * V + id] = u[d];
        }
        const double tau = cg_tau(c.phi[id], rR, rB, c.p);
        cg_collide<L>(fT, rho, u, F, tau, c.p.relax);
        double kR, a[3];
        cg_recolour_coeffs<L>(rR, rB, G, c.p.beta, &kR, a);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) o.gT[q * V + id] = fT[q];
        o.kR[id] = kR;
#pragma unroll
        for (int d = 0; d < 3; ++d) o.a[d * V + id] = a[d];
    }
};

// ---- the perturbation-operator model (LBM_ST_PERTURBATION) in factored form -------------------------------------------
// Same two passes; the collision side is PerturbCollideOp (cg_ops.cuh) up to the recolouring, which is stored as (kR, a) with
// a = beta rhoR rhoB / rho^2 G / |G| (exactly zero gradient: a = 0, like the reference's cos(theta) = 0).
// fT: total population of the node (in: streamed, out: post-collision incl. the perturbation term); G: colour gradient.
template <class L>
LBM_HD void cgp_collide_factored(double* fT, double rR, double rB, double phi, const double* u, const double* G, const CGParams& p,
                                 double* kR, double* a) {
    const double rho = rB + rR;
    double d[L::Q], m[L::NMOM];
    double uu = 0.0;
#pragma unroll
    for (int k = 0; k < L::D; ++k) uu += u[k] * u[k];
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
        double eu = 0.0;
#pragma unroll
        for (int k = 0; k < L::D; ++k)
            if (L::c(q, k) != 0) eu += L::c(q, k) * u[k];
        d[q] = fT[q] - rho * L::w(q) * (1.0 + (3.0 * eu + 4.5 * eu * eu - 1.5 * uu));
    }
    const double tau = 0.5 + 1.0 / ((1.0 + phi) / (2.0 * (p.tauR - 0.5)) + (1.0 - phi) / (2.0 * (p.tauB - 0.5)));
    if (p.relax == 0) {
#pragma unroll
        for (int q = 0; q < L::Q; ++q) d[q] = 1.0 / tau * d[q];
    } else {
        L::to_moments(d, m);
        PerturbCollideOp<L>::scale_moments(m, 1.0 / tau);
        L::from_moments(m, d);
    }
    double g2 = mul_rn(G[0], G[0]);
#pragma unroll
    for (int k = 1; k < L::D; ++k) g2 = add_rn(g2, mul_rn(G[k], G[k]));
    const double gn = sqrt(g2);
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
        double eF = 0.0, eg = 0.0;
#pragma unroll
        for (int k = 0; k < L::D; ++k)
            if (L::c(q, k) != 0) { eF += L::c(q, k) * p.bf[k]; eg += L::c(q, k) * G[k]; }
        double f = -d[q] + (q == 0 ? 0.0 : 3.0 * L::w(q)) * eF + fT[q];
        if (g2 != 0.0) f += p.Ak * gn * (L::w(q) * (eg * eg) / g2 - (q == 0 ? L::w(0) - 2.0 / 3.0 : L::w(q)));
        fT[q] = f;
    }
    *kR = rR / rho;
    const double amp = gn != 0.0 ? p.beta * (rR * rB) / (rho * rho) / gn : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = k < L::D ? amp * G[k] : 0.0;
}

// colour gradient of the perturbation model at (x, y, z): SolidColorDiff on solid neighbours, products rounded one by one
template <class L>
LBM_HD void cgp_gradient_at(const CGFields& c, int x, int y, int z, double* G) {
    const Grid& g = c.g;
    G[0] = G[1] = G[2] = 0.0;
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        const int64_t n = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
        const double pk = (c.cls[n] & CLS_FLUID) ? c.phi[n] : c.p.solid_phi;
#pragma unroll
        for (int k = 0; k < L::D; ++k)
            if (L::c(q, k) != 0) G[k] = add_rn(G[k], mul_rn(3.0 * L::w(q) * L::c(q, k), pk));
    }
}

// entry into the fast path / the patched open rows: streamed populations (fS, rho, u, phi) -> factored post-collision state
template <class L>
struct PerturbCollideFactoredOp {
    CGFields c; FastFields o;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double fT[L::Q], u[3] = {0.0, 0.0, 0.0}, G[3], kR, a[3];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) fT[q] = c.fS[0][q * V + id] + c.fS[1][q * V + id];
#pragma unroll
        for (int k = 0; k < L::D; ++k) u[k] = c.u[k * V + id];
        cgp_gradient_at<L>(c, x, y, z, G);
        cgp_collide_factored<L>(fT, c.rho[0][id], c.rho[1][id], c.phi[id], u, G, c.p, &kR, a);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) o.gT[q * V + id] = fT[q];
        o.kR[id] = kR;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o.a[k * V + id] = a[k]; c.G[k * V + id] = G[k]; }
    }
};

// pass 2 of the perturbation model, one thread per node: pull the total population, velocity (no force term:
// calPhysicalVelocityRKGPU2D), gradient from phi, collision + perturbation, new factored state
template <class L, bool SOLIDS>
struct PullPerturbCollideOp {
    CGFields c; FastFields s, o;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (SOLIDS && !(c.cls[id] & CLS_FLUID)) return;
        double fT[L::Q];
        fT[0] = s.gT[id];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t src = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            fT[q] = (!SOLIDS || (c.cls[src] & CLS_FLUID)) ? s.gT[q * V + src] : s.gT[L::opp(q) * V + id];
        }
        const double rR = c.rho[0][id], rB = c.rho[1][id], rho = rB + rR;
        double mom[3] = {0.0, 0.0, 0.0}, u[3] = {0.0, 0.0, 0.0}, G[3], kR, a[3];
#pragma unroll
        for (int q = 1; q < L::Q; ++q)
#pragma unroll
            for (int k = 0; k < L::D; ++k)
                if (L::c(q, k) != 0) mom[k] += L::c(q, k) * fT[q];
#pragma unroll
        for (int k = 0; k < L::D; ++k) u[k] = mom[k] / rho;      // HeadOp's expression with the zero force of this model
        cgp_gradient_at<L>(c, x, y, z, G);
        cgp_collide_factored<L>(fT, rR, rB, c.phi[id], u, G, c.p, &kR, a);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) o.gT[q * V + id] = fT[q];
        o.kR[id] = kR;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o.a[k * V + id] = a[k]; c.G[k * V + id] = G[k]; }
    }
};

}  // namespace lbm
