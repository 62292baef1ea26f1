// cg_ops.cuh -- colour-gradient (CSF) step as one-thread-per-node operators on the dense grid, in the
// order of the reference's loop body (RKD2Q9.py:1295-1490).  This is the GENERAL path: any mask,
// wetting, open boundaries, SRT/MRT, D2Q9/D3Q19.  Populations are structure-of-arrays [Q][vol].
// The fused fast path for closed (periodic / bounce-back) boxes lives in cg_fast.cu.
#pragma once
#include "../../include/lbmpm.h"
#include "grid.cuh"

namespace lbm {

struct CGFields {
    Grid g;
    CGParams p;
    double* fS[2];      // streamed populations (what the reference holds between iterations), R and B
    double* fC[2];      // post-collision populations
    double* rho[2];     // rhoR, rhoB
    double* u;          // [3][vol]
    double* phi;        // colour field; on wetting solids: calColorValueOnSolid's value
    double* G;          // [3][vol] colour gradient (after the wetting correction)
    double* nrm;        // [3][vol] interface unit normal
    double* F;          // [3][vol] CSF force (lagged by one step when the velocity is evaluated)
    double* K;          // curvature
    const uint8_t* cls;
    const uint32_t* pull;   // [vol] pull masks (tiled kernels, lattices with solids)
    int store_u;            // fast path: the collision pass also stores the velocity (tracers attached)
    const double* ns;   // [3][vol] solid normals
    // open boundaries (global plane numbers along axis 2; -1000 = not on this slab)
    int inlet, outlet;
    int z_in, z_in_ghost, z_out, z_out_ghost, z_out2;
    double v_in, rhoBH, rhoRH, rhoBL, rhoRL;
};

// f = w_i rho at rest (RKD2Q9.py:561-585 with u = 0); also clears the lagged force
template <class L>
struct InitEquilibriumOp {
    CGFields c; const double* rhoR_in; const double* rhoB_in;    // dense owned [n2][n1][n0] on device
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * c.g.plane + i;
        const bool fl = c.cls[id] & CLS_FLUID;
        const double r[2] = {fl ? rhoR_in[i] : 0.0, fl ? rhoB_in[i] : 0.0};
        for (int k = 0; k < 2; ++k) {
            c.rho[k][id] = r[k];
#pragma unroll
            for (int q = 0; q < L::Q; ++q) c.fS[k][q * c.g.vol + id] = L::w(q) * r[k];
        }
        for (int a = 0; a < 3; ++a) { c.F[a * c.g.vol + id] = 0.0; c.u[a * c.g.vol + id] = 0.0; }
    }
};

// AoS [node][Q] (reference layout, RKD2Q9.py:451-452) <-> SoA [Q][vol]
template <class L>
struct AosToSoaOp {
    Grid g; const double* aos; double* soa; const uint8_t* cls; double* rho; const double* rho_in;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * g.plane + i;
        const bool fl = cls[id] & CLS_FLUID;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            const double v = fl ? aos[i * L::Q + q] : 0.0;
            soa[q * g.vol + id] = v;
            s = q == 0 ? v : s + v;
        }
        rho[id] = fl ? (rho_in ? rho_in[i] : s) : 0.0;
    }
};
template <class L>
struct SoaToAosOp {
    Grid g; const double* soa; double* aos;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * g.plane + i;
#pragma unroll
        for (int q = 0; q < L::Q; ++q) aos[i * L::Q + q] = soa[q * g.vol + id];
    }
};

// ------------------------------------------------------------------------------------------------
// open boundaries of the reference's driver (planes along the flow axis; inlet = top, outlet = bottom)
// ------------------------------------------------------------------------------------------------
// The three row treatments below are written for any lattice: "up" is the slab axis (y in 2-D, z in 3-D),
// the unknown populations are those pointing into the domain, N_t = 1/2 sum_{in-plane} c_t f is the
// transverse momentum that Zou-He redistributes over the diagonal unknowns.  For D2Q9 they are the
// reference's formulas term by term; D3Q19 (no reference code) is their Hecht-Harting generalisation.
// Items: one plane of n0 * n1 nodes.

// constantTotalVelocityInlet (AcceleratedRKGPU2D.py:2345-2423): non-equilibrium bounce-back of the
// unknown populations of the TOTAL distribution on plane z_in with u = v e_up, split by mass fraction.
template <class L>
struct InletVelocityOp {
    CGFields c;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t id = (int64_t)(c.z_in + NG) * g.plane + r;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const int64_t V = g.vol;
        double* fR = c.fS[0]; double* fB = c.fS[1];
        double fT[L::Q];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) fT[q] = fR[q * V + id] + fB[q * V + id];
        const double v = c.v_in;
        double s0 = 0.0, sp = 0.0;
        bool first0 = true, firstp = true;
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            if (L::d2(q) == 0) { s0 = first0 ? fT[q] : s0 + fT[q]; first0 = false; }
            if (L::d2(q) == 1) { sp = firstp ? fT[q] : sp + fT[q]; firstp = false; }
        }
        const double rho = (s0 + 2.0 * sp) / (1.0 + v);
        const double vv = v * v;
        double rR = c.rho[0][id], rB = c.rho[1][id];
        const double ratioR = rR / (rR + rB);
        rR = ratioR * rho;
        const double ratioB = rB / (rR + rB);     // uses the already-updated rhoR, like :2399-2407
        rB = ratioB * rho;
        c.rho[0][id] = rR; c.rho[1][id] = rB;
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            if (L::d2(q) != -1) continue;
            const double ev = -v, evo = v;        // e_q . u and e_opp . u
            const double eq = rho * L::w(q) * (1.0 + 3.0 * ev + 4.5 * ev * ev - 1.5 * vv);
            const double eqo = rho * L::w(q) * (1.0 + 3.0 * evo + 4.5 * evo * evo - 1.5 * vv);
            const double t = eq + (fT[L::opp(q)] - eqo);
            fR[q * V + id] = ratioR * t;
            fB[q * V + id] = ratioB * t;
        }
    }
};
// calConstPressureInletGPU (AcceleratedRKGPU2D.py:923-961): Zou-He pressure per colour on plane z_in
template <class L>
struct InletPressureOp {
    CGFields c;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t id = (int64_t)(c.z_in + NG) * g.plane + r;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const int64_t V = g.vol;
        for (int k = 0; k < 2; ++k) {
            double* f = c.fS[k];
            const double p = k == 0 ? c.rhoRH : c.rhoBH;
            double fl[L::Q];
#pragma unroll
            for (int q = 0; q < L::Q; ++q) fl[q] = f[q * V + id];
            double s0 = 0.0, sp = 0.0, N[3] = {0.0, 0.0, 0.0};
            bool first0 = true, firstp = true;
#pragma unroll
            for (int q = 0; q < L::Q; ++q) {
                if (L::d2(q) == 0) {
                    s0 = first0 ? fl[q] : s0 + fl[q]; first0 = false;
                    if (L::d0(q) != 0) N[0] += L::d0(q) * fl[q];
                    if (L::d1(q) != 0) N[1] += L::d1(q) * fl[q];
                }
                if (L::d2(q) == 1) { sp = firstp ? fl[q] : sp + fl[q]; firstp = false; }
            }
            const double v = -1.0 + (s0 + 2.0 * sp) / p;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                if (L::d2(q) != -1) continue;
                // f_q = f_opp - 1/2 (e_q . N) - 6 w_q p v      (2-D: f4 = f2 - 2/3 p v, f7 = f5 + (f1-f3)/2 - p v/6, ...)
                f[q * V + id] = fl[L::opp(q)] + 1.0 / 2.0 * (-(L::d0(q) * N[0] + L::d1(q) * N[1])) - 6.0 * L::w(q) * p * v;
            }
            c.rho[k][id] = p;
        }
    }
};
// calConstPressureLowerGPUTotal (AcceleratedRKGPU2D.py:2557-2602): Zou-He pressure on the total
// distribution on plane z_out with p = rhoBL + rhoRL (RKD2Q9.py:1344), split by mass fraction
template <class L>
struct OutletPressureOp {
    CGFields c;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t id = (int64_t)(c.z_out + NG) * g.plane + r;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const int64_t V = g.vol;
        double* fR = c.fS[0]; double* fB = c.fS[1];
        double fT[L::Q];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) fT[q] = fR[q * V + id] + fB[q * V + id];
        const double p = c.rhoBL + c.rhoRL;
        double s0 = 0.0, sm = 0.0, N[3] = {0.0, 0.0, 0.0};
        bool first0 = true, firstm = true;
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            if (L::d2(q) == 0) {
                s0 = first0 ? fT[q] : s0 + fT[q]; first0 = false;
                if (L::d0(q) != 0) N[0] += L::d0(q) * fT[q];
                if (L::d1(q) != 0) N[1] += L::d1(q) * fT[q];
            }
            if (L::d2(q) == -1) { sm = firstm ? fT[q] : sm + fT[q]; firstm = false; }
        }
        const double v = 1.0 - 1.0 / p * (s0 + 2.0 * sm);
        const double rR = c.rho[0][id], rB = c.rho[1][id];
        const double ratioR = rR / (rR + rB), ratioB = rB / (rR + rB);
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            if (L::d2(q) != 1) continue;
            // f_q = f_opp - 1/2 (e_q . N) + 6 w_q p v       (2-D: f2 = f4 + 2/3 p v, f5 = f7 + (f3-f1)/2 + p v/6, ...)
            const double t = fT[L::opp(q)] + 0.5 * (-(L::d0(q) * N[0] + L::d1(q) * N[1])) + 6.0 * L::w(q) * (p * v);
            fR[q * V + id] = ratioR * t;
            fB[q * V + id] = ratioB * t;
        }
    }
};
// Row copies: ghostPointsConstantVelocityRK (604-650), ghostPointsConstPressureInletRK (966-1001),
// ghostPointsConstPressureLowerRK (1043-1080), convectiveOutletGPU/Ghost2/Ghost3 (698-784):
// row z_dst <- row z_src (populations of both colours); density = sum (sum_rho) or copied.
// Rows are whole planes of axis 2, so the operator is lattice-generic.
template <class L>
struct RowCopyOp {
    CGFields c; int z_dst, z_src; int sum_rho;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t d = (int64_t)(z_dst + NG) * g.plane + r, s = (int64_t)(z_src + NG) * g.plane + r;
        if (!(c.cls[d] & CLS_FLUID) || !(c.cls[s] & CLS_FLUID)) return;
        for (int k = 0; k < 2; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < L::Q; ++q) {
                const double v = c.fS[k][q * g.vol + s];
                c.fS[k][q * g.vol + d] = v;
                acc = q == 0 ? v : acc + v;
            }
            c.rho[k][d] = sum_rho ? acc : c.rho[k][s];
        }
    }
};

// All open-boundary rows of one iteration in ONE launch (RKD2Q9.py:1299-1352 launches 2 + 2..3 kernels over every node;
// the operators above, one launch each, were 4-5 launches of a launch-bound 2-D step).  Items: 2 planes' worth of
// columns -- the first `plane` items walk the outlet rows of their column, the others its inlet rows.  The rows of one
// column only depend on each other (treated row -> ghost row, row 3 -> 2 -> 1 -> 0), so running them back to back in one
// thread is the same arithmetic in the same order.
template <class L>
struct OpenRowsOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const int64_t plane = c.g.plane;
        if (i < plane) {
            if (c.z_out < 0) return;
            if (c.outlet == LBM_OUTLET_CONVECTIVE) {
                RowCopyOp<L>{c, 2, 3, 1}(i); RowCopyOp<L>{c, 1, 2, 1}(i); RowCopyOp<L>{c, 0, 1, 1}(i);
            } else if (c.outlet == LBM_OUTLET_PRESSURE) {
                OutletPressureOp<L>{c}(i); RowCopyOp<L>{c, 0, 1, 0}(i);
            }
        } else {
            const int64_t r = i - plane;
            if (c.z_in < 0) return;
            if (c.inlet == LBM_INLET_VELOCITY) {
                InletVelocityOp<L>{c}(r); RowCopyOp<L>{c, c.z_in_ghost, c.z_in, 1}(r);
            } else if (c.inlet == LBM_INLET_PRESSURE) {
                InletPressureOp<L>{c}(r); RowCopyOp<L>{c, c.z_in_ghost, c.z_in, 0}(r);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// the loop body
// ------------------------------------------------------------------------------------------------
// calTotalFluidPDF (1413-1422) + calPhysicalVelocityRKGPU2DNew1 (2632-2653) + calPhaseFieldPhi
// (1347-1356): u = (sum e fT + F_lagged / 2) / (rhoR + rhoB), phi = (rhoR - rhoB) / (rhoR + rhoB)
template <class L>
struct HeadOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        const int64_t id = (int64_t)NG * g.plane + i, V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double mom[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const double fT = c.fS[0][q * V + id] + c.fS[1][q * V + id];
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) mom[a] += L::c(q, a) * fT;
        }
        const double rR = c.rho[0][id], rB = c.rho[1][id];
        const double rho = rB + rR;
#pragma unroll
        for (int a = 0; a < L::D; ++a) c.u[a * V + id] = (mom[a] + 0.5 * c.F[a * V + id]) / rho;
        c.phi[id] = (rR - rB) / (rR + rB);
    }
};

// calColorValueOnSolid (1559-1580): phi_s = sum_{fluid nb} w phi / sum w of the wetting solid at (x, y, z)
template <class L>
LBM_HD double cg_phi_on_solid(const CGFields& c, int x, int y, int z) {
    const Grid& g = c.g;
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        const int64_t n = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
        if (c.cls[n] & CLS_FLUID) { num = add_rn(num, mul_rn(L::w(q), c.phi[n])); den += L::w(q); }
    }
    return den > 0.0 ? num / den : 0.0;
}
// ... stored into the phi array, planes [-2, n2+2): what the tiled collision kernel stages in shared memory and what
// lbm_download_fields returns.  The one-thread-per-node gradient evaluates it in place (GradientOp) and needs no such pass.
template <class L>
struct PhiSolidOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 2, x, y, z);
        const int64_t id = g.at(x, y, z);
        if (!(c.cls[id] & CLS_WET)) return;
        c.phi[id] = cg_phi_on_solid<L>(c, x, y, z);
    }
};

// the same over the compact list of wetting solids (grid.cuh::WetFillOp)
template <class L>
struct PhiSolidListOp {
    CGFields c; const int64_t* list;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        const int64_t id = list[i];
        const int z = (int)(id / g.plane) - NG;
        const int r = (int)(id % g.plane);
        const int y = r / g.n0, x = r - y * g.n0;
        c.phi[id] = cg_phi_on_solid<L>(c, x, y, z);
    }
};

// calRKInitialGradient (1582-1632) + updateColorGradientOnWetting[New] (1637-1679 / 2428-2492) and the
// unit normal that the curvature stencil gathers; planes [-1, n2+1)
template <class L>
struct GradientOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 1, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double G[3] = {0.0, 0.0, 0.0};
        const bool near_solid = c.cls[id] & CLS_NEAR;
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            int xn, yn, zn;
            g.nb_coords(x, y, z, L::d0(q), L::d1(q), L::d2(q), xn, yn, zn);
            const int64_t nbid = g.at(xn, yn, zn);
            // the colour a wetting solid shows is a function of its fluid neighbours' phi: evaluated here instead of in a
            // pass of its own over the lattice (calColorValueOnSolid + calRKInitialGradient, 1559-1632)
            const double pv = (near_solid && (c.cls[nbid] & CLS_WET)) ? cg_phi_on_solid<L>(c, xn, yn, zn) : c.phi[nbid];
            const double v = mul_rn(L::w(q), pv);
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) G[a] = add_rn(G[a], L::c(q, a) > 0 ? v : -v);
        }
#pragma unroll
        for (int a = 0; a < L::D; ++a) G[a] *= 3.0;
        if (c.cls[id] & CLS_NEAR) {
            double ns[3] = {c.ns[id], c.ns[V + id], L::D == 3 ? c.ns[2 * V + id] : 0.0};
            cg_wetting<L::D>(G, ns, c.p.cosT, c.p.sinT, c.p.wetting_type);
        }
        double n[3];
        cg_unit_normal<L::D>(G, c.p.wetting_type, n);
#pragma unroll
        for (int a = 0; a < L::D; ++a) { c.G[a * V + id] = G[a]; c.nrm[a * V + id] = n[a]; }
    }
};

// curvature and CSF force from the neighbours' unit normals: calForceTermInColorGradient[New]2D
// (1684-1735 / 2497-2552), K = n_a n_b d_a n_b - (n.n) d_a n_a (== the reference's 2-D expression)
template <class L>
LBM_HD void cg_force_at(const CGFields& c, int x, int y, int z, int64_t id, const double* G, const double* n,
                        double* F, double* Kout) {
    const Grid& g = c.g; const int64_t V = g.vol;
    double dn[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};   // dn[a][b] = d_a n_b
#pragma unroll
    for (int q = 1; q < L::Q; ++q) {
        const int64_t nbid = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
        double nk[3];
#pragma unroll
        for (int b = 0; b < L::D; ++b) nk[b] = c.nrm[b * V + nbid];   // zero on solid nodes
#pragma unroll
        for (int a = 0; a < L::D; ++a)
            if (L::c(q, a) != 0)
#pragma unroll
                for (int b = 0; b < L::D; ++b) dn[a][b] = add_rn(dn[a][b], mul_rn(3.0 * L::w(q) * L::c(q, a), nk[b]));
    }
    double K = 0.0, nn = 0.0, div = 0.0;
#pragma unroll
    for (int a = 0; a < L::D; ++a) {
        nn += n[a] * n[a]; div += dn[a][a];
#pragma unroll
        for (int b = 0; b < L::D; ++b) K += n[a] * n[b] * dn[a][b];
    }
    K -= nn * div;
    const double sgn = c.p.wetting_type == 1 ? 0.5 : -0.5;
#pragma unroll
    for (int a = 0; a < L::D; ++a) F[a] = sgn * c.p.sigma * K * G[a];
    *Kout = K;
    (void)id;
}

// curvature only (lbm_download_fields): same stencil as the collision operators, the force state is left alone
template <class L>
struct CurvatureOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double G[3] = {0, 0, 0}, n[3] = {0, 0, 0}, F[3], K;
#pragma unroll
        for (int a = 0; a < L::D; ++a) { G[a] = c.G[a * V + id]; n[a] = c.nrm[a * V + id]; }
        cg_force_at<L>(c, x, y, z, id, G, n, F, &K);
        c.K[id] = K;
    }
};

// force + collision of the total population + recolouring, fS -> fC
// (calForceTermInColorGradient*, calRKCollision1TotalGPU2D{SRT,MRT}M, calPerturbationFromForce2D[MRT],
//  calRecoloringProcessM; RKD2Q9.py:1419-1465)
template <class L>
struct CollideOp {
    CGFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double G[3] = {0, 0, 0}, n[3] = {0, 0, 0}, u[3] = {0, 0, 0}, F[3] = {0, 0, 0}, K;
#pragma unroll
        for (int a = 0; a < L::D; ++a) { G[a] = c.G[a * V + id]; n[a] = c.nrm[a * V + id]; u[a] = c.u[a * V + id]; }
        cg_force_at<L>(c, x, y, z, id, G, n, F, &K);
#pragma unroll
        for (int a = 0; a < L::D; ++a) c.F[a * V + id] = F[a];
        c.K[id] = K;
        double fT[L::Q], fR[L::Q], fB[L::Q];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) fT[q] = c.fS[0][q * V + id] + c.fS[1][q * V + id];
        const double rR = c.rho[0][id], rB = c.rho[1][id];
        const double tau = cg_tau(c.phi[id], rR, rB, c.p);
        cg_collide<L>(fT, rR + rB, u, F, tau, c.p.relax);
        cg_recolour<L>(fT, rR, rB, G, c.p.beta, fR, fB);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) { c.fC[0][q * V + id] = fR[q]; c.fC[1][q * V + id] = fB[q]; }
    }
};

// pull streaming with half-way bounce back + densities: calStreaming1GPU/2GPU (338-417) and
// calMacroDensityRKGPU2D (101-118).  fC -> fS
template <class L>
struct StreamOp {
    CGFields c;
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
        for (int k = 0; k < 2; ++k) {
            const double* fC = c.fC[k]; double* fS = c.fS[k];
            double acc = fC[id];
            fS[id] = acc;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double v = fl[q] ? fC[q * V + src[q]] : fC[L::opp(q) * V + id];
                fS[q * V + id] = v;
                acc += v;
            }
            c.rho[k][id] = acc;
        }
    }
};

// ---- colour gradient with the PERTURBATION surface-tension operator (LBM_ST_PERTURBATION) -------------------------
// One operator for the collision side of runRKColorGradient2DPerturbation's MRT branch (RKD2Q9.py:1157-1219):
//   calRKCollision1GPU2DMRTNew (AcceleratedRKGPU2D.py:1272-1343): fT <- fT - M^-1 S M (fT - feq(rho, u)) + w_F e.F_body,
//       tau(phi) = 1/2 + 1 / ((1 + phi) / (2 (tauR - 1/2)) + (1 - phi) / (2 (tauB - 1/2))), w_F = 3 w_i;
//   calRKCollision23GPUNew (1169-1266): G = 3 sum_k w_k e_k phi(x + e_k), SolidColorDiff on solid neighbours;
//       fT_i += (A_R + A_B)/2 |G| (w_i (e_i.G)^2 / |G|^2 - B_i)   (exactly zero gradient: nothing);
//       fR_i = rho_R/rho fT_i + beta rho_R rho_B / rho^2 w_i cos(theta_i),  fB_i = rho_B/rho fT_i - ...
// B = (w_0 - 2/3, w_i): (-2/9, 1/9, 1/36) for D2Q9 (RKD2Q9.py:131-133), (-1/3, 1/18, 1/36) for D3Q19 (Liu et al. 2012).
// The gradient products are rounded one by one in the reference's order (mul_rn / add_rn): the kernel tests
// `G.G == 0` exactly and normalises any other G, however small, in the recolouring term.
// fS (streamed populations), rho, u, phi -> fC; also leaves G for lbm_download_fields.
template <class L>
struct PerturbCollideOp {
    CGFields c;
    LBM_HD static void scale_moments(double* m, double s_nu) {
        if (L::Q == 9) {      // S = (0, 1.64, 1.54, 0, 1.9, 0, 1.9, 1/tau, 1/tau), RKD2Q9.py:338-340
            m[0] = 0.0; m[1] *= 1.64; m[2] *= 1.54; m[3] = 0.0; m[4] *= 1.9; m[5] = 0.0; m[6] *= 1.9; m[7] *= s_nu; m[8] *= s_nu;
        } else {              // the rates of the D3Q19 specification (lattice.cuh::D3Q19::relax_moments)
            m[0] = 0.0; m[1] *= 1.19; m[2] *= 1.4; m[3] = 0.0; m[4] *= 1.2; m[5] = 0.0; m[6] *= 1.2; m[7] = 0.0; m[8] *= 1.2;
            m[9] *= s_nu; m[10] *= 1.4; m[11] *= s_nu; m[12] *= 1.4; m[13] *= s_nu; m[14] *= s_nu; m[15] *= s_nu;
            m[16] *= 1.98; m[17] *= 1.98; m[18] *= 1.98;
        }
    }
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const double rR = c.rho[0][id], rB = c.rho[1][id], rho = rB + rR, phi = c.phi[id];
        double u[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int a = 0; a < L::D; ++a) u[a] = c.u[a * V + id];
        double fT[L::Q], d[L::Q], m[L::NMOM];
        double uu = 0.0;
#pragma unroll
        for (int a = 0; a < L::D; ++a) uu += u[a] * u[a];
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            fT[q] = c.fS[0][q * V + id] + c.fS[1][q * V + id];
            double eu = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) eu += L::c(q, a) * u[a];
            d[q] = fT[q] - rho * L::w(q) * (1.0 + (3.0 * eu + 4.5 * eu * eu - 1.5 * uu));
        }
        const double tau = 0.5 + 1.0 / ((1.0 + phi) / (2.0 * (c.p.tauR - 0.5)) + (1.0 - phi) / (2.0 * (c.p.tauB - 0.5)));
        if (c.p.relax == 0) {
            // SRT: the reference collides the two colours separately with the same tau (calRKCollision1GPU2DSRTNew,
            // 1125-1163); their sum is this relaxation of the total population (f_eq is linear in rho)
#pragma unroll
            for (int q = 0; q < L::Q; ++q) d[q] = 1.0 / tau * d[q];
        } else {
            L::to_moments(d, m);
            scale_moments(m, 1.0 / tau);
            L::from_moments(m, d);
        }
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            double eF = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) eF += L::c(q, a) * c.p.bf[a];
            fT[q] = -d[q] + (q == 0 ? 0.0 : 3.0 * L::w(q)) * eF + fT[q];
        }
        double G[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            const int64_t n = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
            const double pk = (c.cls[n] & CLS_FLUID) ? c.phi[n] : c.p.solid_phi;
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) G[a] = add_rn(G[a], mul_rn(3.0 * L::w(q) * L::c(q, a), pk));
        }
        double g2 = mul_rn(G[0], G[0]);
#pragma unroll
        for (int a = 1; a < L::D; ++a) g2 = add_rn(g2, mul_rn(G[a], G[a]));
        const double gn = sqrt(g2);
        const double A = c.p.Ak, kR = rR / rho, kB = rB / rho, amp = c.p.beta * (rR * rB) / (rho * rho);
#pragma unroll
        for (int q = 0; q < L::Q; ++q) {
            double eg = 0.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a)
                if (L::c(q, a) != 0) eg += L::c(q, a) * G[a];
            double f = fT[q];
            if (g2 != 0.0) f += A * gn * (L::w(q) * (eg * eg) / g2 - (q == 0 ? L::w(0) - 2.0 / 3.0 : L::w(q)));
            const double cosT = (q != 0 && gn != 0.0) ? eg / (L::enorm(q) * gn) : 0.0;
            const double a_ = amp * L::w(q) * cosT;
            c.fC[0][q * V + id] = kR * f + a_;
            c.fC[1][q * V + id] = kB * f - a_;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) c.G[a * V + id] = G[a];
    }
};

// sum of a density over the owned void nodes is done on the host side of the ABI from a download
// (mass check only; never on the timed path)

}  // namespace lbm
