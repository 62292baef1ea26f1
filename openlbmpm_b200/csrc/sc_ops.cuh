// sc_ops.cuh -- multi-component Shan-Chen models on the dense grid, one thread per node, written once for any
// lattice (D2Q9: the reference's; D3Q19: `ShanChenD3Q19`, which main.py:17,73-77 names but the reference does not
// ship -- the direct generalisation fixed in oracle/sc_dense.py, isotropy 4):
//   * original Shan-Chen      (ShanChenD2Q9.runOptimizedLBM,   ShanChenD2Q9.py:1433-1629; kernels in
//                              ShanChen2D/OptimizedD2Q9GPU.py)
//   * explicit forcing SRT/MRT (ShanChenD2Q9.runOptimizedEFLBM, ShanChenD2Q9.py:1631-2087; kernels in
//                              ShanChen2D/ExplicitD2Q9GPU.py), isotropy 4, 8 and 10.
// The reference materialises f_eq, the force distribution and (MRT) their images under C = M^-1 S M as
// five extra [nf, N, 9] arrays and runs 7-9 kernels over them; here the state is (f, rho, F, u_eq) and
// the equilibrium / force distributions live in registers inside the collision operator.  The MRT
// relaxation is applied in moment space (M, diag(S), M^-1 hand-factored) instead of a dense 9x9 product.
#pragma once
#include "../../include/lbmpm.h"
#include "grid.cuh"

namespace lbm {

constexpr int SC_MAXC = 4;

struct SCParams {
    int nc, relax, inlet, outlet;
    int scheme;      // isotropy of the explicit force: 4 (8 neighbours), 8 (24) or 10 (36)
    double tau[SC_MAXC], G[SC_MAXC * SC_MAXC], Gs[SC_MAXC], vin[SC_MAXC], rho_out[SC_MAXC];
};

struct SCFields {
    Grid g;
    SCParams p;
    int Q, D;        // lattice of the run (9 / 2 or 19 / 3)
    double* fS;      // [nc][Q][vol] populations (EFS: the transformed populations f - fF/2)
    double* fC;      // [nc][Q][vol] post-collision
    double* rho;     // [nc][vol]
    double* F;       // [nc][D][vol]
    double* ueq;     // [D][vol]   common equilibrium velocity (EFS)
    double* uph;     // [D][vol]   physical velocity
    double* fold;    // [nc][Q][3 planes] populations of planes 0..2 before the collision (convective outlet)
    const uint8_t* cls;
    int z_in, z_in_ghost, z_out;
    LBM_HD double* f(double* base, int c, int q) const { return base + ((int64_t)c * Q + q) * g.vol; }
    LBM_HD double& Fc(int k, int a, int64_t id) const { return F[((int64_t)k * D + a) * g.vol + id]; }
};

template <class L>
LBM_HD double sc_feq(int q, double rho, const double* u) {
    double eu = 0.0, uu = 0.0;
#pragma unroll
    for (int a = 0; a < L::D; ++a) {
        if (L::c(q, a) != 0) eu += L::c(q, a) * u[a];
        uu += u[a] * u[a];
    }
    return L::w(q) * rho * (1.0 + 3.0 * eu + 9.0 / 2.0 * (eu * eu) - 3.0 / 2.0 * uu);
}
// first moment sum_q e_q f_q, summed in the order of q (the reference's f1 - f3 + f5 - f6 - f7 + f8)
template <class L>
LBM_HD void sc_momentum(const double* f, double* m) {
#pragma unroll
    for (int a = 0; a < L::D; ++a) m[a] = 0.0;
#pragma unroll
    for (int q = 1; q < L::Q; ++q)
#pragma unroll
        for (int a = 0; a < L::D; ++a) {
            if (L::c(q, a) > 0) m[a] += f[q];
            else if (L::c(q, a) < 0) m[a] -= f[q];
        }
}

// f = w rho at rest (ShanChenD2Q9.py:759-768)
template <class L>
struct ScInitOp {
    SCFields c; const double* rho_in;     // [nc][owned]
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g; const int64_t id = (int64_t)NG * g.plane + i, owned = g.plane * g.n2;
        const bool fl = c.cls[id] & CLS_FLUID;
        for (int k = 0; k < c.p.nc; ++k) {
            const double r = fl ? rho_in[k * owned + i] : 0.0;
            c.rho[k * g.vol + id] = r;
            for (int q = 0; q < L::Q; ++q) c.f(c.fS, k, q)[id] = L::w(q) * r;
            for (int a = 0; a < L::D; ++a) c.Fc(k, a, id) = 0.0;
        }
        for (int a = 0; a < L::D; ++a) c.ueq[a * g.vol + id] = c.uph[a * g.vol + id] = 0.0;
    }
};
template <class L>
struct ScUploadOp {     // AoS [node][Q] of one component -> SoA; rho = given or sum
    SCFields c; int k; const double* aos; const double* rho_in;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g; const int64_t id = (int64_t)NG * g.plane + i;
        const bool fl = c.cls[id] & CLS_FLUID;
        double s = 0.0;
        for (int q = 0; q < L::Q; ++q) {
            const double v = fl ? aos[i * L::Q + q] : 0.0;
            c.f(c.fS, k, q)[id] = v;
            s = q == 0 ? v : s + v;
        }
        c.rho[k * g.vol + id] = fl ? (rho_in ? rho_in[i] : s) : 0.0;
    }
};
template <class L>
struct ScDownloadOp {
    SCFields c; int k; double* aos;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * c.g.plane + i;
        for (int q = 0; q < L::Q; ++q) aos[i * L::Q + q] = c.f(c.fS, k, q)[id];
    }
};

// calFluidRhoGPU (OptimizedD2Q9GPU.py:84-93)
template <class L>
struct ScRhoOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * c.g.plane + i;
        if (!(c.cls[id] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            double s = c.f(c.fS, k, 0)[id];
            for (int q = 1; q < L::Q; ++q) s += c.f(c.fS, k, q)[id];
            c.rho[k * c.g.vol + id] = s;
        }
    }
};

// calPhysicalVelocity (OptimizedD2Q9GPU.py:156-175): u = sum_k (sum_q e_q f_k,q + F_k / 2) / sum_k rho_k
template <class L>
struct ScPhysicalVelocityOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g; const int64_t id = (int64_t)NG * g.plane + i;
        if (!(c.cls[id] & CLS_FLUID)) return;
        double v[3] = {0.0, 0.0, 0.0}, r = 0.0;
        for (int k = 0; k < c.p.nc; ++k) {
            double f[L::Q], m[3];
            for (int q = 1; q < L::Q; ++q) f[q] = c.f(c.fS, k, q)[id];
            sc_momentum<L>(f, m);
            for (int a = 0; a < L::D; ++a) v[a] += (m[a] + 1.0 / 2.0 * c.Fc(k, a, id));
            r += c.rho[k * g.vol + id];
        }
        for (int a = 0; a < L::D; ++a) c.uph[a * g.vol + id] = v[a] / r;
    }
};

// ---- open boundaries: planes along the flow axis (y in 2-D, z in 3-D; inlet on top, outlet at the bottom) -------
// Written for any lattice: the unknown populations of a plane are those pointing into the domain, s0 / s1 the sums of
// the in-plane / outward populations, N_t the in-plane transverse momentum that Zou-He redistributes over the diagonal
// unknowns.  For D2Q9 these are the reference's formulas term by term (f4 = f2 - 2/3 rho v, f7 = f5 + (f1 - f3)/2 -
// rho v / 6, ...); for D3Q19 (no reference code) their Hecht-Harting generalisation, as for the colour gradient.
// Items: the n0 * n1 nodes of one plane.
template <class L>
LBM_HD void sc_plane_sums(const SCFields& c, int k, int64_t id, int sign, double* fl, double* s0, double* s1, double* N) {
    bool first0 = true, first1 = true;
    N[0] = N[1] = 0.0; *s0 = *s1 = 0.0;
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
        fl[q] = c.f(c.fS, k, q)[id];
        if (L::d2(q) == 0) {
            *s0 = first0 ? fl[q] : *s0 + fl[q]; first0 = false;
            if (L::d0(q) != 0) N[0] += L::d0(q) * fl[q];
            if (L::d1(q) != 0) N[1] += L::d1(q) * fl[q];
        } else if (L::d2(q) == sign) {
            *s1 = first1 ? fl[q] : *s1 + fl[q]; first1 = false;
        }
    }
}
// constantVelocityZouHeBoundaryHigher (OptimizedD2Q9GPU.py:839-861): per-component Zou-He velocity on plane z_in
template <class L>
struct ScInletVelocityOp {
    SCFields c;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g; const int64_t id = (int64_t)(c.z_in + NG) * g.plane + r;
        if (!(c.cls[id] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            const double v = c.p.vin[k];
            double fl[L::Q], s0, sp, N[2];
            sc_plane_sums<L>(c, k, id, +1, fl, &s0, &sp, N);
            const double rho = (s0 + 2.0 * sp) / (1.0 + v);
            c.rho[k * g.vol + id] = rho;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                if (L::d2(q) != -1) continue;
                c.f(c.fS, k, q)[id] = fl[L::opp(q)] + 1.0 / 2.0 * (-(L::d0(q) * N[0] + L::d1(q) * N[1])) - 6.0 * L::w(q) * rho * v;
            }
        }
    }
};
// constantPressureZouHeBoundaryLower (OptimizedD2Q9GPU.py:555-584) on plane z_out.  The reference ignores its
// densityL argument and uses the hard-coded densities [1.0, 0.02]; they arrive here as p.rho_out.
template <class L>
struct ScOutletPressureOp {
    SCFields c;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g; const int64_t id = (int64_t)(c.z_out + NG) * g.plane + r;
        if (!(c.cls[id] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            const double d = c.p.rho_out[k];
            double fl[L::Q], s0, sm, N[2];
            sc_plane_sums<L>(c, k, id, -1, fl, &s0, &sm, N);
            const double vy = 1.0 - (s0 + 2.0 * sm) / d;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                if (L::d2(q) != 1) continue;
                c.f(c.fS, k, q)[id] = fl[L::opp(q)] + 1.0 / 2.0 * (-(L::d0(q) * N[0] + L::d1(q) * N[1])) + 6.0 * L::w(q) * d * vy;
            }
            c.rho[k * g.vol + id] = d;
        }
    }
};
// plane copies with rho = sum: ghostPointsConstantVelocityInlet (710-736), ghostPointsConstantPressureOutlet
// (743-768), convectiveOutletGPU / Ghost2 / Ghost3 (960-1036)
template <class L>
struct ScRowCopyOp {
    SCFields c; int z_dst, z_src;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t d = (int64_t)(z_dst + NG) * g.plane + r, s = (int64_t)(z_src + NG) * g.plane + r;
        if (!(c.cls[d] & CLS_FLUID) || !(c.cls[s] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            double acc = 0.0;
            for (int q = 0; q < L::Q; ++q) {
                const double v = c.f(c.fS, k, q)[s];
                c.f(c.fS, k, q)[d] = v;
                acc = q == 0 ? v : acc + v;
            }
            c.rho[k * g.vol + d] = acc;
        }
    }
};
// savePDFLastStep (OptimizedD2Q9GPU.py:70-78), restricted to the planes the convective outlet reads (0..2)
template <class L>
struct ScSaveRowsOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {      // i over 3 planes
        const Grid& g = c.g; const int64_t id = (int64_t)NG * g.plane + i;
        for (int k = 0; k < c.p.nc; ++k)
            for (int q = 0; q < L::Q; ++q) c.fold[((int64_t)k * L::Q + q) * 3 * g.plane + i] = c.f(c.fS, k, q)[id];
    }
};
// convectiveOutletEachGPU / Each2 / Each3 (OptimizedD2Q9GPU.py:1044-1119): plane z <- (f_old + |u_up(plane 3)| f(plane z+1)) / (1 + |u_up|)
template <class L>
struct ScConvectiveEachOp {
    SCFields c; int z;
    LBM_HD void operator()(int64_t r) const {
        const Grid& g = c.g;
        const int64_t d = (int64_t)(z + NG) * g.plane + r, s = (int64_t)(z + 1 + NG) * g.plane + r;
        const int64_t r3 = (int64_t)(3 + NG) * g.plane + r;
        if (!(c.cls[d] & CLS_FLUID)) return;
        const double v = fabs(c.uph[(int64_t)(L::D - 1) * g.vol + r3]);
        for (int k = 0; k < c.p.nc; ++k) {
            double acc = 0.0;
            for (int q = 0; q < L::Q; ++q) {
                const double fo = c.fold[((int64_t)k * L::Q + q) * 3 * g.plane + (int64_t)z * g.plane + r];
                const double val = (fo + v * c.f(c.fS, k, q)[s]) / (1.0 + v);
                c.f(c.fS, k, q)[d] = val;
                acc += val;
            }
            c.rho[k * g.vol + d] = acc;
        }
    }
};

// All boundary rows of one iteration in ONE launch (the reference: 3-6 kernels over every node; one launch per row
// operator here was 5-7 launches of a launch-bound 2-D step).  Items: 2 planes' worth of columns -- the first `plane`
// items walk the outlet rows of their column, the others its inlet rows.  Rows of one column only depend on each other,
// so one thread runs them back to back: same arithmetic, same order.  `efs`: the explicit-forcing loop
// (ShanChenD2Q9.py:1933-2014: convective-each / pressure outlet, inlet, then calFluidRhoGPU -- of which only the two
// Zou-He planes are not already sums); otherwise the original loop's pieces selected by `part`:
// 1 = inlet at the top of an iteration (+ calFluidRhoGPU on its plane), 2 = convective outlet copies after the streaming
// (3 = both: the two-pass form treats the rows of iteration k's end and iteration k + 1's top between its passes).
template <class L>
struct ScOpenRowsOp {
    SCFields c; int efs, part, do_in, do_out, rho_after_inlet;
    LBM_HD void inlet_column(int64_t r) const {
        ScInletVelocityOp<L>{c}(r);
        for (int zr = c.z_in; zr < c.z_in_ghost; ++zr) ScRowCopyOp<L>{c, zr + 1, zr}(r);
        if (rho_after_inlet) ScRhoOp<L>{c}((int64_t)c.z_in * c.g.plane + r);
    }
    LBM_HD void operator()(int64_t i) const {
        const int64_t plane = c.g.plane;
        if (i < plane) {
            if (!do_out) return;
            if (efs) {
                if (c.p.outlet == LBM_OUTLET_CONVECTIVE) {
                    ScPhysicalVelocityOp<L>{c}(3 * plane + i);          // |u_up| of plane 3 is all the convective rows read
                    ScConvectiveEachOp<L>{c, 2}(i); ScConvectiveEachOp<L>{c, 1}(i); ScConvectiveEachOp<L>{c, 0}(i);
                } else if (c.p.outlet == LBM_OUTLET_PRESSURE) {
                    ScOutletPressureOp<L>{c}(i);
                    for (int zr = c.z_out; zr > 0; --zr) ScRowCopyOp<L>{c, zr - 1, zr}(i);
                    ScRhoOp<L>{c}((int64_t)c.z_out * plane + i);
                }
            } else if ((part & 2) && c.p.outlet == LBM_OUTLET_CONVECTIVE) {
                ScRowCopyOp<L>{c, 2, 3}(i); ScRowCopyOp<L>{c, 1, 2}(i); ScRowCopyOp<L>{c, 0, 1}(i);
            }
        } else {
            if (!do_in || c.p.inlet != LBM_INLET_VELOCITY) return;
            if (efs || (part & 1)) inlet_column(i - plane);
        }
    }
};

// pull streaming with half-way bounce back + densities (calStreaming1GPU/2GPU 450-548, calFluidRhoGPU)
template <class L>
struct ScStreamOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z);
        if (!(c.cls[id] & CLS_FLUID)) return;
        int64_t src[L::Q]; bool fl[L::Q];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            src[q] = g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q));
            fl[q] = c.cls[src[q]] & CLS_FLUID;
        }
        for (int k = 0; k < c.p.nc; ++k) {
            double acc = c.f(c.fC, k, 0)[id];
            c.f(c.fS, k, 0)[id] = acc;
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double v = fl[q] ? c.f(c.fC, k, q)[src[q]] : c.f(c.fC, k, L::opp(q))[id];
                c.f(c.fS, k, q)[id] = v;
                acc += v;
            }
            c.rho[k * g.vol + id] = acc;
        }
    }
};

// interactionCollisionProcess (OptimizedD2Q9GPU.py:1274-1446): common velocity u', Shan-Chen force with
// psi = rho (fluid-fluid through the neighbours, fluid-solid with the lattice weights 1/9, 1/36 | 1/18, 1/36),
// SRT collision towards f_eq(rho, u' + tau F / rho).  fS -> fC, writes F.
template <class L>
struct ScCollideOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const int nc = c.p.nc;
        double vt[3] = {0.0, 0.0, 0.0}, rt = 0.0;
        for (int k = 0; k < nc; ++k) {
            double f[L::Q], m[3];
#pragma unroll
            for (int q = 1; q < L::Q; ++q) f[q] = c.f(c.fS, k, q)[id];
            sc_momentum<L>(f, m);
#pragma unroll
            for (int a = 0; a < L::D; ++a) vt[a] += m[a] / c.p.tau[k];
            rt += c.rho[k * V + id] / c.p.tau[k];
        }
        double up[3];
#pragma unroll
        for (int a = 0; a < L::D; ++a) up[a] = vt[a] / rt;
        int64_t nb[L::Q]; bool fl[L::Q];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            nb[q] = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
            fl[q] = c.cls[nb[q]] & CLS_FLUID;
        }
        for (int k = 0; k < nc; ++k) {
            const double psi = c.rho[k * V + id];
            double F[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double wI = L::w(q);
                if (fl[q]) {
                    for (int j = 0; j < nc; ++j) {
                        const double t = -wI * c.p.G[k * SC_MAXC + j] * psi * c.rho[j * V + nb[q]];
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
#pragma unroll
            for (int q = 0; q < L::Q; ++q) {
                double eu = 0.0;
#pragma unroll
                for (int a = 0; a < L::D; ++a)
                    if (L::c(q, a) != 0) eu += L::c(q, a) * u[a];
                c.f(c.fC, k, q)[id] = (1.0 - 1.0 / tau) * c.f(c.fS, k, q)[id] +
                                      L::w(q) * psi / tau * (1.0 + 3.0 * eu + 4.5 * (eu * eu) - 1.5 * uu);
            }
        }
    }
};

// Higher-isotropy explicit force: calExplicit8thOrderScheme (ExplicitD2Q9GPU.py:627-953, 24 neighbours) and
// calExplicit10thOrderScheme (957-1372, 36 neighbours).  Neighbour slots in the order of fillNeighboringNodesISO8/10
// (392-592).  A far neighbour only contributes when it is fluid AND "visible": the nearer node(s) on the way are
// fluid (the gates below, read off the kernels' if-conditions).  Solid neighbours act through the first 8 slots
// only, with the hard-coded weights 1/9 and 1/36.  The 8th-order kernel uses psi_j(x+e) - psi_j(x), the
// 10th-order kernel plain psi_j(x+e).
LBM_HD void efs_force_iso(const SCFields& c, int x, int y, int z, int64_t id, int k, double* fx_out, double* fy_out) {
    constexpr int OX[36] = {1, 0, -1, 0, 1, -1, -1, 1, 2, 0, -2, 0, 2, -2, -2, 2, 2, 1, -1, -2, -2, -1, 1, 2,
                            3, 0, -3, 0, 3, 1, -1, -3, -3, -1, 1, 3};
    constexpr int OY[36] = {0, 1, 0, -1, 1, 1, -1, -1, 0, 2, 0, -2, 2, 2, -2, -2, 1, 2, 2, 1, -1, -2, -2, -1,
                            0, 3, 0, -3, 1, 3, 3, 1, -1, -3, -3, -1};
    constexpr int KA[8] = {0, 1, 1, 2, 2, 3, 3, 0}, KB[8] = {4, 4, 5, 5, 6, 6, 7, 7};    // knight moves: axis / diagonal gate
    const Grid& g = c.g; const int64_t V = g.vol;
    const int ns = c.p.scheme == 8 ? 24 : 36;
    int64_t nb[36]; bool fl[36];
    for (int s = 0; s < ns; ++s) {
        nb[s] = g.nb(x, y, z, OX[s], 0, OY[s]);
        fl[s] = c.cls[nb[s]] & CLS_FLUID;
    }
    const double psi = c.rho[k * V + id];
    double fx = 0.0, fy = 0.0;
    for (int s = 0; s < ns; ++s) {
        bool gate = true;
        double w;
        if (c.p.scheme == 8) w = s < 4 ? 4.0 / 21.0 : s < 8 ? 4.0 / 45.0 : s < 12 ? 1.0 / 60.0 : s < 16 ? 1.0 / 5040.0 : 2.0 / 315.0;
        else w = s < 4 ? 262.0 / 1785.0 : s < 8 ? 93.0 / 1190.0 : s < 12 ? 7.0 / 340.0 : s < 16 ? 9.0 / 9520.0 :
                 s < 24 ? 6.0 / 595.0 : s < 28 ? 2.0 / 5355.0 : 1.0 / 7140.0;
        if (s >= 8 && s < 16) gate = fl[s - 8];
        else if (s >= 16 && s < 24) gate = fl[KA[s - 16]] || fl[KB[s - 16]];
        else if (s >= 24 && s < 28) gate = fl[s - 24] && fl[s - 16];
        else if (s >= 28) {
            const int t = s - 28;                       // same quadrant pairing as the knight moves
            const int axis = KA[t], diag = KB[t];
            gate = (fl[axis] && fl[axis + 8]) || (fl[diag] && fl[16 + t]);
        }
        if (fl[s] && gate) {
            for (int j = 0; j < c.p.nc; ++j) {
                const double d = c.p.scheme == 8 ? c.rho[j * V + nb[s]] - c.rho[j * V + id] : c.rho[j * V + nb[s]];
                const double t = -6.0 * w * c.p.G[k * SC_MAXC + j] * psi * d;
                if (OX[s] != 0) fx += OX[s] * t;
                if (OY[s] != 0) fy += OY[s] * t;
            }
        } else if (s < 8 && !fl[s]) {
            const double t = -(s < 4 ? 1.0 / 9.0 : 1.0 / 36.0) * c.p.Gs[k] * psi;
            if (OX[s] != 0) fx += t * OX[s];
            if (OY[s] != 0) fy += t * OY[s];
        }
    }
    *fx_out = fx; *fy_out = fy;
}

// calExplicit4thOrderScheme (ExplicitD2Q9GPU.py:51-217) + calEquilibriumVEFGPU (340-363, SRT) /
// transformEquilibriumVelocity (1426-1449, MRT: weights s_0 = 1 instead of 1/tau).  Writes F and u_eq.
// Interaction weights 3 w_q (D2Q9: 1/3, 1/12 -- ShanChenD2Q9.py:1675; D3Q19: 1/6, 1/12).
template <class L>
struct EfsForceOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const Grid& g = c.g;
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z), V = g.vol;
        if (!(c.cls[id] & CLS_FLUID)) return;
        const int nc = c.p.nc;
        int64_t nb[L::Q]; bool fl[L::Q];
#pragma unroll
        for (int q = 1; q < L::Q; ++q) {
            nb[q] = g.nb(x, y, z, L::d0(q), L::d1(q), L::d2(q));
            fl[q] = c.cls[nb[q]] & CLS_FLUID;
        }
        // physical velocity of the output point, u = sum_k (sum_q e_q f_k,q + F_k / 2) / sum_k rho_k with the force of the
        // PREVIOUS evaluation (calPhysicalVelocity, OptimizedD2Q9GPU.py:156-175, launched by the reference right before
        // this kernel, ShanChenD2Q9.py:2016-2027): same inputs, so it rides along instead of costing a pass of its own
        {
            double v[3] = {0.0, 0.0, 0.0}, r = 0.0;
            for (int k = 0; k < nc; ++k) {
                double f[L::Q], m[3];
#pragma unroll
                for (int q = 1; q < L::Q; ++q) f[q] = c.f(c.fS, k, q)[id];
                sc_momentum<L>(f, m);
#pragma unroll
                for (int a = 0; a < L::D; ++a) v[a] += (m[a] + 1.0 / 2.0 * c.Fc(k, a, id));
                r += c.rho[k * V + id];
            }
#pragma unroll
            for (int a = 0; a < L::D; ++a) c.uph[a * V + id] = v[a] / r;
        }
        double mt[3] = {0.0, 0.0, 0.0}, rt = 0.0;
        for (int k = 0; k < nc; ++k) {
            const double psi = c.rho[k * V + id];
            double gr[3] = {0.0, 0.0, 0.0}, sl[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 1; q < L::Q; ++q) {
                const double wI = L::D == 2 ? (q < 5 ? 1.0 / 3.0 : 1.0 / 12.0) : (q < 7 ? 1.0 / 6.0 : 1.0 / 12.0);
                if (fl[q]) {
                    for (int j = 0; j < nc; ++j) {
                        const double t = wI * (c.rho[j * V + nb[q]] - c.rho[j * V + id]);
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
            double F[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < L::D; ++a) F[a] = -6.0 * psi * gr[a] + sl[a];
            if (L::D == 2 && c.p.scheme != 4) efs_force_iso(c, x, y, z, id, k, &F[0], &F[1]);
            double f[L::Q], e[3];
#pragma unroll
            for (int q = 1; q < L::Q; ++q) f[q] = c.f(c.fS, k, q)[id];
            sc_momentum<L>(f, e);
            const double wgt = c.p.relax == 0 ? 1.0 / c.p.tau[k] : 1.0;
#pragma unroll
            for (int a = 0; a < L::D; ++a) {
                c.Fc(k, a, id) = F[a];
                const double ea = e[a] + 1.0 / 2.0 * F[a];
                if (c.p.relax == 0) mt[a] += ea / c.p.tau[k]; else mt[a] += ea * wgt;
            }
            if (c.p.relax == 0) rt = rt + psi / c.p.tau[k]; else rt += psi * wgt;
        }
#pragma unroll
        for (int a = 0; a < L::D; ++a) c.ueq[a * V + id] = mt[a] / rt;
    }
};

// equilibrium and force distributions of component k at a node (calEquilibriumFuncEFGPU 227-248,
// calForceDistrGPU 255-272)
template <class L>
LBM_HD void efs_feq_ff_at(double r, const double* u, const double* F, double* feq, double* ff) {
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
        feq[q] = sc_feq<L>(q, r, u);
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < L::D; ++a) s += F[a] * (L::c(q, a) - u[a]);
        ff[q] = s * feq[q] / (1.0 / 3.0 * r);
    }
}
template <class L>
LBM_HD void efs_feq_ff(const SCFields& c, int k, int64_t id, double* feq, double* ff) {
    const int64_t V = c.g.vol;
    const double r = c.rho[k * V + id];
    double u[3] = {0.0, 0.0, 0.0}, F[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < L::D; ++a) { u[a] = c.ueq[a * V + id]; F[a] = c.Fc(k, a, id); }
    efs_feq_ff_at<L>(r, u, F, feq, ff);
}
// transformPDFGPU (ExplicitD2Q9GPU.py:278-287), once before the first iteration: f <- f - fF / 2
template <class L>
struct EfsTransformOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * c.g.plane + i;
        if (!(c.cls[id] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            double feq[L::Q], ff[L::Q];
            efs_feq_ff<L>(c, k, id, feq, ff);
#pragma unroll
            for (int q = 0; q < L::Q; ++q) c.f(c.fS, k, q)[id] = c.f(c.fS, k, q)[id] - 1.0 / 2.0 * ff[q];
        }
    }
};
// relaxation rates of the non-conserved moments: D2Q9 s = [1, 0.6, 1.5, 1, 1.2, 1, 1.2, 1/tau, 1/tau] for the first
// two components, 1 (and 1/tau on the stress moments) for further ones (ShanChenD2Q9.py:99-106, 484-496); D3Q19:
// the d'Humieres rates of the colour-gradient specification (1.19, 1.4, 1.2, 1.4, 1.98), 1/tau on the five
// stress moments (oracle/sc_dense.py)
template <class L>
LBM_HD void efs_scale_moments(double* m, double st, bool tuned) {
    if (L::Q == 9) {
        if (tuned) { m[1] *= 0.6; m[2] *= 1.5; m[4] *= 1.2; m[6] *= 1.2; }
        m[7] *= st; m[8] *= st;
    } else {
        if (tuned) {
            m[1] *= 1.19; m[2] *= 1.4; m[4] *= 1.2; m[6] *= 1.2; m[8] *= 1.2; m[10] *= 1.4; m[12] *= 1.4;
            m[16] *= 1.98; m[17] *= 1.98; m[18] *= 1.98;
        }
        m[9] *= st; m[11] *= st; m[13] *= st; m[14] *= st; m[15] *= st;
    }
}
// calCollisionEXGPU (294-304) / transfromForceTerm + transformPDFandEquil + calAfterCollisionMRT
// (1379-1469): f <- f + C (feq - f - fF/2) + fF with C = 1/tau (SRT) or M^-1 diag(s) M (MRT)
template <class L>
struct EfsCollideOp {
    SCFields c;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)NG * c.g.plane + i;
        if (!(c.cls[id] & CLS_FLUID)) return;
        for (int k = 0; k < c.p.nc; ++k) {
            double feq[L::Q], ff[L::Q], f[L::Q], d[L::Q];
            efs_feq_ff<L>(c, k, id, feq, ff);
#pragma unroll
            for (int q = 0; q < L::Q; ++q) { f[q] = c.f(c.fS, k, q)[id]; d[q] = feq[q] - f[q] - 1.0 / 2.0 * ff[q]; }
            if (c.p.relax == 0) {
#pragma unroll
                for (int q = 0; q < L::Q; ++q) c.f(c.fC, k, q)[id] = f[q] + 1.0 / c.p.tau[k] * d[q] + 1.0 * ff[q];
            } else {
                double m[L::Q], cd[L::Q];
                L::to_moments(d, m);
                efs_scale_moments<L>(m, 1.0 / c.p.tau[k], k < 2);
                L::from_moments(m, cd);
#pragma unroll
                for (int q = 0; q < L::Q; ++q) c.f(c.fC, k, q)[id] = f[q] + cd[q] + 1.0 * ff[q];
            }
        }
    }
};

}  // namespace lbm
