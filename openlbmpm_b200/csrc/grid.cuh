// grid.cuh -- the dense masked lattice of one slab and the geometry operators.
//
// Every field is a dense array over the padded slab [n2 + 2*NG][n1][n0], x (axis 0) fastest; axis 2 is
// the flow / slab axis (z in 3-D, the reference's y in 2-D, where n1 = 1).  NG ghost planes on each
// side of axis 2 hold the periodic images (one GPU) or the neighbouring ranks' planes (slab
// decomposition); axes 0 and 1 wrap inside the kernels.  The reference instead keeps a compact
// fluid-node list with an int64 neighbour table (RKD2Q9.py:657-736); that structure is only
// exported on request (lbm_export_indexing), never used on the hot path.
#pragma once
#include "lattice.cuh"

namespace lbm {

constexpr int NG = 3;   // phi needs radius 3 when wetting solids are present (phi_s -> G -> n -> K)

enum : uint8_t { CLS_FLUID = 1, CLS_WET = 2, CLS_NEAR = 4 };

struct Grid {
    int n0, n1, n2;
    int64_t plane;   // n0 * n1
    int64_t vol;     // plane * (n2 + 2 NG)
    int wrap2;       // 1: a single slab whose one-thread-per-node operators wrap axis 2 by index arithmetic, so that the
                     //    per-step ghost-plane copies (2 launches of a launch-bound 2-D step) are not needed at all
    LBM_HD int64_t at(int x, int y, int z) const { return (int64_t)(z + NG) * plane + (int64_t)y * n0 + x; }
    // neighbour (x+dx, y+dy, z+dz): periodic in axes 0 and 1; along axis 2 the ghost planes, or the periodic image itself
    LBM_HD int64_t nb(int x, int y, int z, int dx, int dy, int dz) const {
        int xn = x + dx, yn = y + dy, zn = z + dz;
        if (xn < 0) xn += n0; else if (xn >= n0) xn -= n0;
        if (yn < 0) yn += n1; else if (yn >= n1) yn -= n1;
        if (wrap2) { if (zn < 0) zn += n2; else if (zn >= n2) zn -= n2; }
        return at(xn, yn, zn);
    }
    // the coordinates nb() addresses
    LBM_HD void nb_coords(int x, int y, int z, int dx, int dy, int dz, int& xn, int& yn, int& zn) const {
        xn = x + dx; yn = y + dy; zn = z + dz;
        if (xn < 0) xn += n0; else if (xn >= n0) xn -= n0;
        if (yn < 0) yn += n1; else if (yn >= n1) yn -= n1;
        if (wrap2) { if (zn < 0) zn += n2; else if (zn >= n2) zn -= n2; }
    }
    // item i of a launch over planes [-ext, n2 + ext)
    LBM_HD void decode(int64_t i, int ext, int& x, int& y, int& z) const {
        int r;
        if ((uint64_t)(i | plane) >> 31 == 0) {     // every lattice below 2^31 nodes: one 32-bit division instead of an emulated 64-bit one
            const uint32_t ii = (uint32_t)i, p = (uint32_t)plane, zz = ii / p;
            z = (int)zz - ext;
            r = (int)(ii - zz * p);
        } else {
            z = (int)(i / plane) - ext;
            r = (int)(i % plane);
        }
        y = n1 == 1 ? 0 : r / n0;
        x = r - y * n0;
    }
    LBM_HD int64_t count(int ext) const { return plane * (int64_t)(n2 + 2 * ext); }
};

// The 27 neighbour ids of a node from three small tables (wrapped x, y * n0, plane base of z): one select per wrap and two
// additions per neighbour instead of the compare / branch chain of Grid::nb per direction.  Same ids as Grid::nb.
struct NbTable {
    int64_t xs[3], ys[3], zs[3];
    LBM_HD NbTable(const Grid& g, int x, int y, int z) {
        xs[0] = x == 0 ? g.n0 - 1 : x - 1; xs[1] = x; xs[2] = x == g.n0 - 1 ? 0 : x + 1;
        const int ym = y == 0 ? g.n1 - 1 : y - 1, yp = y == g.n1 - 1 ? 0 : y + 1;
        ys[0] = (int64_t)ym * g.n0; ys[1] = (int64_t)y * g.n0; ys[2] = (int64_t)yp * g.n0;
        int zm = z - 1, zp = z + 1;
        if (g.wrap2) { zm = zm < 0 ? zm + g.n2 : zm; zp = zp >= g.n2 ? zp - g.n2 : zp; }
        zs[0] = (int64_t)(zm + NG) * g.plane; zs[1] = (int64_t)(z + NG) * g.plane; zs[2] = (int64_t)(zp + NG) * g.plane;
    }
    LBM_HD int64_t operator()(int dx, int dy, int dz) const { return zs[dz + 1] + ys[dy + 1] + xs[dx + 1]; }
};

// an operator restricted to a range of planes: item i of the launch is item i + off of the operator
template <class Op>
struct PlaneRangeOp {
    Op op; int64_t off;
    LBM_HD void operator()(int64_t i) const { op(i + off); }
};

// ... and to two disjoint ranges in one launch (the outlet planes and the inlet planes of an open box)
template <class Op>
struct TwoRangeOp {
    Op op; int64_t off0, cnt0, off1;
    LBM_HD void operator()(int64_t i) const { op(i < cnt0 ? i + off0 : i - cnt0 + off1); }
};

// Fill the ghost planes of `narr` arrays (stride `stride` elements apart) from the slab's own opposite
// planes: periodic wrap on one GPU.  item = (array, side, ghost plane j, node in plane)
template <class T>
struct GhostWrapOp {
    Grid g; T* base; int64_t stride; int narr; int gp;
    LBM_HD void operator()(int64_t i) const {
        const int64_t r = i % g.plane; int64_t q = i / g.plane;
        const int j = (int)(q % gp); q /= gp;
        const int side = (int)(q % 2); const int a = (int)(q / 2);
        T* f = base + a * stride;
        if (side == 0) f[(int64_t)(NG - 1 - j) * g.plane + r] = f[(int64_t)(NG + g.n2 - 1 - j) * g.plane + r];
        else f[(int64_t)(NG + g.n2 + j) * g.plane + r] = f[(int64_t)(NG + j) * g.plane + r];
    }
    int64_t items() const { return (int64_t)narr * 2 * gp * g.plane; }
};

// One-sided ghost-plane exchange: store my boundary planes of `narr` arrays into the ghost planes of the neighbour slabs
// (`up` / `down` are the neighbours' images of `base`).  item = (array, side, plane j, node in plane); dirs[a]: 0 both
// ways, +1 only upwards, -1 only downwards, 2 none (comm.cu::ring_exchange has the same filter).
struct PeerPushOp {
    Grid g; const double* base; double* up; double* down; int64_t stride; int narr; int gp; int8_t dirs[48];
    LBM_HD void operator()(int64_t i) const {
        const int64_t r = i % g.plane; int64_t q = i / g.plane;
        const int j = (int)(q % gp); q /= gp;
        const int side = (int)(q % 2); const int a = (int)(q / 2);
        const int dir = dirs[a];
        const double* f = base + a * stride;
        if (side == 0) {            // my top planes -> low ghost of the slab above
            if (dir == 0 || dir == 1) up[a * stride + (int64_t)(NG - gp + j) * g.plane + r] = f[(int64_t)(NG + g.n2 - gp + j) * g.plane + r];
        } else {                    // my bottom planes -> high ghost of the slab below
            if (dir == 0 || dir == -1) down[a * stride + (int64_t)(NG + g.n2 + j) * g.plane + r] = f[(int64_t)(NG + j) * g.plane + r];
        }
    }
    int64_t items() const { return (int64_t)narr * 2 * gp * g.plane; }
};

// Node classes from the void mask (1 = void): the dense equivalent of optimizeFluidandSolidArray's
// wetting-solid marking (RKD2Q9.py:677-689) and of sortOutFluidNodesToSolid (RKD2Q9.py:741-760).
template <int D>
struct ClassifyOp {
    Grid g; const uint8_t* dom; uint8_t* cls;
    LBM_HD void operator()(int64_t i) const {
        int x, y, z; g.decode(i, NG, x, y, z);
        if (z < -2 || z >= g.n2 + 2) {      // outermost ghost planes: only the fluid bit is known (and needed)
            const int64_t id = g.at(x, y, z);
            cls[id] = dom[id] ? (uint8_t)CLS_FLUID : (uint8_t)0;
            return;
        }
        int nfl = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = (D == 3 ? -1 : 0); dy <= (D == 3 ? 1 : 0); ++dy)
                for (int dx = -1; dx <= 1; ++dx) nfl += dom[g.nb(x, y, z, dx, dy, dz)] ? 1 : 0;
        const int full = D == 3 ? 27 : 9;
        const int64_t id = g.at(x, y, z);
        cls[id] = dom[id] ? (uint8_t)(CLS_FLUID | (nfl < full ? CLS_NEAR : 0)) : (uint8_t)(nfl > 0 ? CLS_WET : 0);
    }
};

// Unit normal of the solid surface seen from a fluid node next to solid: calVectorNormaltoSolid
// (RKD2Q9.py:768-892), n_s = sum over solid cells c in the 5^D box of w(|c|^2) c, normalised.
// 2-D weights are the reference's (RKD2Q9.py:806-880); 3-D uses the 8th-order isotropic set.
template <int D>
struct SolidNormalOp {
    Grid g; const uint8_t* dom; const uint8_t* cls; double* ns;   // ns: [3][vol]
    LBM_HD static double weight(int c2) {
        if (D == 2) {
            switch (c2) { case 1: return 4.0 / 21.0; case 2: return 4.0 / 45.0; case 4: return 1.0 / 60.0;
                          case 5: return 2.0 / 315.0; case 8: return 1.0 / 5040.0; default: return 0.0; }
        }
        switch (c2) { case 1: return 4.0 / 45.0; case 2: return 1.0 / 21.0; case 3: return 2.0 / 105.0;
                      case 4: return 5.0 / 504.0; case 5: return 1.0 / 315.0; case 6: return 1.0 / 630.0;
                      case 8: return 1.0 / 5040.0; default: return 0.0; }
    }
    LBM_HD void operator()(int64_t i) const {
        int x, y, z; g.decode(i, 1, x, y, z);
        const int64_t id = g.at(x, y, z);
        if (!(cls[id] & CLS_NEAR)) return;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;    // along array axes 0, 1, 2
        for (int dz = -2; dz <= 2; ++dz)
            for (int dy = (D == 3 ? -2 : 0); dy <= (D == 3 ? 2 : 0); ++dy)
                for (int dx = -2; dx <= 2; ++dx) {
                    const double w = weight(dx * dx + dy * dy + dz * dz);
                    if (w == 0.0) continue;
                    int xn = x + dx, yn = y + dy;     // |d| = 2 may wrap twice on tiny grids
                    xn = ((xn % g.n0) + g.n0) % g.n0; yn = ((yn % g.n1) + g.n1) % g.n1;
                    if (!dom[g.at(xn, yn, z + dz)]) { s0 += w * dx; s1 += w * dy; s2 += w * dz; }
                }
        const double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
        // physical components: 2-D (x, y) = axes (0, 2); 3-D (x, y, z) = axes (0, 1, 2)
        if (D == 2) { ns[id] = s0 / nrm; ns[g.vol + id] = s2 / nrm; }
        else { ns[id] = s0 / nrm; ns[g.vol + id] = s1 / nrm; ns[2 * g.vol + id] = s2 / nrm; }
    }
};

// Pull mask of an owned node for the tiled kernels: bit 0 = the node is fluid, bit q (1..Q-1) = the upstream node
// x - e_q is fluid (the population is pulled; otherwise it bounces back from the node's own opposite direction),
// bit 31 = fluid node next to solid.  One coalesced 4-byte load per node replaces Q byte gathers of the class
// array, and -- unlike those -- it can be requested one plane ahead of the populations it steers.
constexpr uint32_t PULL_NEAR = 0x80000000u;
template <class L>
struct PullMaskOp {
    Grid g; const uint8_t* cls; uint32_t* pull;
    LBM_HD void operator()(int64_t i) const {
        int x, y, z; g.decode(i, 0, x, y, z);
        const int64_t id = g.at(x, y, z);
        uint32_t m = 0;
        if (cls[id] & CLS_FLUID) {
            m = 1u | ((cls[id] & CLS_NEAR) ? PULL_NEAR : 0u);
#pragma unroll
            for (int q = 1; q < L::Q; ++q)
                if (cls[g.nb(x, y, z, -L::d0(q), -L::d1(q), -L::d2(q))] & CLS_FLUID) m |= 1u << q;
        }
        pull[id] = m;
    }
};

// ---- compact list of the wetting-solid nodes of planes [-2, n2 + 2), grouped by plane ------------------------------
// (the tiled kernels stage phi with bulk copies, so the colour a wetting solid shows has to be IN the phi array; evaluating it
// over a list of the few per cent of nodes that are wetting solids replaces a pass over the whole lattice)
LBM_HD int64_t lbm_atomic_inc(int64_t* p) {
#ifdef __CUDA_ARCH__
    return (int64_t)atomicAdd((unsigned long long*)p, 1ull);
#else
    return (*p)++;
#endif
}
struct WetCountOp {      // counts[plane + 2] += 1 for every wetting solid; item = node of planes [-2, n2 + 2)
    Grid g; const uint8_t* cls; int64_t* counts;
    LBM_HD void operator()(int64_t i) const {
        if (cls[(int64_t)(NG - 2) * g.plane + i] & CLS_WET) lbm_atomic_inc(counts + i / g.plane);
    }
};
struct WetFillOp {       // list[cursor[plane]++] = flat id (unordered inside a plane: the items are independent)
    Grid g; const uint8_t* cls; int64_t* cursor; int64_t* list;
    LBM_HD void operator()(int64_t i) const {
        const int64_t id = (int64_t)(NG - 2) * g.plane + i;
        if (cls[id] & CLS_WET) list[lbm_atomic_inc(cursor + i / g.plane)] = id;
    }
};

// ---- export of the reference's compact index structures (single slab; bit-exact contract) -------
struct FlagOp {          // flag[i] = 1 where the owned node has all bits of `mask` set, row-major order
    Grid g; const uint8_t* cls; uint8_t mask; int64_t* flag;
    LBM_HD void operator()(int64_t i) const { flag[i] = (cls[(int64_t)NG * g.plane + i] & mask) ? 1 : 0; }
};
// list[rank[i]] = i for flagged nodes; id[i] = (negate ? -2 - rank : rank)
struct CompactOp {
    const int64_t* flag; const int64_t* rank; int64_t* list; int64_t* id; int negate;
    LBM_HD void operator()(int64_t i) const {
        if (!flag[i]) return;
        list[rank[i]] = i;
        if (id) id[i] = negate ? -2 - rank[i] : rank[i];
    }
};
// neighbour table of the listed nodes: slot k <-> direction k+1, periodic in every axis
// (fillNeighboringNodes / fillNeighboringWettingNodes, AcceleratedRKGPU2D.py:14-95)
template <class L>
struct NeighbourTableOp {
    Grid g; const int64_t* list; const int64_t* id; int64_t* out;
    LBM_HD void operator()(int64_t i) const {
        const int64_t node = list[i];
        int x, y, z; g.decode(node, 0, x, y, z);
#pragma unroll
        for (int k = 1; k < L::Q; ++k) {
            int zn = z + L::d2(k);
            if (zn < 0) zn += g.n2; else if (zn >= g.n2) zn -= g.n2;
            const int64_t nbn = g.nb(x, y, zn, L::d0(k), L::d1(k), 0) - (int64_t)NG * g.plane;
            out[i * (L::Q - 1) + (k - 1)] = id[nbn];
        }
    }
};
// fluid nodes next to solid: compact id, flat id and unit normal (component-major)
struct NearSolidExportOp {
    Grid g; int D; const int64_t* list; const int64_t* id; const double* ns; int64_t n;
    int64_t* compact; int64_t* flat; double* ns_out;
    LBM_HD void operator()(int64_t i) const {
        const int64_t node = list[i];
        if (compact) compact[i] = id[node];
        if (flat) flat[i] = node;
        if (ns_out)
            for (int a = 0; a < D; ++a) ns_out[a * n + i] = ns[a * g.vol + (int64_t)NG * g.plane + node];
    }
};

}  // namespace lbm
