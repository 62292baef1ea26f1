"""CPU ORACLE (test infrastructure, NOT product code) -- D2Q9 colour-gradient CSF path.

A NumPy restatement, kernel by kernel and in the reference's launch order, of
openLBMPM's `RKColorGradientLBM.runRKColorGradient2DCSF`
(/root/reference/RKCG2D/RKD2Q9.py:1225-1490, commit 3d84189) and of the
Numba-CUDA kernels it launches (/root/reference/RKCG2D/AcceleratedRKGPU2D.py).
It keeps the reference's own data layout (compact fluid-node list, AoS `[N, 9]`
populations, int64 neighbour table `[8N]`) so that every intermediate can be
compared with the reference one to one.

Pinned: `tests/test_oracle_cg2d.py` checks this file against the golden
vectors in `tests/golden/cg2d_*.npz`, which were produced by running the
reference driver and kernels themselves (tests/golden/gen_goldens_cg2d.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import this module; the product (openlbmpm_b200) never does.

Conventions: arrays `[y, x]`, flat id `i*nx + j`,
e = [(0,0),(1,0),(0,1),(-1,0),(0,-1),(1,1),(-1,1),(-1,-1),(1,-1)],
inlet = top (BC row ny-2, ghost row ny-1), outlet = bottom (BC row 1, ghost 0).
"""
import numpy as np

EX = np.array([0., 1., 0., -1., 0., 1., -1., -1., 1.])
EY = np.array([0., 0., 1., 0., -1., 1., 1., -1., -1.])
W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
OPP = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])
# neighbour slot k (0..7) <-> direction k+1; (dy, dx)
NB_OFF = [(0, 1), (1, 0), (0, -1), (-1, 0), (1, 1), (1, -1), (-1, -1), (-1, 1)]


def mrt_matrix():
    """RKD2Q9.py:308-336 (rows: rho, e, eps, jx, qx, jy, qy, pxx, pxy)."""
    M = np.array([
        [1, 1, 1, 1, 1, 1, 1, 1, 1],
        [-4, -1, -1, -1, -1, 2, 2, 2, 2],
        [4, -2, -2, -2, -2, 1, 1, 1, 1],
        [0, 1, 0, -1, 0, 1, -1, -1, 1],
        [0, -2, 0, 2, 0, 1, -1, -1, 1],
        [0, 0, 1, 0, -1, 1, 1, -1, -1],
        [0, 0, -2, 0, 2, 1, 1, -1, -1],
        [0, 1, -1, 1, -1, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 1, -1, 1, -1]], dtype=np.float64)
    return M, np.linalg.inv(M)


S_FIXED = np.array([0., 1.64, 1.54, 0., 1.9, 0., 1.9, 0., 0.])  # RKD2Q9.py:338-340


# ----------------------------------------------------------------------------
# host-side indexing  (RKD2Q9.py:657-892)
# ----------------------------------------------------------------------------
def build_indexing(is_domain):
    """optimizeFluidandSolidArray + sortOutFluidNodesToSolid + calVectorNormaltoSolid."""
    dom = np.asarray(is_domain, dtype=bool)
    ny, nx = dom.shape
    flat = dom.ravel()
    fluid_nodes = np.flatnonzero(flat).astype(np.int64)
    new_idx = -np.ones((ny, nx), dtype=np.int64)
    new_idx.ravel()[fluid_nodes] = np.arange(fluid_nodes.size, dtype=np.int64)
    # number of fluid / solid cells in the periodic 3x3 neighbourhood
    nfluid = np.zeros((ny, nx), dtype=np.int64)
    for m in (-1, 0, 1):
        for n in (-1, 0, 1):
            nfluid += np.roll(np.roll(dom, -m, 0), -n, 1)
    wet_solid = (~dom) & (nfluid > 0)                      # RKD2Q9.py:677-689
    wetting_nodes = np.flatnonzero(wet_solid.ravel()).astype(np.int64)
    new_idx.ravel()[wetting_nodes] = -2 - np.arange(wetting_nodes.size, dtype=np.int64)

    def neighbours(nodes):                                  # AcceleratedRKGPU2D.py:14-95
        i = nodes // nx
        j = nodes % nx
        out = np.empty((nodes.size, 8), dtype=np.int64)
        for k, (dy, dx) in enumerate(NB_OFF):
            out[:, k] = new_idx[(i + dy) % ny, (j + dx) % nx]
        return out.ravel()

    nb = neighbours(fluid_nodes)
    nb_s = neighbours(wetting_nodes)
    near = dom & (nfluid < 9)                               # RKD2Q9.py:741-760
    fn_orig = np.flatnonzero(near.ravel()).astype(np.int64)
    fn_gpu = new_idx.ravel()[fn_orig]
    # solid normals, same accumulation order as RKD2Q9.py:806-880
    offs = [((1, 0), 4. / 21.), ((0, 1), 4. / 21.), ((-1, 0), 4. / 21.), ((0, -1), 4. / 21.),
            ((1, 1), 4. / 45.), ((-1, 1), 4. / 45.), ((-1, -1), 4. / 45.), ((1, -1), 4. / 45.),
            ((2, 0), 1. / 60.), ((0, 2), 1. / 60.), ((-2, 0), 1. / 60.), ((0, -2), 1. / 60.),
            ((2, 1), 2. / 315.), ((1, 2), 2. / 315.), ((-1, 2), 2. / 315.), ((-2, 1), 2. / 315.),
            ((-2, -1), 2. / 315.), ((-1, -2), 2. / 315.), ((1, -2), 2. / 315.), ((2, -1), 2. / 315.),
            ((2, 2), 1. / 5040.), ((-2, 2), 1. / 5040.), ((-2, -2), 1. / 5040.), ((2, -2), 1. / 5040.)]
    yi = fn_orig // nx
    xi = fn_orig % nx
    sx = np.zeros(fn_orig.size)
    sy = np.zeros(fn_orig.size)
    for (cx, cy), wgt in offs:
        solid = ~dom[(yi + cy) % ny, (xi + cx) % nx]
        sx = sx + np.where(solid, wgt * 1. * cx, 0.)
        sy = sy + np.where(solid, wgt * 1. * cy, 0.)
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = np.sqrt(sx * sx + sy * sy)
        nsx = sx / nrm
        nsy = sy / nrm
    return dict(nx=nx, ny=ny, fluidNodes=fluid_nodes, newIndicesDomain=new_idx,
                neighboringNodes=nb, wettingSolidNodes=wetting_nodes,
                neighboringWettingSolidNodes=nb_s, fluidNodesWithSolidGPU=fn_gpu,
                fluidNodesWithSolidOriginal=fn_orig, nsX=nsx, nsY=nsy)


def feq(rho, ux, uy):
    """calEquilibriumRK2D, AcceleratedRKGPU2D.py:169-175 -> [N, 9]."""
    eu = EX[None, :] * ux[:, None] + EY[None, :] * uy[:, None]
    uu = (ux * ux + uy * uy)[:, None]
    return rho[:, None] * W[None, :] * (1 + (3. * eu + 4.5 * eu * eu - 1.5 * uu))


class CG2D:
    """State + one-step operators in the reference's launch order."""

    def __init__(self, is_domain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98,
                 tauR=1.0, tauB=1.0, tautype=2, relax="MRT", inlet="Periodic", outlet="Periodic",
                 vy_inlet=0.0, dBH=5e-8, dRH=1.0, dBL=1.0, dRL=5e-8):
        self.idx = build_indexing(is_domain)
        self.dom = np.asarray(is_domain, bool)
        self.ny, self.nx = self.dom.shape
        self.N = self.idx["fluidNodes"].size
        self.nb = self.idx["neighboringNodes"].reshape(self.N, 8)
        self.nbS = self.idx["neighboringWettingSolidNodes"].reshape(-1, 8)
        self.Ns = self.nbS.shape[0]
        self.sigma, self.beta, self.delta = sigma, beta, delta
        self.cosT = np.cos(theta_deg / 180. * np.pi)       # RKD2Q9.py:86-87
        self.sinT = np.sin(theta_deg / 180. * np.pi)
        self.wetting, self.tauR, self.tauB, self.tautype = wetting, tauR, tauB, tautype
        self.relax, self.inlet, self.outlet = relax, inlet, outlet
        self.vy, self.dBH, self.dRH, self.dBL, self.dRL = vy_inlet, dBH, dRH, dBL, dRL
        self.M, self.Mi = mrt_matrix()
        fn = self.idx["fluidNodes"]
        nx, ny = self.nx, self.ny
        self.row_in = np.flatnonzero((fn >= (ny - 2) * nx) & (fn < (ny - 1) * nx))
        self.row_in_ghost = np.flatnonzero((fn >= (ny - 1) * nx) & (fn < ny * nx))
        self.row0 = np.flatnonzero(fn < nx)
        self.row1 = np.flatnonzero((fn >= nx) & (fn < 2 * nx))
        self.row2 = np.flatnonzero((fn >= 2 * nx) & (fn < 3 * nx))

    # -- initial condition (RKD2Q9.py:445-490, 561-585; zero velocity) ---------
    def set_densities(self, rhoR2d, rhoB2d):
        fn = self.idx["fluidNodes"]
        self.rhoR = np.asarray(rhoR2d, float).ravel()[fn].copy()
        self.rhoB = np.asarray(rhoB2d, float).ravel()[fn].copy()
        z = np.zeros(self.N)
        self.fR = feq(self.rhoR, z, z)
        self.fB = feq(self.rhoB, z, z)
        self.fT = self.fR + self.fB                        # RKD2Q9.py:1259
        self.ux = z.copy(); self.uy = z.copy()
        self.Fx = z.copy(); self.Fy = z.copy()             # zero-initialised
        self.phi = z.copy(); self.Gx = z.copy(); self.Gy = z.copy(); self.K = z.copy()
        self.phiS = np.zeros(self.Ns)

    def set_pdfs(self, fR2d, fB2d, rhoR2d, rhoB2d):
        fn = self.idx["fluidNodes"]
        self.set_densities(rhoR2d, rhoB2d)
        self.fR = np.asarray(fR2d, float).reshape(-1, 9)[fn].copy()
        self.fB = np.asarray(fB2d, float).reshape(-1, 9)[fn].copy()
        self.fT = self.fR + self.fB

    # -- open boundaries --------------------------------------------------------
    def _inlet_neumann(self):
        """constantTotalVelocityInlet (2345-2423) + ghostPointsConstantVelocityRK (604-650)."""
        r = self.row_in; v = self.vy; fT = self.fT
        if r.size:
            rho = (fT[r, 0] + fT[r, 1] + fT[r, 3] + 2. * (fT[r, 2] + fT[r, 5] + fT[r, 6])) / (1. + v)
            eq = lambda w, ev: rho * w * (1. + 3. * ev + 4.5 * ev * ev - 1.5 * (v * v))
            fT[r, 4] = eq(1. / 9., -v) + (fT[r, 2] - eq(1. / 9., v))
            fT[r, 7] = eq(1. / 36., -v) + (fT[r, 5] - eq(1. / 36., v))
            fT[r, 8] = eq(1. / 36., -v) + (fT[r, 6] - eq(1. / 36., v))
            ratioR = self.rhoR[r] / (self.rhoR[r] + self.rhoB[r])
            self.rhoR[r] = ratioR * rho
            for k in (4, 7, 8):
                self.fR[r, k] = ratioR * fT[r, k]
            ratioB = self.rhoB[r] / (self.rhoR[r] + self.rhoB[r])   # uses the UPDATED rhoR (2399-2407)
            self.rhoB[r] = ratioB * rho
            for k in (4, 7, 8):
                self.fB[r, k] = ratioB * fT[r, k]
            self.uy[r] = v
        g = self.row_in_ghost
        if g.size:
            src = self.nb[g, 3]
            self.fR[g] = self.fR[src]; self.fB[g] = self.fB[src]
            self.rhoR[g] = _sum9(self.fR[g]); self.rhoB[g] = _sum9(self.fB[g])

    def _inlet_dirichlet(self):
        """calConstPressureInletGPU (923-961) + ghostPointsConstPressureInletRK (966-1001)."""
        r = self.row_in
        for f, rho, p in ((self.fB, self.rhoB, self.dBH), (self.fR, self.rhoR, self.dRH)):
            if r.size:
                v = -1. + (f[r, 0] + f[r, 1] + f[r, 3] + 2. * (f[r, 2] + f[r, 5] + f[r, 6])) / p
                f[r, 4] = f[r, 2] - 2. / 3. * p * v
                f[r, 7] = f[r, 5] + 1. / 2. * (f[r, 1] - f[r, 3]) - 1. / 6. * p * v
                f[r, 8] = f[r, 6] - 1. / 2. * (f[r, 1] - f[r, 3]) - 1. / 6. * p * v
                rho[r] = p
        g = self.row_in_ghost
        if g.size:
            src = self.nb[g, 3]
            self.fR[g] = self.fR[src]; self.fB[g] = self.fB[src]
            self.rhoR[g] = self.rhoR[src]; self.rhoB[g] = self.rhoB[src]

    def _outlet_convective(self):
        """convectiveOutletGPU / Ghost2 / Ghost3 (698-784): rows 2, 1, 0 in that order."""
        for r in (self.row2, self.row1, self.row0):
            if r.size:
                src = self.nb[r, 1]
                self.fR[r] = self.fR[src]; self.fB[r] = self.fB[src]
                self.rhoR[r] = _sum9(self.fR[r]); self.rhoB[r] = _sum9(self.fB[r])

    def _outlet_dirichlet(self):
        """calConstPressureLowerGPUTotal (2557-2602) + ghostPointsConstPressureLowerRK (1043-1080)."""
        r = self.row1; fT = self.fT; p = self.dBL + self.dRL   # RKD2Q9.py:1344
        if r.size:
            v = 1. - 1. / p * (fT[r, 0] + fT[r, 1] + fT[r, 3] + 2. * (fT[r, 4] + fT[r, 7] + fT[r, 8]))
            fT[r, 2] = fT[r, 4] + 2. / 3. * (p * v)
            fT[r, 5] = fT[r, 7] + 0.5 * (fT[r, 3] - fT[r, 1]) + 1. / 6. * p * v
            fT[r, 6] = fT[r, 8] + 0.5 * (fT[r, 1] - fT[r, 3]) + 1. / 6. * p * v
            self.uy[r] = v
            ratioR = self.rhoR[r] / (self.rhoR[r] + self.rhoB[r])
            ratioB = self.rhoB[r] / (self.rhoR[r] + self.rhoB[r])
            for k in (2, 5, 6):
                self.fR[r, k] = ratioR * fT[r, k]
                self.fB[r, k] = ratioB * fT[r, k]
        g = np.arange(min(self.nx, self.N))                # `indices < nx` on COMPACT ids (1055)
        src = self.nb[g, 1]
        self.fR[g] = self.fR[src]; self.fB[g] = self.fB[src]
        self.rhoR[g] = self.rhoR[src]; self.rhoB[g] = self.rhoB[src]

    # -- per-node pieces --------------------------------------------------------
    def _tau(self):
        """tau(phi) shared by the four collision/forcing kernels (e.g. 1820-1834)."""
        phi = self.phi
        tau = np.ones(self.N)
        hi = phi > self.delta; lo = phi < -self.delta; mid = np.abs(phi) <= self.delta
        tau[hi] = self.tauR; tau[lo] = self.tauB
        with np.errstate(invalid="ignore", divide="ignore"):
            if self.tautype == 1:
                t = 0.5 + 1. / ((1. + phi) / (2. * (self.tauR - 0.5)) + (1. - phi) / (2. * (self.tauB - 0.5)))
                tau[mid] = t[mid]
            elif self.tautype == 2:
                ratioR = self.rhoR / (self.rhoR + self.rhoB)
                ratioB = self.rhoB / (self.rhoR + self.rhoB)
                miu = 1. / (ratioR * (3. / (self.tauR - 0.5)) + ratioB * (3. / (self.tauB - 0.5)))
                tau[mid] = (3. * miu + 0.5)[mid]
        return tau

    def head(self):
        """Everything up to the output point of the loop (RKD2Q9.py:1299-1368)."""
        if self.inlet == "Neumann":
            self._inlet_neumann()
        if self.inlet == "Dirichlet":
            self._inlet_dirichlet()
        if self.outlet == "Convective":
            self._outlet_convective()
        elif self.outlet == "Dirichlet":
            self._outlet_dirichlet()
        self.fT = self.fR + self.fB                                        # calTotalFluidPDF
        fT = self.fT
        with np.errstate(invalid="ignore", divide="ignore"):
            rho = self.rhoB + self.rhoR                                    # calPhysicalVelocityRKGPU2DNew1
            self.ux = (fT[:, 1] - fT[:, 3] + fT[:, 5] - fT[:, 6] - fT[:, 7] + fT[:, 8] + 0.5 * self.Fx) / rho
            self.uy = (fT[:, 2] - fT[:, 4] + fT[:, 5] + fT[:, 6] - fT[:, 7] - fT[:, 8] + 0.5 * self.Fy) / rho
            self.phi = (self.rhoR - self.rhoB) / (self.rhoR + self.rhoB)   # calPhaseFieldPhi

    def body(self):
        """The rest of the loop body (RKD2Q9.py:1395-1490)."""
        N, nb = self.N, self.nb
        fluid_nb = nb >= 0
        with np.errstate(invalid="ignore", divide="ignore"):
            # calColorValueOnSolid (1559-1580)
            if self.Ns > 0:
                m = self.nbS >= 0
                vals = np.where(m, self.phi[np.where(m, self.nbS, 0)], 0.)
                num = np.zeros(self.Ns); den = np.zeros(self.Ns)
                for k in range(8):
                    num = num + np.where(m[:, k], W[k + 1] * vals[:, k], 0.)
                    den = den + np.where(m[:, k], W[k + 1], 0.)
                self.phiS = num / den
            # calRKInitialGradient (1582-1632)
            phis_pad = self.phiS if self.Ns > 0 else np.zeros(1)
            gx = np.zeros(N); gy = np.zeros(N)
            for k in range(8):
                n = nb[:, k]
                val = np.where(n >= 0, self.phi[np.where(n >= 0, n, 0)],
                               phis_pad[np.clip(-n - 2, 0, phis_pad.size - 1)])
                gx = gx + W[k + 1] * val * EX[k + 1]
                gy = gy + W[k + 1] * val * EY[k + 1]
            self.Gx = 3. * gx; self.Gy = 3. * gy
            # wetting boundary condition on the colour gradient
            if self.Ns > 0:
                if self.wetting == 1:
                    self._wetting_xu()
                elif self.wetting == 2:
                    self._wetting_akai()
            # CSF force
            self._force()
            # collision on the total population + forcing
            tau = self._tau()
            fe = feq(self.rhoR, self.ux, self.uy) + feq(self.rhoB, self.ux, self.uy)
            if self.relax == "SRT":
                self.fT = -1. / tau[:, None] * (self.fT - fe) + self.fT          # 1801-1849
                ux, uy = self.ux[:, None], self.uy[:, None]
                eu = EX[None, :] * ux + EY[None, :] * uy
                src = W[None, :] * ((3. * (EX[None, :] - ux) + 9. * EX[None, :] * eu) * self.Fx[:, None] +
                                    (3. * (EY[None, :] - uy) + 9. * EY[None, :] * eu) * self.Fy[:, None]) * \
                    (1. - 1. / (2. * tau))[:, None]
                self.fT = self.fT + src                                           # 1740-1796
            else:
                S = np.tile(S_FIXED, (N, 1)); S[:, 7] = 1. / tau; S[:, 8] = 1. / tau
                m = (self.fT - fe) @ self.M.T                                     # 1934-2018
                self.fT = self.fT - (m * S) @ self.Mi.T
                ux, uy, Fx, Fy = self.ux[:, None], self.uy[:, None], self.Fx[:, None], self.Fy[:, None]
                ex, ey = EX[None, :], EY[None, :]
                src = W[None, :] * (ex * Fx * 3. + ey * Fy * 3. + (ex * ex - 1. / 3.) * ux * Fx * 9. +
                                    ex * ey * uy * Fx * 9. + ey * ex * ux * Fy * 9. +
                                    (ey * ey - 1. / 3.) * uy * Fy * 9.)
                S2 = 1. - 0.5 * S
                self.fT = self.fT + ((src @ self.M.T) * S2) @ self.Mi.T           # 2023-2114
            # calRecoloringProcessM (1854-1900)
            gn = np.sqrt(self.Gx * self.Gx + self.Gy * self.Gy)
            en = np.sqrt(EX * EX + EY * EY)
            ok = (gn[:, None] > 1.0e-8) & (en[None, :] > 1.0e-8)
            cost = np.where(ok, (EX[None, :] * self.Gx[:, None] + EY[None, :] * self.Gy[:, None]) /
                            np.where(ok, en[None, :] * gn[:, None], 1.), 0.)
            tot = self.rhoR + self.rhoB
            a = (self.beta * self.rhoR * self.rhoB / tot)[:, None] * W[None, :] * cost * en[None, :]
            self.fR = (self.rhoR / tot)[:, None] * self.fT + a
            self.fB = (self.rhoB / tot)[:, None] * self.fT - a
        # calStreaming1GPU / calStreaming2GPU (338-417): push + half-way bounce back
        for f in (self.fR, self.fB):
            new = f.copy()
            me = np.arange(N)
            for k in range(8):
                d = k + 1
                n = nb[:, k]
                fl = n >= 0
                new[n[fl], d] = f[me[fl], d]
                new[me[~fl], OPP[d]] = f[me[~fl], d]
            f[:, 1:] = new[:, 1:]
        self.fT = self.fR + self.fB
        self.rhoR = _sum9(self.fR); self.rhoB = _sum9(self.fB)     # calMacroDensityRKGPU2D (101-118)

    def step(self, n=1):
        for _ in range(n):
            self.head()
            self.body()

    # -- pieces of body() ---------------------------------------------------------
    def _wetting_xu(self):
        """updateColorGradientOnWetting (1637-1679), Xu et al. 2017."""
        loc = self.idx["fluidNodesWithSolidGPU"]
        nsx, nsy, c, s = self.idx["nsX"], self.idx["nsY"], self.cosT, self.sinT
        n1x = nsx * c - nsy * s; n1y = nsy * c + nsx * s
        n2x = nsx * c + nsy * s; n2y = nsy * c - nsx * s
        gx, gy = self.Gx[loc], self.Gy[loc]
        gn = np.sqrt(gx * gx + gy * gy)
        big = gn > 1.0e-8
        ugx = np.where(big, gx / np.where(big, gn, 1.), 0.)
        ugy = np.where(big, gy / np.where(big, gn, 1.), 0.)
        d1 = np.sqrt((ugx - n1x) ** 2 + (ugy - n1y) ** 2)
        d2 = np.sqrt((ugx - n2x) ** 2 + (ugy - n2y) ** 2)
        mx = np.where(d1 < d2, n1x, np.where(d1 > d2, n2x, np.where(d1 == d2, nsx, 0.)))
        my = np.where(d1 < d2, n1y, np.where(d1 > d2, n2y, np.where(d1 == d2, nsy, 0.)))
        self.Gx[loc] = gn * mx; self.Gy[loc] = gn * my

    def _wetting_akai(self):
        """updateColorGradientOnWettingNew (2428-2492), Akai et al. 2018."""
        loc = self.idx["fluidNodesWithSolidGPU"]
        nsx, nsy, c, s = self.idx["nsX"], self.idx["nsY"], self.cosT, self.sinT
        gx, gy = self.Gx[loc], self.Gy[loc]
        gn = np.sqrt(gx * gx + gy * gy)
        big = gn > 1.0e-8
        ux = np.where(big, -gx / np.where(big, gn, 1.), 0.)
        uy = np.where(big, -gy / np.where(big, gn, 1.), 0.)
        dot = ux * nsx + uy * nsy
        th = np.arccos(np.clip(dot, -1., 1.))      # GPU: NaN outside [-1,1] -> same "no update" outcome
        sth = np.sin(th); cth = np.cos(th)
        ok = np.abs(sth) > 1.0e-9
        sth_ = np.where(ok, sth, 1.)
        c1 = np.where(ok, s * cth / sth_, 0.); c2 = np.where(ok, s / sth_, 0.)
        c3 = np.where(ok, -s * cth / sth_, 0.); c4 = np.where(ok, -s / sth_, 0.)
        x1 = (c - c1) * nsx + c2 * ux; y1 = (c - c1) * nsy + c2 * uy
        x2 = (c - c3) * nsx + c4 * ux; y2 = (c - c3) * nsy + c4 * uy
        d1 = np.sqrt((x1 - ux) ** 2 + (y1 - uy) ** 2)
        d2 = np.sqrt((x2 - ux) ** 2 + (y2 - uy) ** 2)
        newx = np.where(d1 < d2, -gn * x1, np.where(d1 > d2, -gn * x2, gx))
        newy = np.where(d1 < d2, -gn * y1, np.where(d1 > d2, -gn * y2, gy))
        self.Gx[loc] = newx; self.Gy[loc] = newy

    def _force(self):
        """calForceTermInColorGradient2D (1684-1735) / ...New2D (2497-2552)."""
        nb = self.nb
        gn = np.sqrt(self.Gx * self.Gx + self.Gy * self.Gy)
        if self.wetting == 1:
            big = gn > 0.
            sign = 1.
        else:
            big = gn > 1.0e-8
            sign = -1.
        safe = np.where(big, gn, 1.)
        unx = np.where(big, sign * self.Gx / safe, 0.)
        uny = np.where(big, sign * self.Gy / safe, 0.)
        pyx = np.zeros(self.N); pxy = np.zeros(self.N); px = np.zeros(self.N); py = np.zeros(self.N)
        for k in range(8):
            n = nb[:, k]; fl = n >= 0; nn = np.where(fl, n, 0)
            nxk = np.where(fl, unx[nn], 0.); nyk = np.where(fl, uny[nn], 0.)
            pyx = pyx + 3. * W[k + 1] * nyk * EX[k + 1]
            pxy = pxy + 3. * W[k + 1] * nxk * EY[k + 1]
            px = px + 3. * W[k + 1] * nxk * EX[k + 1]
            py = py + 3. * W[k + 1] * nyk * EY[k + 1]
        self.K = unx * uny * (pyx + pxy) - uny * uny * px - unx * unx * py
        self.Fx = sign * 0.5 * self.sigma * self.K * self.Gx
        self.Fy = sign * 0.5 * self.sigma * self.K * self.Gy

    # -- what the reference writes to HDF5 (RKD2Q9.py:902-957) --------------------
    def to_dense(self):
        fn = self.idx["fluidNodes"]
        out = {}
        for name, a in (("rhoR", self.rhoR), ("rhoB", self.rhoB), ("ux", self.ux), ("uy", self.uy)):
            d = np.zeros(self.ny * self.nx); d[fn] = a
            out[name] = d.reshape(self.ny, self.nx)
        for name, a in (("pdfR", self.fR), ("pdfB", self.fB)):
            d = np.zeros((self.ny * self.nx, 9)); d[fn] = a
            out[name] = d.reshape(self.ny, self.nx, 9)
        return out


def _sum9(f):
    """Left-to-right sum over the 9 directions (101-118)."""
    s = f[:, 0].copy()
    for k in range(1, 9):
        s = s + f[:, k]
    return s
