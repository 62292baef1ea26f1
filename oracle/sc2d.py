"""CPU ORACLE (test infrastructure, NOT product code) -- D2Q9 Shan-Chen paths on a dense masked grid.

NumPy restatement of the two live Shan-Chen drivers of openLBMPM (reference commit 3d84189):
  * original Shan-Chen:        ShanChenD2Q9.runOptimizedLBM   (ShanChen2D/ShanChenD2Q9.py:1433-1629)
    kernels in ShanChen2D/OptimizedD2Q9GPU.py
  * explicit forcing SRT/MRT:  ShanChenD2Q9.runOptimizedEFLBM (ShanChen2D/ShanChenD2Q9.py:1631-2087)
    kernels in ShanChen2D/ExplicitD2Q9GPU.py, isotropy 4, 8 and 10
in the reference's kernel order, keeping its materialised f_eq / force-distribution arrays and its dense
C = M^-1 S M product for MRT.

PINNED: tests/test_oracle_sc2d.py checks it against tests/golden/sc2d_*.npz, produced by the reference
drivers themselves (tests/golden/gen_goldens_sc2d.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import numpy as np

EX = np.array([0., 1., 0., -1., 0., 1., -1., -1., 1.])
EY = np.array([0., 0., 1., 0., -1., 1., 1., -1., -1.])
W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
OPP = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])


def mrt_matrix():
    """SimpleD2Q9.py:107-124"""
    return np.array([[1, 1, 1, 1, 1, 1, 1, 1, 1], [-4, -1, -1, -1, -1, 2, 2, 2, 2], [4, -2, -2, -2, -2, 1, 1, 1, 1],
                     [0, 1, 0, -1, 0, 1, -1, -1, 1], [0, -2, 0, 2, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1],
                     [0, 0, -2, 0, 2, 1, 1, -1, -1], [0, 1, -1, 1, -1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, -1, 1, -1]], float)


def shift(a, k):
    """value at x + e_k (periodic) of a dense [..., y, x] array"""
    return np.roll(a, (-int(EY[k]), -int(EX[k])), axis=(-2, -1))


class SC2D:
    def __init__(self, is_domain, model="ShanChen", relax="SRT", tau=(1., 1.), G=3.8, Gs=(-0.4, 0.4),
                 inlet="Periodic", outlet="Periodic", vy=(0., 0.), rho_out=(1.0, 0.02), scheme=4):
        self.dom = np.asarray(is_domain, bool)
        self.ny, self.nx = self.dom.shape
        self.model, self.relax, self.inlet, self.outlet = model, relax, inlet, outlet
        self.tau = np.asarray(tau, float); self.nc = self.tau.size
        self.G = np.zeros((self.nc, self.nc))
        self.G[0, 1] = self.G[1, 0] = G                               # ShanChenD2Q9.py:224-231
        self.Gs = np.asarray(Gs, float); self.vy = np.asarray(vy, float); self.rho_out = np.asarray(rho_out, float)
        if relax == "MRT":                                             # ShanChenD2Q9.py:96-106, 484-496
            M = mrt_matrix(); Mi = np.linalg.inv(M)
            self.C = []
            for k in range(self.nc):
                s = np.ones(9)
                if k < 2:
                    s[1], s[2], s[4], s[6] = 0.6, 1.5, 1.2, 1.2
                s[7] = s[8] = 1. / self.tau[k]
                self.C.append(Mi @ np.diag(s) @ M)
        self.scheme = int(scheme)         # [ForceScheme] ExplicitScheme: 4, 8 or 10
        deep = 1 if (self.scheme == 8 and model == "EFS") else 0      # ...Higher8 / ...Lower8 act one row further in
        self.z_in, self.z_out = self.ny - 2 - deep, 1 + deep
        self.prepared = False

    def set_densities(self, rho):
        """rho: [nc, ny, nx]; f = w rho (ShanChenD2Q9.py:759-768)"""
        self.rho = np.where(self.dom, np.asarray(rho, float), 0.)
        self.f = self.rho[:, :, :, None] * W                        # [nc, ny, nx, 9]
        self.F = np.zeros((self.nc, 2, self.ny, self.nx))
        self.uph = np.zeros((2, self.ny, self.nx))
        self.prepared = False

    # -- shared kernels -------------------------------------------------------------------------
    def _rho(self):
        """calFluidRhoGPU (OptimizedD2Q9GPU.py:84-93)"""
        s = self.f[..., 0].copy()
        for q in range(1, 9):
            s = s + self.f[..., q]
        self.rho = np.where(self.dom, s, 0.)

    def _stream(self):
        """calStreaming1GPU / calStreaming2GPU (OptimizedD2Q9GPU.py:450-548): push + half-way bounce back"""
        new = self.f.copy()
        for q in range(1, 9):
            src_fluid = np.roll(self.dom, (int(EY[q]), int(EX[q])), axis=(0, 1))       # node at x - e_q
            pulled = np.roll(self.f[..., q], (int(EY[q]), int(EX[q])), axis=(1, 2))
            new[..., q] = np.where(src_fluid, pulled, self.f[..., OPP[q]])
        self.f = np.where(self.dom[None, :, :, None], new, 0.)

    def _uphys(self):
        """calPhysicalVelocity (OptimizedD2Q9GPU.py:156-175)"""
        f = self.f
        vx = np.zeros((self.ny, self.nx)); vy = np.zeros((self.ny, self.nx)); r = np.zeros((self.ny, self.nx))
        for k in range(self.nc):
            vx = vx + (f[k, ..., 1] - f[k, ..., 3] + f[k, ..., 5] - f[k, ..., 6] - f[k, ..., 7] + f[k, ..., 8] + 0.5 * self.F[k, 0])
            vy = vy + (f[k, ..., 2] - f[k, ..., 4] + f[k, ..., 5] + f[k, ..., 6] - f[k, ..., 7] - f[k, ..., 8] + 0.5 * self.F[k, 1])
            r = r + self.rho[k]
        with np.errstate(invalid="ignore", divide="ignore"):
            self.uph = np.where(self.dom, np.stack([vx, vy]) / r, 0.)

    def _inlet(self):
        """constantVelocityZouHeBoundaryHigher (839-861) + ghostPointsConstantVelocityInlet (710-736)"""
        if self.inlet != "Neumann":
            return
        r = self.z_in
        m = self.dom[r]
        for k in range(self.nc):
            f = self.f[k, r]; v = self.vy[k]
            rho = (f[:, 0] + f[:, 1] + f[:, 3] + 2. * (f[:, 2] + f[:, 5] + f[:, 6])) / (1. + v)
            self.rho[k, r] = np.where(m, rho, 0.)
            f4 = f[:, 2] - 2. / 3. * rho * v
            f7 = f[:, 5] + (f[:, 1] - f[:, 3]) / 2. - 1. / 6. * rho * v
            f8 = f[:, 6] - (f[:, 1] - f[:, 3]) / 2. - 1. / 6. * rho * v
            for q, val in ((4, f4), (7, f7), (8, f8)):
                f[:, q] = np.where(m, val, f[:, q])
        for row in range(r, self.ny - 1):          # ghost rows, one after the other (ghostPointsConstantVelocity8 / 82)
            self._row_copy(row + 1, row)

    def _row_copy(self, dst, src):
        m = self.dom[dst] & self.dom[src]
        self.f[:, dst] = np.where(m[None, :, None], self.f[:, src], self.f[:, dst])
        s = self.f[:, dst, :, 0].copy()
        for q in range(1, 9):
            s = s + self.f[:, dst, :, q]
        self.rho[:, dst] = np.where(m[None], s, self.rho[:, dst])

    def _outlet_pressure(self):
        """constantPressureZouHeBoundaryLower (555-584, densities hard-coded) + ghostPointsConstantPressureOutlet (743-768)"""
        zo = self.z_out
        m = self.dom[zo]
        for k in range(self.nc):
            f = self.f[k, zo]; d = self.rho_out[k]
            vy = 1. - (f[:, 0] + f[:, 1] + f[:, 3] + 2. * (f[:, 4] + f[:, 7] + f[:, 8])) / d
            f2 = f[:, 4] + 2. / 3. * vy * d
            f5 = f[:, 7] + 1. / 2. * (f[:, 3] - f[:, 1]) + 1. / 6. * d * vy
            f6 = f[:, 8] - 1. / 2. * (f[:, 3] - f[:, 1]) + 1. / 6. * d * vy
            for q, val in ((2, f2), (5, f5), (6, f6)):
                f[:, q] = np.where(m, val, f[:, q])
            self.rho[k, zo] = np.where(m, d, 0.)
        for row in range(zo, 0, -1):               # ghostPointsConstantPressureOutlet8 / 82
            self._row_copy(row - 1, row)

    # -- original Shan-Chen ---------------------------------------------------------------------
    def _sc_iteration(self):
        self._inlet()
        self._rho()
        psi = self.rho
        f = self.f
        with np.errstate(invalid="ignore", divide="ignore"):
            vxt = sum((f[k, ..., 1] - f[k, ..., 3] + f[k, ..., 5] - f[k, ..., 6] - f[k, ..., 7] + f[k, ..., 8]) / self.tau[k]
                      for k in range(self.nc))
            vyt = sum((f[k, ..., 2] - f[k, ..., 4] + f[k, ..., 5] + f[k, ..., 6] - f[k, ..., 7] - f[k, ..., 8]) / self.tau[k]
                      for k in range(self.nc))
            rt = sum(self.rho[k] / self.tau[k] for k in range(self.nc))
            upx, upy = vxt / rt, vyt / rt
            for k in range(self.nc):                                   # interactionCollisionProcess (1274-1446)
                fx = np.zeros((self.ny, self.nx)); fy = np.zeros((self.ny, self.nx))
                for q in range(1, 9):
                    wI = 1. / 9. if q < 5 else 1. / 36.
                    fl = shift(self.dom, q)
                    t_f = sum(-wI * self.G[k, j] * psi[k] * shift(psi[j], q) for j in range(self.nc))
                    t = np.where(fl, t_f, -wI * self.Gs[k] * psi[k])
                    fx = fx + t * EX[q]; fy = fy + t * EY[q]
                self.F[k, 0] = np.where(self.dom, fx, 0.); self.F[k, 1] = np.where(self.dom, fy, 0.)
                ux = upx + self.tau[k] * fx / psi[k]; uy = upy + self.tau[k] * fy / psi[k]
                uu = ux * ux + uy * uy
                for q in range(9):
                    eu = EX[q] * ux + EY[q] * uy
                    f[k, ..., q] = (1. - 1. / self.tau[k]) * f[k, ..., q] + W[q] * psi[k] / self.tau[k] * \
                        (1. + 3. * eu + 4.5 * (eu * eu) - 1.5 * uu)
        self.f = np.where(self.dom[None, :, :, None], f, 0.)
        self._stream()
        self._rho()
        if self.outlet == "Convective":                                # convectiveOutletGPU / Ghost2 / Ghost3 (960-1036)
            self._row_copy(2, 3); self._row_copy(1, 2); self._row_copy(0, 1)
        self._uphys()

    # -- explicit forcing -----------------------------------------------------------------------
    def _efs_force_ueq(self):
        """calExplicit4thOrderScheme (ExplicitD2Q9GPU.py:51-217) + calEquilibriumVEFGPU (340-363) /
        transformEquilibriumVelocity (1426-1449)"""
        psi = self.rho
        mx = np.zeros((self.ny, self.nx)); my = np.zeros((self.ny, self.nx)); rt = np.zeros((self.ny, self.nx))
        for k in range(self.nc):
            gx = np.zeros((self.ny, self.nx)); gy = gx.copy(); sx = gx.copy(); sy = gx.copy()
            for q in range(1, 9):
                wI = 1. / 3. if q < 5 else 1. / 12.
                fl = shift(self.dom, q)
                grad = sum(wI * (shift(psi[j], q) - psi[j]) * self.G[k, j] for j in range(self.nc))
                gx = gx + np.where(fl, grad * EX[q], 0.); gy = gy + np.where(fl, grad * EY[q], 0.)
                s = -wI * self.Gs[k] * psi[k]
                sx = sx + np.where(fl, 0., s * EX[q]); sy = sy + np.where(fl, 0., s * EY[q])
            fx = -6.0 * psi[k] * gx + sx; fy = -6.0 * psi[k] * gy + sy
            if self.scheme != 4:
                fx, fy = self._force_iso(k)
            self.F[k, 0] = np.where(self.dom, fx, 0.); self.F[k, 1] = np.where(self.dom, fy, 0.)
            ex = (self.f[k] * EX).sum(-1) + 0.5 * self.F[k, 0]; ey = (self.f[k] * EY).sum(-1) + 0.5 * self.F[k, 1]
            wgt = 1. / self.tau[k] if self.relax == "SRT" else 1.0
            mx = mx + ex * wgt; my = my + ey * wgt; rt = rt + psi[k] * wgt
        with np.errstate(invalid="ignore", divide="ignore"):
            self.ueq = np.where(self.dom, np.stack([mx, my]) / rt, 0.)
        ux, uy = self.ueq
        eu = EX * ux[..., None] + EY * uy[..., None]
        poly = 1. + 3. * eu + 4.5 * (eu * eu) - 1.5 * (ux * ux + uy * uy)[..., None]
        self.feq = W * self.rho[..., None] * poly[None]                # calEquilibriumFuncEFGPU (227-248)
        with np.errstate(invalid="ignore", divide="ignore"):           # calForceDistrGPU (255-272)
            self.fF = (self.F[:, 0, ..., None] * (EX - ux[..., None]) + self.F[:, 1, ..., None] * (EY - uy[..., None])) * \
                self.feq / (1. / 3. * self.rho[..., None])
        self.fF = np.where(self.dom[None, :, :, None], self.fF, 0.)

    # neighbour slots of fillNeighboringNodesISO8 / ISO10 (ExplicitD2Q9GPU.py:392-592), (dx, dy)
    ISO_OFF = [(1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1),
               (2, 0), (0, 2), (-2, 0), (0, -2), (2, 2), (-2, 2), (-2, -2), (2, -2),
               (2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1), (-1, -2), (1, -2), (2, -1),
               (3, 0), (0, 3), (-3, 0), (0, -3), (3, 1), (1, 3), (-1, 3), (-3, 1), (-3, -1), (-1, -3), (1, -3), (3, -1)]
    ISO_W = {8: [4. / 21.] * 4 + [4. / 45.] * 4 + [1. / 60.] * 4 + [1. / 5040.] * 4 + [2. / 315.] * 8,
             10: [262. / 1785.] * 4 + [93. / 1190.] * 4 + [7. / 340.] * 4 + [9. / 9520.] * 4 + [6. / 595.] * 8 +
                 [2. / 5355.] * 4 + [1. / 7140.] * 8}        # ShanChenD2Q9.py:1677-1689
    KA = [0, 1, 1, 2, 2, 3, 3, 0]
    KB = [4, 4, 5, 5, 6, 6, 7, 7]

    def _force_iso(self, k):
        """calExplicit8thOrderScheme (ExplicitD2Q9GPU.py:627-953) / calExplicit10thOrderScheme (957-1372): a far
        neighbour contributes only if it is fluid and the nearer node(s) towards it are fluid; solids act through
        the first 8 slots with weights 1/9, 1/36; order 8 uses psi(x+e) - psi(x), order 10 plain psi(x+e)."""
        ns = 24 if self.scheme == 8 else 36
        w = self.ISO_W[self.scheme]
        sh = lambda a, d: np.roll(a, (-d[1], -d[0]), axis=(-2, -1))
        fl = [sh(self.dom, self.ISO_OFF[s]) for s in range(ns)]
        psi = self.rho
        fx = np.zeros((self.ny, self.nx)); fy = np.zeros((self.ny, self.nx))
        for s in range(ns):
            dx, dy = self.ISO_OFF[s]
            if s < 8:
                gate = np.ones_like(self.dom)
            elif s < 16:
                gate = fl[s - 8]
            elif s < 24:
                gate = fl[self.KA[s - 16]] | fl[self.KB[s - 16]]
            elif s < 28:
                gate = fl[s - 24] & fl[s - 16]
            else:
                t = s - 28
                gate = (fl[self.KA[t]] & fl[self.KA[t] + 8]) | (fl[self.KB[t]] & fl[16 + t])
            term = np.zeros((self.ny, self.nx))
            for j in range(self.nc):
                d = sh(psi[j], (dx, dy)) - psi[j] if self.scheme == 8 else sh(psi[j], (dx, dy))
                term = term + (-6.0 * w[s] * self.G[k, j]) * psi[k] * d
            on = fl[s] & gate
            fx = fx + np.where(on, term * dx, 0.); fy = fy + np.where(on, term * dy, 0.)
            if s < 8:
                ws = 1. / 9. if s < 4 else 1. / 36.
                sol = -ws * self.Gs[k] * psi[k]
                fx = fx + np.where(fl[s], 0., sol * dx); fy = fy + np.where(fl[s], 0., sol * dy)
        return fx, fy

    def _efs_prepare(self):
        """pre-loop, ShanChenD2Q9.py:1714-1849"""
        self._efs_force_ueq()
        self.f = self.f - 0.5 * self.fF                                # transformPDFGPU (278-287)
        rho_keep = self.rho.copy()
        self._inlet()
        if self.outlet == "Dirichlet":
            self._outlet_pressure()
        self.rho = rho_keep        # boundary densities set here are overwritten by calFluidRhoGPU before any use
        self.prepared = True

    def _efs_iteration(self):
        f_old = self.f.copy()                                          # savePDFLastStep
        d = self.feq - self.f - 0.5 * self.fF
        if self.relax == "SRT":                                        # calCollisionEXGPU (294-304)
            self.f = self.f + d / self.tau[:, None, None, None] + self.fF
        else:                                                          # 1379-1469
            self.f = self.f + np.stack([d[k] @ self.C[k].T for k in range(self.nc)]) + self.fF
        self.f = np.where(self.dom[None, :, :, None], self.f, 0.)
        self._stream()
        self._rho()
        self._uphys()
        if self.outlet == "Convective":                                # convectiveOutletEach{,2,3}GPU (1044-1119)
            v = np.abs(self.uph[1, 3])
            for row in (2, 1, 0):
                m = self.dom[row]
                new = (f_old[:, row] + v[None, :, None] * self.f[:, row + 1]) / (1. + v)[None, :, None]
                self.f[:, row] = np.where(m[None, :, None], new, self.f[:, row])
        elif self.outlet == "Dirichlet":
            self._outlet_pressure()
        self._inlet()
        self._rho()
        self._uphys()
        self._efs_force_ueq()

    def step(self, n=1):
        for _ in range(n):
            if self.model == "ShanChen":
                self._sc_iteration()
            else:
                if not self.prepared:
                    self._efs_prepare()
                self._efs_iteration()
