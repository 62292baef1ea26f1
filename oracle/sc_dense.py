"""CPU ORACLE (test infrastructure, NOT product code) -- lattice-generic Shan-Chen step on a dense masked grid.

Second restatement of the two live Shan-Chen loops of openLBMPM (ShanChen2D/ShanChenD2Q9.py:1492-1629 original
Shan-Chen; :1714-2087 explicit forcing SRT/MRT, isotropy 4; kernels in OptimizedD2Q9GPU.py / ExplicitD2Q9GPU.py),
written once for any lattice on `[z, y, x]` arrays (periodic + half-way bounce back; open boundaries along the flow
axis: per-component Zou-He velocity inlet on top, Zou-He pressure or convective outlet at the bottom -- for D2Q9 the
reference's row kernels term by term, OptimizedD2Q9GPU.py:555-584, 710-768, 839-861, 960-1119; for D3Q19 their
Hecht-Harting generalisation: unknown = the populations pointing into the domain, N_t = the in-plane transverse
momentum), so that the SAME code runs
  * D2Q9  -- pinned against the reference's golden vectors (tests/golden/sc2d_*.npz, periodic and solid cases) in
             tests/test_oracle_sc_dense.py, and
  * D3Q19 -- for which the reference ships no code (`ShanChenD3Q19` is imported by main.py:17 but absent).  The 3-D
             specification is the direct generalisation: interaction weights 3 w_k (explicit forcing) / w_k (original),
             solid weights w_k, c_s^2 = 1/3, d'Humieres moment basis with rates 1 (conserved), 1.19, 1.4, 1.2, 1.4,
             1.98 and 1/tau on the five stress moments.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import it.
"""
import numpy as np

from . import cg_dense


def shift(a, e):
    """value at x + e (periodic) for a dense [..., z, y, x] array"""
    return np.roll(a, (-int(e[2]), -int(e[1]), -int(e[0])), axis=(-3, -2, -1))


class SCDense:
    def __init__(self, lattice, is_domain, model="EFS", relax="SRT", tau=(1., 1.), G=0.2, Gs=(-0.14, 0.14),
                 inlet="Periodic", outlet="Periodic", v_in=(0., 0.), rho_out=(1.0, 0.02)):
        L = self.L = lattice
        dom = np.asarray(is_domain, bool)
        self.dom = dom[None] if dom.ndim == 2 else dom
        self.shape = self.dom.shape
        self.model, self.relax = model, relax
        self.tau = np.asarray(tau, float); self.nc = self.tau.size
        self.G = np.zeros((self.nc, self.nc)); self.G[0, 1] = self.G[1, 0] = G
        self.Gs = np.asarray(Gs, float)
        self.ef = L.e.astype(float)
        self.inlet, self.outlet = inlet, outlet
        self.v_in = np.asarray(v_in, float); self.rho_out = np.asarray(rho_out, float)
        self.nflow = self.shape[0] if L.D == 3 else self.shape[1]
        self.z_in, self.z_out = self.nflow - 2, 1
        self.up = L.e[:, L.D - 1]                       # component along the flow axis (y in 2-D, z in 3-D)
        if relax == "MRT":
            if L.Q == 9:      # ShanChenD2Q9.py:99-106
                base = np.array([1., 0.6, 1.5, 1., 1.2, 1., 1.2, np.nan, np.nan])
            else:             # 3-D: conserved 1, the d'Humieres rates of the colour-gradient spec, stress moments 1/tau
                n = np.nan
                base = np.array([1., 1.19, 1.4, 1., 1.2, 1., 1.2, 1., 1.2, n, 1.4, n, 1.4, n, n, n, 1.98, 1.98, 1.98])
            self.C = []
            for k in range(self.nc):
                s = np.where(np.isnan(base), 1. / self.tau[k], base)
                if k >= 2:
                    s = np.where(np.isnan(base), 1. / self.tau[k], 1.0)
                self.C.append(L.Mi @ np.diag(s) @ L.M)
        self.prepared = False

    def set_densities(self, rho):
        self.rho = np.where(self.dom, np.asarray(rho, float).reshape((self.nc,) + self.shape), 0.)
        self.f = self.L.w[None, :, None, None, None] * self.rho[:, None]          # [nc, Q, z, y, x]
        self.F = np.zeros((self.nc, 3) + self.shape)
        self.uph = np.zeros((3,) + self.shape)
        self.prepared = False

    def _rho(self):
        s = self.f[:, 0].copy()
        for q in range(1, self.L.Q):
            s = s + self.f[:, q]
        self.rho = np.where(self.dom, s, 0.)

    def _stream(self):
        L = self.L
        new = self.f.copy()
        for q in range(1, L.Q):
            src_fluid = shift(self.dom, -L.e[q])
            new[:, q] = np.where(src_fluid, shift(self.f[:, q], -L.e[q]), self.f[:, L.opp[q]])
        self.f = np.where(self.dom, new, 0.)

    def _momentum(self, k):
        return [sum(self.ef[q, a] * self.f[k, q] for q in range(self.L.Q)) for a in range(3)]

    def _uphys(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            r = sum(self.rho[k] for k in range(self.nc))
            for a in range(self.L.D):
                v = sum(self._momentum(k)[a] + 0.5 * self.F[k, a] for k in range(self.nc))
                self.uph[a] = np.where(self.dom, v / r, 0.)

    # -- open boundaries: planes along the flow axis ----------------------------------------------------------
    def _ix(self, r):
        return (r, slice(None), slice(None)) if self.L.D == 3 else (0, r, slice(None))

    def _row_copy(self, dst, src):
        """ghostPointsConstantVelocityInlet / ...PressureOutlet / convectiveOutletGPU...: plane dst <- plane src, rho = sum"""
        d, s_ = self._ix(dst), self._ix(src)
        m = self.dom[d] & self.dom[s_]
        for k in range(self.nc):
            acc = None
            for q in range(self.L.Q):
                v = np.where(m, self.f[k, q][s_], self.f[k, q][d])
                self.f[k, q][d] = v
                acc = v.copy() if acc is None else acc + v
            self.rho[k][d] = np.where(m, acc, self.rho[k][d])

    def _plane_sums(self, k, ix, sign):
        """s0 = sum of the in-plane populations, s1 = sum of those with e_up = sign, N_t = in-plane transverse momentum"""
        L = self.L
        s0 = s1 = None
        N = [0., 0.]
        for q in range(L.Q):
            fq = self.f[k, q][ix]
            if self.up[q] == 0:
                s0 = fq.copy() if s0 is None else s0 + fq
                for a in range(L.D - 1):
                    if L.e[q, a] != 0:
                        N[a] = N[a] + self.ef[q, a] * fq
            elif self.up[q] == sign:
                s1 = fq.copy() if s1 is None else s1 + fq
        return s0, s1, N

    def _inlet(self):
        """constantVelocityZouHeBoundaryHigher (OptimizedD2Q9GPU.py:839-861) + ghost plane (710-736)"""
        if self.inlet != "Neumann":
            return
        L = self.L; ix = self._ix(self.z_in); m = self.dom[ix]
        for k in range(self.nc):
            v = self.v_in[k]
            s0, sp, N = self._plane_sums(k, ix, +1)
            rho = (s0 + 2. * sp) / (1. + v)
            self.rho[k][ix] = np.where(m, rho, 0.)
            old = [self.f[k, q][ix].copy() for q in range(L.Q)]
            for q in range(1, L.Q):
                if self.up[q] != -1:
                    continue
                val = old[L.opp[q]] - 0.5 * sum(self.ef[q, a] * N[a] for a in range(L.D - 1)) - 6. * L.w[q] * rho * v
                self.f[k, q][ix] = np.where(m, val, old[q])
        self._row_copy(self.z_in + 1, self.z_in)

    def _outlet_pressure(self):
        """constantPressureZouHeBoundaryLower (555-584, the hard-coded densities arrive as rho_out) + ghost plane (743-768)"""
        L = self.L; ix = self._ix(self.z_out); m = self.dom[ix]
        for k in range(self.nc):
            d = self.rho_out[k]
            s0, sm, N = self._plane_sums(k, ix, -1)
            vy = 1. - (s0 + 2. * sm) / d
            old = [self.f[k, q][ix].copy() for q in range(L.Q)]
            for q in range(1, L.Q):
                if self.up[q] != 1:
                    continue
                val = old[L.opp[q]] - 0.5 * sum(self.ef[q, a] * N[a] for a in range(L.D - 1)) + 6. * L.w[q] * d * vy
                self.f[k, q][ix] = np.where(m, val, old[q])
            self.rho[k][ix] = np.where(m, d, 0.)
        self._row_copy(self.z_out - 1, self.z_out)

    # -- original Shan-Chen: interactionCollisionProcess (OptimizedD2Q9GPU.py:1274-1446) ----------------------
    def _sc_iteration(self):
        L = self.L
        self._inlet()
        self._rho()
        psi = self.rho
        with np.errstate(invalid="ignore", divide="ignore"):
            rt = sum(self.rho[k] / self.tau[k] for k in range(self.nc))
            up = [sum(self._momentum(k)[a] / self.tau[k] for k in range(self.nc)) / rt for a in range(3)]
            for k in range(self.nc):
                F = [np.zeros(self.shape) for _ in range(3)]
                for q in range(1, L.Q):
                    fl = shift(self.dom, L.e[q])
                    t_f = sum(-L.w[q] * self.G[k, j] * psi[k] * shift(psi[j], L.e[q]) for j in range(self.nc))
                    t = np.where(fl, t_f, -L.w[q] * self.Gs[k] * psi[k])
                    for a in range(L.D):
                        if L.e[q, a] != 0:
                            F[a] = F[a] + t * self.ef[q, a]
                for a in range(3):
                    self.F[k, a] = np.where(self.dom, F[a], 0.)
                u = [up[a] + self.tau[k] * F[a] / psi[k] for a in range(3)]
                uu = sum(u[a] * u[a] for a in range(L.D))
                for q in range(L.Q):
                    eu = sum(self.ef[q, a] * u[a] for a in range(L.D))
                    self.f[k, q] = (1. - 1. / self.tau[k]) * self.f[k, q] + L.w[q] * psi[k] / self.tau[k] * \
                        (1. + 3. * eu + 4.5 * (eu * eu) - 1.5 * uu)
        self.f = np.where(self.dom, self.f, 0.)
        self._stream()
        self._rho()
        if self.outlet == "Convective":                  # convectiveOutletGPU / Ghost2 / Ghost3 (960-1036)
            self._row_copy(2, 3); self._row_copy(1, 2); self._row_copy(0, 1)
        self._uphys()

    # -- explicit forcing (ExplicitD2Q9GPU.py:51-363, 1379-1469) ---------------------------------------------
    def _efs_force_ueq(self):
        L = self.L
        psi = self.rho
        m = [np.zeros(self.shape) for _ in range(3)]; rt = np.zeros(self.shape)
        for k in range(self.nc):
            g = [np.zeros(self.shape) for _ in range(3)]; sl = [np.zeros(self.shape) for _ in range(3)]
            for q in range(1, L.Q):
                wI = 3. * L.w[q]
                fl = shift(self.dom, L.e[q])
                grad = sum(wI * (shift(psi[j], L.e[q]) - psi[j]) * self.G[k, j] for j in range(self.nc))
                sol = -wI * self.Gs[k] * psi[k]
                for a in range(L.D):
                    if L.e[q, a] != 0:
                        g[a] = g[a] + np.where(fl, grad * self.ef[q, a], 0.)
                        sl[a] = sl[a] + np.where(fl, 0., sol * self.ef[q, a])
            mom = self._momentum(k)
            wgt = 1. / self.tau[k] if self.relax == "SRT" else 1.0
            for a in range(3):
                self.F[k, a] = np.where(self.dom, -6.0 * psi[k] * g[a] + sl[a], 0.)
                m[a] = m[a] + (mom[a] + 0.5 * self.F[k, a]) * wgt
            rt = rt + psi[k] * wgt
        with np.errstate(invalid="ignore", divide="ignore"):
            self.ueq = np.where(self.dom, np.stack(m) / rt, 0.)
            u = self.ueq
            uu = sum(u[a] * u[a] for a in range(L.D))
            self.feq = np.empty_like(self.f); self.fF = np.empty_like(self.f)
            for q in range(L.Q):
                eu = sum(self.ef[q, a] * u[a] for a in range(L.D))
                for k in range(self.nc):
                    self.feq[k, q] = L.w[q] * self.rho[k] * (1. + 3. * eu + 4.5 * (eu * eu) - 1.5 * uu)
                    self.fF[k, q] = sum(self.F[k, a] * (self.ef[q, a] - u[a]) for a in range(L.D)) * self.feq[k, q] / \
                        (1. / 3. * self.rho[k])
        self.fF = np.where(self.dom, self.fF, 0.)

    def _efs_iteration(self):
        if not self.prepared:                            # pre-loop, ShanChenD2Q9.py:1714-1849
            self._efs_force_ueq()
            self.f = self.f - 0.5 * self.fF
            keep = self.rho.copy()
            self._inlet()
            if self.outlet == "Dirichlet":
                self._outlet_pressure()
            self.rho = keep          # boundary densities set here are overwritten by calFluidRhoGPU before any use
            self.prepared = True
        f_old = self.f.copy()                            # savePDFLastStep
        d = self.feq - self.f - 0.5 * self.fF
        if self.relax == "SRT":
            self.f = self.f + d / self.tau[:, None, None, None, None] + self.fF
        else:
            self.f = self.f + np.stack([np.tensordot(self.C[k], d[k], axes=(1, 0)) for k in range(self.nc)]) + self.fF
        self.f = np.where(self.dom, self.f, 0.)
        self._stream()
        self._rho()
        self._uphys()
        if self.outlet == "Convective":                  # convectiveOutletEach{,2,3}GPU (OptimizedD2Q9GPU.py:1044-1119)
            v = np.abs(self.uph[self.L.D - 1][self._ix(3)])
            for row in (2, 1, 0):
                ix, ixn = self._ix(row), self._ix(row + 1)
                m = self.dom[ix]
                for k in range(self.nc):
                    for q in range(self.L.Q):
                        new = (f_old[k, q][ix] + v * self.f[k, q][ixn]) / (1. + v)
                        self.f[k, q][ix] = np.where(m, new, self.f[k, q][ix])
        elif self.outlet == "Dirichlet":
            self._outlet_pressure()
        self._inlet()
        if self.inlet != "Periodic" or self.outlet != "Periodic":
            self._rho()
        self._uphys()
        self._efs_force_ueq()

    def step(self, n=1):
        for _ in range(n):
            if self.model == "ShanChen":
                self._sc_iteration()
            else:
                self._efs_iteration()


d2q9, d3q19 = cg_dense.d2q9, cg_dense.d3q19
