"""CPU ORACLE (test infrastructure, NOT product code) -- lattice-generic Shan-Chen step on a dense masked grid.

Second restatement of the two live Shan-Chen loops of openLBMPM (ShanChen2D/ShanChenD2Q9.py:1492-1629 original
Shan-Chen; :1714-2087 explicit forcing SRT/MRT, isotropy 4; kernels in OptimizedD2Q9GPU.py / ExplicitD2Q9GPU.py),
written once for any lattice on `[z, y, x]` arrays (closed boxes: periodic + half-way bounce back), so that the
SAME code runs
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
    def __init__(self, lattice, is_domain, model="EFS", relax="SRT", tau=(1., 1.), G=0.2, Gs=(-0.14, 0.14)):
        L = self.L = lattice
        dom = np.asarray(is_domain, bool)
        self.dom = dom[None] if dom.ndim == 2 else dom
        self.shape = self.dom.shape
        self.model, self.relax = model, relax
        self.tau = np.asarray(tau, float); self.nc = self.tau.size
        self.G = np.zeros((self.nc, self.nc)); self.G[0, 1] = self.G[1, 0] = G
        self.Gs = np.asarray(Gs, float)
        self.ef = L.e.astype(float)
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

    # -- original Shan-Chen: interactionCollisionProcess (OptimizedD2Q9GPU.py:1274-1446) ----------------------
    def _sc_iteration(self):
        L = self.L
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
        if not self.prepared:
            self._efs_force_ueq()
            self.f = self.f - 0.5 * self.fF
            self.prepared = True
        d = self.feq - self.f - 0.5 * self.fF
        if self.relax == "SRT":
            self.f = self.f + d / self.tau[:, None, None, None, None] + self.fF
        else:
            self.f = self.f + np.stack([np.tensordot(self.C[k], d[k], axes=(1, 0)) for k in range(self.nc)]) + self.fF
        self.f = np.where(self.dom, self.f, 0.)
        self._stream()
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
