"""CPU ORACLE (test infrastructure, NOT product code) -- lattice-generic colour-gradient
CSF step on a DENSE masked grid, D2Q9 and D3Q19.

Second, independent restatement of the reference's colour-gradient loop
(/root/reference/RKCG2D/RKD2Q9.py:1295-1490 with the kernels of
/root/reference/RKCG2D/AcceleratedRKGPU2D.py) written with `np.roll` on dense
`[z, y, x]` arrays instead of the reference's compact node list, and
parameterised by the lattice so that the SAME code runs

  * D2Q9 -- pinned against the reference's golden vectors (tests/golden/cg2d_*.npz,
    periodic and solid/wetting cases) in tests/test_oracle_dense.py, and
  * D3Q19 -- for which the reference ships NO code (main.py:22 imports a module
    that is not in the repository).  The D3Q19 specification below is the
    generalisation fixed in SURVEY.md section 8 (a-3D) / DESIGN.md; its parity is
    "pinned by construction": identical code path as the D2Q9 instantiation
    that matches the reference, plus physical invariants (mass, symmetry,
    Laplace law) in the tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import it.
"""
import numpy as np


class Lattice:
    pass


def d2q9():
    """RKD2Q9.py:299-340."""
    L = Lattice()
    L.name = "D2Q9"; L.D = 2; L.Q = 9
    L.e = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0),
                    (1, 1, 0), (-1, 1, 0), (-1, -1, 0), (1, -1, 0)], dtype=np.int64)
    L.w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
    ex, ey = L.e[:, 0].astype(float), L.e[:, 1].astype(float)
    c2 = ex * ex + ey * ey
    L.M = np.stack([np.ones(9), -4 + 3 * c2, 4 - 10.5 * c2 + 4.5 * c2 * c2, ex, (-5 + 3 * c2) * ex,
                    ey, (-5 + 3 * c2) * ey, ex * ex - ey * ey, ex * ey])
    L.S_fixed = np.array([0., 1.64, 1.54, 0., 1.9, 0., 1.9, np.nan, np.nan])   # nan -> 1/tau
    # solid-normal stencil weights by |c|^2 (RKD2Q9.py:806-880)
    L.ns_weights = {1: 4. / 21., 2: 4. / 45., 4: 1. / 60., 5: 2. / 315., 8: 1. / 5040.}
    _finish(L)
    return L


def d3q19():
    """d'Humieres et al. 2002 velocity order and moment basis."""
    L = Lattice()
    L.name = "D3Q19"; L.D = 3; L.Q = 19
    L.e = np.array([(0, 0, 0),
                    (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
                    (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0),
                    (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1),
                    (0, 1, 1), (0, -1, 1), (0, 1, -1), (0, -1, -1)], dtype=np.int64)
    L.w = np.array([1. / 3.] + [1. / 18.] * 6 + [1. / 36.] * 12)
    ex, ey, ez = (L.e[:, k].astype(float) for k in range(3))
    c2 = ex * ex + ey * ey + ez * ez
    L.M = np.stack([
        np.ones(19), 19 * c2 - 30, (21 * c2 * c2 - 53 * c2 + 24) / 2,
        ex, (5 * c2 - 9) * ex, ey, (5 * c2 - 9) * ey, ez, (5 * c2 - 9) * ez,
        3 * ex * ex - c2, (3 * c2 - 5) * (3 * ex * ex - c2),
        ey * ey - ez * ez, (3 * c2 - 5) * (ey * ey - ez * ez),
        ex * ey, ey * ez, ex * ez,
        (ey * ey - ez * ez) * ex, (ez * ez - ex * ex) * ey, (ex * ex - ey * ey) * ez])
    n = np.nan   # -> 1/tau(phi): the five second-order (viscous-stress) moments
    L.S_fixed = np.array([0., 1.19, 1.4, 0., 1.2, 0., 1.2, 0., 1.2, n, 1.4, n, 1.4, n, n, n, 1.98, 1.98, 1.98])
    # 3-D 8th-order isotropic weights by |c|^2 (Sbragaglia et al. 2007), our choice for 3-D
    L.ns_weights = {1: 4. / 45., 2: 1. / 21., 3: 2. / 105., 4: 5. / 504., 5: 1. / 315., 6: 1. / 630., 8: 1. / 5040.}
    _finish(L)
    return L


def _finish(L):
    L.Mi = np.linalg.inv(L.M)
    L.opp = np.array([int(np.flatnonzero((L.e == -L.e[i]).all(1))[0]) for i in range(L.Q)])
    L.enorm = np.sqrt((L.e.astype(float) ** 2).sum(1))
    r = range(-2, 3)
    offs = []
    for cz in (r if L.D == 3 else (0,)):
        for cy in r:
            for cx in r:
                c2 = cx * cx + cy * cy + cz * cz
                if c2 in L.ns_weights:
                    offs.append((cx, cy, cz, L.ns_weights[c2]))
    L.ns_offsets = offs


def shift(a, e):
    """value at x + e (periodic) for a dense [z, y, x] array."""
    return np.roll(a, (-int(e[2]), -int(e[1]), -int(e[0])), axis=(-3, -2, -1))


class CGDense:
    def __init__(self, lattice, is_domain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98,
                 tauR=1.0, tauB=1.0, tautype=2, relax="MRT", inlet="Periodic", outlet="Periodic",
                 v_inlet=0.0, dBH=5e-8, dRH=1.0, dBL=1.0, dRL=5e-8):
        L = self.L = lattice
        self.inlet, self.outlet, self.v_in = inlet, outlet, v_inlet
        self.dBH, self.dRH, self.dBL, self.dRL = dBH, dRH, dBL, dRL
        dom = np.asarray(is_domain, bool)
        if dom.ndim == 2:
            dom = dom[None]
        self.dom = dom
        self.shape = dom.shape
        self.sigma, self.beta, self.delta = sigma, beta, delta
        self.cosT = np.cos(theta_deg / 180. * np.pi); self.sinT = np.sin(theta_deg / 180. * np.pi)
        self.wetting, self.tauR, self.tauB, self.tautype, self.relax = wetting, tauR, tauB, tautype, relax
        # geometry-derived fields
        nfl = sum(shift(dom, (cx, cy, cz)).astype(int)
                  for cz in ((-1, 0, 1) if L.D == 3 else (0,)) for cy in (-1, 0, 1) for cx in (-1, 0, 1))
        full = 27 if L.D == 3 else 9
        self.wet_solid = (~dom) & (nfl > 0)
        self.near_solid = dom & (nfl < full)
        ns = np.zeros((3,) + dom.shape)
        for cx, cy, cz, wgt in L.ns_offsets:
            sol = ~shift(dom, (cx, cy, cz))
            ns[0] += np.where(sol, wgt * cx, 0.); ns[1] += np.where(sol, wgt * cy, 0.); ns[2] += np.where(sol, wgt * cz, 0.)
        with np.errstate(invalid="ignore", divide="ignore"):
            nrm = np.sqrt((ns * ns).sum(0))
            self.ns = np.where(self.near_solid, ns / nrm, 0.)

    def set_densities(self, rhoR, rhoB):
        L = self.L
        rhoR = np.asarray(rhoR, float).reshape(self.shape) * self.dom
        rhoB = np.asarray(rhoB, float).reshape(self.shape) * self.dom
        self.rhoR, self.rhoB = rhoR.copy(), rhoB.copy()
        self.fR = L.w[:, None, None, None] * rhoR[None]
        self.fB = L.w[:, None, None, None] * rhoB[None]
        self.F = np.zeros((3,) + self.shape)
        self.u = np.zeros((3,) + self.shape)

    def set_pdfs(self, fR, fB):
        """fR, fB: [Q, z, y, x]; densities are the population sums."""
        self.fR = np.asarray(fR, float).copy() * self.dom
        self.fB = np.asarray(fB, float).copy() * self.dom
        self.rhoR = _qsum(self.fR); self.rhoB = _qsum(self.fB)
        self.F = np.zeros((3,) + self.shape); self.u = np.zeros((3,) + self.shape)

    # -- open boundaries: planes along the flow axis ("up" = y in 2-D, z in 3-D; inlet on top, outlet at the
    # bottom).  For D2Q9 these are the reference's row kernels term by term (AcceleratedRKGPU2D.py:2345-2423,
    # 923-961, 2557-2602, 604-784, 966-1080); for D3Q19 their Hecht-Harting generalisation: unknown = the
    # populations pointing into the domain, N_t = 1/2 sum_{in-plane} c_t f the transverse momentum.
    def _pl(self, a, k):
        """view of plane k (along the flow axis) of a [..., z, y, x] array"""
        return a[..., k, :, :] if self.L.D == 3 else a[..., 0, k, :][..., None, :]

    def _up(self, q):
        return int(self.L.e[q, self.L.D - 1])

    def _copy_plane(self, dst, src, sum_rho):
        m = self._pl(self.dom, dst) & self._pl(self.dom, src)
        for f in (self.fR, self.fB):
            self._pl(f, dst)[...] = np.where(m, self._pl(f, src), self._pl(f, dst))
        for f, r in ((self.fR, self.rhoR), (self.fB, self.rhoB)):
            new = _qsum(self._pl(f, dst)) if sum_rho else self._pl(r, src)
            self._pl(r, dst)[...] = np.where(m, new, self._pl(r, dst))

    def _inplane_sums(self, f, sign):
        L = self.L
        s0 = None; s1 = None; N = [0.0, 0.0]
        for q in range(L.Q):
            if self._up(q) == 0:
                s0 = f[q] if s0 is None else s0 + f[q]
                for t in range(L.D - 1):
                    if L.e[q, t] != 0:
                        N[t] = N[t] + float(L.e[q, t]) * f[q]
            if self._up(q) == sign:
                s1 = f[q] if s1 is None else s1 + f[q]
        return s0, s1, N

    def boundaries(self):
        L = self.L
        n_up = self.shape[0] if L.D == 3 else self.shape[1]
        zin, zgh = n_up - 2, n_up - 1
        with np.errstate(invalid="ignore", divide="ignore"):
            if self.inlet == "Neumann":
                m = self._pl(self.dom, zin)
                fR, fB = self._pl(self.fR, zin), self._pl(self.fB, zin)
                fT = fR + fB
                s0, sp, _ = self._inplane_sums(fT, 1)
                v = self.v_in
                rho = (s0 + 2. * sp) / (1. + v)
                rR, rB = self._pl(self.rhoR, zin), self._pl(self.rhoB, zin)
                ratioR = rR / (rR + rB)
                rRn = ratioR * rho
                ratioB = rB / (rRn + rB)                      # updated rhoR (2399-2407)
                rBn = ratioB * rho
                for q in range(1, L.Q):
                    if self._up(q) != -1:
                        continue
                    eq = rho * L.w[q] * (1. + 3. * (-v) + 4.5 * v * v - 1.5 * v * v)
                    eqo = rho * L.w[q] * (1. + 3. * v + 4.5 * v * v - 1.5 * v * v)
                    t = eq + (fT[L.opp[q]] - eqo)
                    fR[q] = np.where(m, ratioR * t, fR[q]); fB[q] = np.where(m, ratioB * t, fB[q])
                rR[...] = np.where(m, rRn, rR); rB[...] = np.where(m, rBn, rB)
                self._copy_plane(zgh, zin, True)
            elif self.inlet == "Dirichlet":
                m = self._pl(self.dom, zin)
                for f, r, p in ((self._pl(self.fB, zin), self._pl(self.rhoB, zin), self.dBH),
                                (self._pl(self.fR, zin), self._pl(self.rhoR, zin), self.dRH)):
                    s0, sp, N = self._inplane_sums(f, 1)
                    v = -1. + (s0 + 2. * sp) / p
                    old = f.copy()
                    for q in range(1, L.Q):
                        if self._up(q) != -1:
                            continue
                        eN = sum(float(L.e[q, t]) * N[t] for t in range(L.D - 1))
                        f[q] = np.where(m, old[L.opp[q]] + 0.5 * (-eN) - 6. * L.w[q] * p * v, old[q])
                    r[...] = np.where(m, p, r)
                self._copy_plane(zgh, zin, False)
            if self.outlet == "Convective":
                self._copy_plane(2, 3, True); self._copy_plane(1, 2, True); self._copy_plane(0, 1, True)
            elif self.outlet == "Dirichlet":
                m = self._pl(self.dom, 1)
                fR, fB = self._pl(self.fR, 1), self._pl(self.fB, 1)
                fT = fR + fB
                p = self.dBL + self.dRL
                s0, sm, N = self._inplane_sums(fT, -1)
                v = 1. - 1. / p * (s0 + 2. * sm)
                rR, rB = self._pl(self.rhoR, 1), self._pl(self.rhoB, 1)
                ratioR = rR / (rR + rB); ratioB = rB / (rR + rB)
                for q in range(1, L.Q):
                    if self._up(q) != 1:
                        continue
                    eN = sum(float(L.e[q, t]) * N[t] for t in range(L.D - 1))
                    t = fT[L.opp[q]] + 0.5 * (-eN) + 6. * L.w[q] * (p * v)
                    fR[q] = np.where(m, ratioR * t, fR[q]); fB[q] = np.where(m, ratioB * t, fB[q])
                self._copy_plane(0, 1, False)

    def head(self):
        L = self.L
        self.boundaries()
        self.fT = self.fR + self.fB
        with np.errstate(invalid="ignore", divide="ignore"):
            rho = self.rhoB + self.rhoR
            for a in range(3):
                mom = sum(float(L.e[i, a]) * self.fT[i] for i in range(L.Q) if L.e[i, a] != 0) \
                    if a < L.D else 0.0
                self.u[a] = np.where(self.dom, (mom + 0.5 * self.F[a]) / rho, 0.)
            self.phi = np.where(self.dom, (self.rhoR - self.rhoB) / (self.rhoR + self.rhoB), 0.)

    def _tau(self):
        phi = self.phi
        tau = np.ones(self.shape)
        tau[phi > self.delta] = self.tauR; tau[phi < -self.delta] = self.tauB
        mid = np.abs(phi) <= self.delta
        with np.errstate(invalid="ignore", divide="ignore"):
            if self.tautype == 1:
                t = 0.5 + 1. / ((1. + phi) / (2. * (self.tauR - 0.5)) + (1. - phi) / (2. * (self.tauB - 0.5)))
            else:
                xR = self.rhoR / (self.rhoR + self.rhoB); xB = self.rhoB / (self.rhoR + self.rhoB)
                t = 3. * (1. / (xR * (3. / (self.tauR - 0.5)) + xB * (3. / (self.tauB - 0.5)))) + 0.5
        tau[mid] = t[mid]
        return tau

    def gradient(self):
        """colour value on the wetting solids, G = 3 sum w_k e_k phi(x + e_k), wetting correction -> self.G (what body()
        evaluates first; a function of phi only, so callers that need G between head() and body() may ask for it)"""
        L = self.L; dom = self.dom; Q = L.Q
        ef = L.e.astype(float)
        with np.errstate(invalid="ignore", divide="ignore"):
            num = np.zeros(self.shape); den = np.zeros(self.shape)
            for k in range(1, Q):
                fl = shift(dom, L.e[k])
                num += np.where(fl, L.w[k] * shift(self.phi, L.e[k]), 0.); den += np.where(fl, L.w[k], 0.)
            phi_ext = np.where(dom, self.phi, np.where(self.wet_solid, num / den, 0.))
            G = np.zeros((3,) + self.shape)
            for k in range(1, Q):
                pk = shift(phi_ext, L.e[k])
                for a in range(L.D):
                    if L.e[k, a] != 0:
                        G[a] += L.w[k] * pk * ef[k, a]
            G *= 3.
            G = np.where(dom, G, 0.)
            if self.wet_solid.any():
                G = self._wetting(G)
        self.G = G
        return G

    def stream_only(self):
        """one streaming step with half-way bounce back + densities, no collision (drivers whose loop STARTS with the
        streaming: runTransport2DMPMCRKNew, Transport2DRK.py:1180-1200)"""
        L = self.L; dom = self.dom
        for name in ("fR", "fB"):
            f = getattr(self, name)
            new = np.empty_like(f)
            new[0] = f[0]
            for i in range(1, L.Q):
                src_fluid = shift(dom, -L.e[i])
                new[i] = np.where(src_fluid, shift(f[i], -L.e[i]), f[L.opp[i]])
            setattr(self, name, np.where(dom, new, 0.))
        self.fT = self.fR + self.fB
        self.rhoR = _qsum(self.fR); self.rhoB = _qsum(self.fB)

    def body(self):
        L = self.L; dom = self.dom; Q = L.Q
        ef = L.e.astype(float)
        with np.errstate(invalid="ignore", divide="ignore"):
            # colour value on wetting solids, then G
            num = np.zeros(self.shape); den = np.zeros(self.shape)
            for k in range(1, Q):
                fl = shift(dom, L.e[k])
                num += np.where(fl, L.w[k] * shift(self.phi, L.e[k]), 0.); den += np.where(fl, L.w[k], 0.)
            phi_ext = np.where(dom, self.phi, np.where(self.wet_solid, num / den, 0.))
            G = np.zeros((3,) + self.shape)
            for k in range(1, Q):
                pk = shift(phi_ext, L.e[k])
                for a in range(L.D):
                    if L.e[k, a] != 0:
                        G[a] += L.w[k] * pk * ef[k, a]
            G *= 3.
            G = np.where(dom, G, 0.)
            if self.wet_solid.any():
                G = self._wetting(G)
            self.G = G
            # curvature and CSF force
            gn = np.sqrt((G * G).sum(0))
            if self.wetting == 1:
                big = gn > 0.; sign = 1.
            else:
                big = gn > 1.0e-8; sign = -1.
            n = np.where(big & dom, sign * G / np.where(big, gn, 1.), 0.)
            dn = np.zeros((3, 3) + self.shape)        # dn[a, b] = d_a n_b
            for k in range(1, Q):
                nk = shift(n, L.e[k])                 # zero on solids = "fluid neighbours only"
                for a in range(L.D):
                    if L.e[k, a] != 0:
                        dn[a] += 3. * L.w[k] * ef[k, a] * nk
            # K = n_a n_b d_a n_b - (n.n) d_a n_a : identical to the reference's 2-D expression
            # (AcceleratedRKGPU2D.py:2547-2548), including K = 0 where |G| <= 1e-8 (n = 0)
            nn = (n * n).sum(0)
            K = np.zeros(self.shape)
            for a in range(L.D):
                for b in range(L.D):
                    K += n[a] * n[b] * dn[a, b]
                K -= nn * dn[a, a]
            self.K = K
            self.F = sign * 0.5 * self.sigma * K * G
            # collision of the total population
            tau = self._tau()
            u = self.u; rho = self.rhoR + self.rhoB
            uu = (u * u).sum(0)
            fe = np.empty((Q,) + self.shape)
            for i in range(Q):
                eu = sum(ef[i, a] * u[a] for a in range(L.D))
                fe[i] = rho * L.w[i] * (1. + (3. * eu + 4.5 * eu * eu - 1.5 * uu))
            fT = self.fT
            F = self.F
            if self.relax == "SRT":
                fT = -1. / tau * (fT - fe) + fT
                src = np.empty_like(fT)
                for i in range(Q):
                    eu = sum(ef[i, a] * u[a] for a in range(L.D))
                    src[i] = L.w[i] * sum((3. * (ef[i, a] - u[a]) + 9. * ef[i, a] * eu) * F[a]
                                          for a in range(L.D)) * (1. - 1. / (2. * tau))
                fT = fT + src
            else:
                S = np.where(np.isnan(L.S_fixed)[:, None, None, None], 1. / tau[None], L.S_fixed[:, None, None, None])
                m = np.tensordot(L.M, fT - fe, axes=(1, 0))
                fT = fT - np.tensordot(L.Mi, S * m, axes=(1, 0))
                src = np.empty_like(fT)
                for i in range(Q):
                    t = 0.
                    for a in range(L.D):
                        t = t + 3. * ef[i, a] * F[a]
                        for b in range(L.D):
                            t = t + 9. * (ef[i, a] * ef[i, b] - (1. / 3. if a == b else 0.)) * u[a] * F[b]
                    src[i] = L.w[i] * t
                ms = np.tensordot(L.M, src, axes=(1, 0))
                fT = fT + np.tensordot(L.Mi, (1. - 0.5 * S) * ms, axes=(1, 0))
            # recolouring
            tot = self.rhoR + self.rhoB
            fR = np.empty_like(fT); fB = np.empty_like(fT)
            for i in range(Q):
                ok = (gn > 1.0e-8) & (L.enorm[i] > 1.0e-8)
                cost = np.where(ok, sum(ef[i, a] * G[a] for a in range(L.D)) /
                                np.where(ok, L.enorm[i] * gn, 1.), 0.)
                a_ = self.beta * self.rhoR * self.rhoB / tot * L.w[i] * cost * L.enorm[i]
                fR[i] = self.rhoR / tot * fT[i] + a_
                fB[i] = self.rhoB / tot * fT[i] - a_
            fR = np.where(dom, fR, 0.); fB = np.where(dom, fB, 0.)
        # streaming with half-way bounce back (pull form of AcceleratedRKGPU2D.py:338-417)
        for name, f in (("fR", fR), ("fB", fB)):
            new = np.empty_like(f)
            new[0] = f[0]
            for i in range(1, Q):
                src_fluid = shift(dom, -L.e[i])
                new[i] = np.where(src_fluid, shift(f[i], -L.e[i]), f[L.opp[i]])
            setattr(self, name, np.where(dom, new, 0.))
        self.fT = self.fR + self.fB
        self.rhoR = _qsum(self.fR); self.rhoB = _qsum(self.fB)

    def step(self, n=1):
        for _ in range(n):
            self.head(); self.body()

    def _wetting(self, G):
        L = self.L
        near = self.near_solid
        ns = self.ns; c, s = self.cosT, self.sinT
        gn = np.sqrt((G * G).sum(0))
        big = gn > 1.0e-8
        if self.wetting == 1:
            assert L.D == 2, "WettingType 1 (Xu 2017) is a 2-D rotation; 3-D uses type 2"
            n1 = np.stack([ns[0] * c - ns[1] * s, ns[1] * c + ns[0] * s, 0 * ns[0]])
            n2 = np.stack([ns[0] * c + ns[1] * s, ns[1] * c - ns[0] * s, 0 * ns[0]])
            ug = np.where(big, G / np.where(big, gn, 1.), 0.)
            d1 = np.sqrt(((ug - n1) ** 2).sum(0)); d2 = np.sqrt(((ug - n2) ** 2).sum(0))
            mod = np.where(d1 < d2, n1, np.where(d1 > d2, n2, ns))
            return np.where(near, gn * mod, G)
        ug = np.where(big, -G / np.where(big, gn, 1.), 0.)
        dot = (ug * ns).sum(0)
        th = np.arccos(np.clip(dot, -1., 1.))
        sth = np.sin(th); cth = np.cos(th)
        ok = np.abs(sth) > 1.0e-9
        sth_ = np.where(ok, sth, 1.)
        c1 = np.where(ok, s * cth / sth_, 0.); c2 = np.where(ok, s / sth_, 0.)
        n1 = (c - c1) * ns + c2 * ug
        n2 = (c + c1) * ns - c2 * ug
        d1 = np.sqrt(((n1 - ug) ** 2).sum(0)); d2 = np.sqrt(((n2 - ug) ** 2).sum(0))
        new = np.where(d1 < d2, -gn * n1, np.where(d1 > d2, -gn * n2, G))
        return np.where(near, new, G)


def _qsum(f):
    s = f[0].copy()
    for k in range(1, f.shape[0]):
        s = s + f[k]
    return s
