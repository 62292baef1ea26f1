"""CPU ORACLE (test infrastructure, NOT product code) -- colour-gradient step with the PERTURBATION surface-tension
operator on a dense masked grid, lattice-generic (D2Q9 and D3Q19).

Restates the kernel sequence of the reference's `runRKColorGradient2DPerturbation`, MRT branch
(/root/reference/RKCG2D/RKD2Q9.py:1046-1219), the one reading under which that work-in-progress driver is
self-consistent (tests/golden/gen_goldens_cgp2d.py explains what is left out and why):
    streaming with half-way bounce back           calStreaming1GPU / calStreaming2GPU   AcceleratedRKGPU2D.py:338-417
    fT = fR + fB                                   calTotalFluidPDF                      :1413-1422
    rho_R, rho_B; u = sum e (fR + fB) / rho        calMacroDensityRKGPU2D :101-118, calPhysicalVelocityRKGPU2D :125-146
    phi = (rho_R - rho_B) / (rho_R + rho_B)        calPhaseFieldPhi                      :1347-1356
    MRT collision of fT, tau harmonic in phi,      calRKCollision1GPU2DMRTNew            :1272-1343
      body force w_F e.F with w_F = 3 w_i
    G = 3 sum w_k e_k phi(x + e_k) with phi =      calRKCollision23GPUNew                :1169-1266
      SolidColorDiff on solid neighbours;
      fT_i += (A_R + A_B)/2 |G| (w_i (e_i.G)^2/|G|^2 - B_i);
      fR_i = rho_R/rho fT_i + beta rho_R rho_B/rho^2 w_i cos(theta_i), fB_i = rho_B/rho fT_i - ...
The D2Q9 instantiation is PINNED against tests/golden/cgp2d_*.npz (the reference's kernels executed verbatim under
Numba's CUDA simulator) in tests/test_oracle_cgp.py.  The D3Q19 instantiation is the direct generalisation --
B = (-1/3, 1/18 x 6, 1/36 x 12) (Liu et al. 2012; B_0 = w_0 - 2/3 like the reference's D2Q9 set (-2/9, 1/9, 1/36),
RKD2Q9.py:131-133), d'Humieres MRT basis with the rates of the CSF specification -- which is what the reference's 3-D
ini parameterises (IniFiles/RKtwophasesetup3D.ini: AkR, AkB, no [SurfaceTension] section); pinned by construction +
invariants (mass, Laplace law).
Only tests/ and __graft_entry__.smoke() may import it.
"""
import numpy as np

from . import cg_dense
from .cg_dense import shift, _qsum


def constant_B(L):
    B = L.w.copy()
    B[0] = L.w[0] - 2. / 3.
    return B


class CGPDense:
    def __init__(self, lattice, is_domain, beta=0.7, AkR=1.4e-2, AkB=1.4e-2, tauR=1.0, tauB=1.0, solid_phi=0.5,
                 body_force=(0., 0., 0.), relax="MRT", **open_boundaries):
        L = self.L = lattice
        dom = np.asarray(is_domain, bool)
        self.dom = dom[None] if dom.ndim == 2 else dom
        self.shape = self.dom.shape
        self.beta, self.AkR, self.AkB, self.tauR, self.tauB, self.solid_phi = beta, AkR, AkB, tauR, tauB, solid_phi
        self.bf = np.zeros(3); self.bf[:len(body_force)] = body_force
        self.B = constant_B(L)
        self.ef = L.e.astype(float)
        self.relax = relax      # "SRT": relaxation of the total population with tau(phi) = the sum of the reference's two per-colour
                                # SRT collisions (calRKCollision1GPU2DSRTNew), whose result its driver drops
        # Open boundaries (inlet=, outlet=, v_inlet=, dBH=, dRH=, dBL=, dRL= as for cg_dense.CGDense).  The reference's driver
        # treats the open rows AFTER forming the total population (RKD2Q9.py:1063 vs :1065-1118), so their effect never
        # reaches its collision; the consistent order -- the one of its CSF loop -- is taken here: streaming, densities, the
        # row operators of the CSF loop (cg_dense.CGDense.boundaries), velocity, phi, collision.  This is what the
        # reference's 3-D ini asks for (perturbation parameters + velocity inlet + pressure outlet); no reference vector
        # exists for it.
        self.bc = cg_dense.CGDense(L, self.dom, **open_boundaries) if open_boundaries else None

    def set_densities(self, rhoR, rhoB):
        """f = w rho at rest, then the streaming the reference's loop starts with: the state is its first output"""
        L = self.L
        rhoR = np.asarray(rhoR, float).reshape(self.shape) * self.dom
        rhoB = np.asarray(rhoB, float).reshape(self.shape) * self.dom
        self.fR = L.w[:, None, None, None] * rhoR[None]
        self.fB = L.w[:, None, None, None] * rhoB[None]
        self._stream_and_macros()

    def _stream_and_macros(self):
        L = self.L; dom = self.dom
        for name in ("fR", "fB"):
            f = getattr(self, name)
            new = np.empty_like(f)
            new[0] = f[0]
            for i in range(1, L.Q):
                src_fluid = shift(dom, -L.e[i])
                new[i] = np.where(src_fluid, shift(f[i], -L.e[i]), f[L.opp[i]])
            setattr(self, name, np.where(dom, new, 0.))
        self.rhoR = _qsum(self.fR); self.rhoB = _qsum(self.fB)
        if self.bc is not None:
            b = self.bc
            b.fR, b.fB, b.rhoR, b.rhoB = self.fR, self.fB, self.rhoR, self.rhoB
            b.boundaries()
            self.fR, self.fB, self.rhoR, self.rhoB = b.fR, b.fB, b.rhoR, b.rhoB
        with np.errstate(invalid="ignore", divide="ignore"):
            rho = self.rhoB + self.rhoR
            self.u = np.zeros((3,) + self.shape)
            for a in range(L.D):
                m = np.zeros(self.shape)
                for f in (self.fR, self.fB):
                    for q in range(1, L.Q):
                        if L.e[q, a] != 0:
                            m = m + self.ef[q, a] * f[q]
                self.u[a] = np.where(dom, m / rho, 0.)
            self.phi = np.where(dom, (self.rhoR - self.rhoB) / (self.rhoR + self.rhoB), 0.)

    def step(self, n=1):
        L = self.L; dom = self.dom; Q = L.Q; ef = self.ef
        for _ in range(n):
            with np.errstate(invalid="ignore", divide="ignore"):
                fT = self.fR + self.fB
                phi = self.phi
                tau = 0.5 + 1. / ((1. + phi) / (2. * (self.tauR - 0.5)) + (1. - phi) / (2. * (self.tauB - 0.5)))
                rho = self.rhoB + self.rhoR
                u = self.u
                uu = (u * u).sum(0)
                fe = np.empty((Q,) + self.shape)
                for i in range(Q):
                    eu = sum(ef[i, a] * u[a] for a in range(L.D))
                    fe[i] = rho * L.w[i] * (1. + (3. * eu + 4.5 * eu * eu - 1.5 * uu))
                S = np.where(np.isnan(L.S_fixed)[:, None, None, None], 1. / tau[None], L.S_fixed[:, None, None, None])
                m = np.tensordot(L.M, fT, axes=(1, 0)) - np.tensordot(L.M, fe, axes=(1, 0))
                wF = 3. * L.w; wF[0] = 0.
                force = np.stack([wF[i] * sum(ef[i, a] * self.bf[a] for a in range(L.D)) * np.ones(self.shape) for i in range(Q)])
                if self.relax == "SRT":
                    fT = -(1. / tau) * (fT - fe) + force + fT
                else:
                    fT = -np.tensordot(L.Mi, S * m, axes=(1, 0)) + force + fT
                # colour gradient: phi of the fluid neighbours, SolidColorDiff on solid ones
                G = np.zeros((3,) + self.shape)
                for k in range(1, Q):
                    pk = np.where(shift(dom, L.e[k]), shift(phi, L.e[k]), self.solid_phi)
                    for a in range(L.D):
                        if L.e[k, a] != 0:
                            G[a] = G[a] + 3. * L.w[k] * ef[k, a] * pk
                g2 = (G * G).sum(0); gn = np.sqrt(g2)
                nz = g2 != 0.
                for i in range(Q):
                    eg = sum(ef[i, a] * G[a] for a in range(L.D))
                    part = L.w[i] * (eg * eg) / np.where(nz, g2, 1.)
                    fT[i] = fT[i] + np.where(nz, (self.AkR + self.AkB) * 0.5 * gn * (part - self.B[i]), 0.)
                rmul = self.rhoR * self.rhoB; r2 = rho * rho
                fR = np.empty_like(fT); fB = np.empty_like(fT)
                for i in range(Q):
                    ok = nz & (L.enorm[i] != 0.)
                    cost = np.where(ok, sum(ef[i, a] * G[a] for a in range(L.D)) / np.where(ok, L.enorm[i] * gn, 1.), 0.)
                    a_ = (self.beta * rmul / r2) * L.w[i] * cost
                    fR[i] = self.rhoR / rho * fT[i] + a_
                    fB[i] = self.rhoB / rho * fT[i] - a_
                self.fR = np.where(dom, fR, 0.); self.fB = np.where(dom, fB, 0.)
                self.G = np.where(dom, G, 0.)
            self._stream_and_macros()


d2q9, d3q19 = cg_dense.d2q9, cg_dense.d3q19
