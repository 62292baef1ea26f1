"""CPU ORACLE (test infrastructure, NOT product code) -- solute tracers riding on the colour-gradient CSF flow
(SURVEY.md section 8, row f-3), lattice-generic on a dense masked grid.

Restates the tracer part of the reference's `runTransport2DMPMCRKNew`, NumberSchemes = 9
(/root/reference/RKCG2D/Transport2DRK.py:1341-1425; kernels in AccelerateTransport2DRK.py), in the driver's order:
    loop:  flow streaming, densities, velocity (lagged force), phi, colour gradient          cg_dense.CGDense
           indicator = -1 where rho_R <= criterion else 0                                    calValueTransportDomain :957-971
           SRT  g_j <- g_j - (g_j - C w_j (1 + 3 e_j.u)) / tau                               calCollisionQ9 :704-730
           MRT  g <- g - M^-1 S^-1 (M g - M g_eq), S = 1 except the flux moments (3, 4: 1/2 + 3 D_xx; 5, 6: 1/2 + 3 D_yy;
                S[3,5] = 3 D_xy, S[5,3] = 3 D_yx)                                            calCollisionTransportLinearEqlMRTGPUD2Q9 :1053-1105, Transport2DRK.py:367-391
           g_j += beta * indicator * w_j C cos(angle(e_j, -G)), j > 0, |G| > 1e-8            calTransportWithInterfaceD2Q9 :1019-1047
           streaming with half-way bounce back                                               calStreaming1GPU / 2GPU :736-835
           C = sum_j g_j                                                                     calConcentrationGPU :78-90
           flow force, collision, recolouring
The D2Q9 instantiation is PINNED against tests/golden/tr2d_*.npz (the reference's kernels executed verbatim,
tests/golden/gen_goldens_tr2d.py; all-fluid periodic boxes -- next to wetting solids the reference's tracer streaming
writes through a negative index, here it bounces back).  D3Q19: SRT by generalisation.
Only tests/ may import it.
"""
import numpy as np

from . import cg_dense
from .cg_dense import shift, _qsum


class TracerDense:
    def __init__(self, flow, relax="SRT", tau=(0.8,), dxx=(0.05,), dyy=(0.08,), dxy=(0.0,), dyx=(0.0,), beta=(0.6,), criterion=0.5):
        self.flow = flow                                  # cg_dense.CGDense (CSF), closed box
        L = self.L = flow.L
        self.relax, self.criterion = relax, criterion
        self.tau = np.asarray(tau, float); self.beta = np.asarray(beta, float)
        self.nt = self.tau.size
        if relax == "MRT":
            assert L.Q == 9, "the reference defines the tracer MRT for D2Q9 only"
            self.A = []
            for i in range(self.nt):
                S = np.eye(9)
                S[3, 3] = S[4, 4] = 0.5 + 3. * dxx[i]; S[5, 5] = S[6, 6] = 0.5 + 3. * dyy[i]
                S[3, 5] = 3. * dxy[i]; S[5, 3] = 3. * dyx[i]
                self.A.append(-np.dot(L.Mi, np.linalg.inv(S)))

    def set_concentrations(self, conc):
        """call after flow.set_densities: g = w C at rest; the flow streams once (the loop starts with the streaming)"""
        f = self.flow
        self.conc = np.asarray(conc, float).reshape((self.nt,) + f.shape) * f.dom
        self.g = self.L.w[None, :, None, None, None] * self.conc[:, None]
        f.stream_only()

    def step(self, n=1):
        f = self.flow; L = self.L; dom = f.dom; ef = L.e.astype(float)
        for _ in range(n):
            f.head()
            G = f.gradient()
            with np.errstate(invalid="ignore", divide="ignore"):
                value = np.where(f.rhoR > self.criterion, -0.0, -1.0)
                gn = np.sqrt((G * G).sum(0))
                big = gn > 1.0e-8
                ug = np.where(big, -G / np.where(big, gn, 1.), 0.)
                un = np.sqrt((ug * ug).sum(0))
                for i in range(self.nt):
                    C = self.conc[i]
                    geq = np.stack([C * L.w[j] * (1. + 3. * sum(ef[j, a] * f.u[a] for a in range(L.D))) for j in range(L.Q)])
                    if self.relax == "SRT":
                        self.g[i] = -(self.g[i] - geq) / self.tau[i] + self.g[i]
                    else:
                        diff = np.tensordot(L.M, self.g[i], axes=(1, 0)) - np.tensordot(L.M, geq, axes=(1, 0))
                        self.g[i] = self.g[i] + np.tensordot(self.A[i], diff, axes=(1, 0))
                    for j in range(1, L.Q):
                        ok = (un > 1.0e-8) & (L.enorm[j] > 1.0e-8)
                        cos = np.where(ok, sum(ef[j, a] * ug[a] for a in range(L.D)) / np.where(ok, L.enorm[j] * un, 1.), 0.)
                        self.g[i, j] = self.g[i, j] + self.beta[i] * value * (L.w[j] * C) * cos
                    new = np.empty_like(self.g[i])
                    new[0] = self.g[i, 0]
                    for j in range(1, L.Q):
                        src_fluid = shift(dom, -L.e[j])
                        new[j] = np.where(src_fluid, shift(self.g[i, j], -L.e[j]), self.g[i, L.opp[j]])
                    self.g[i] = np.where(dom, new, 0.)
                    c = np.zeros(f.shape)
                    for j in range(L.Q):
                        c = c + self.g[i, j]
                    self.conc[i] = c
            f.body()
