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
NumberSchemes = 5 (`TracerDenseQ5`, branch :1344-1384 of the same driver; D2Q5 tracers on the D2Q9 flow):
           MRT  g <- g - M^-1 S^-1 (M g - M g_eq), g_eq = C w_j (1 + 3 e_j.u), w = (1/3, 1/6 x4), rows of M: (1,1,1,1,1),
                j_x, j_y, (4,-1,-1,-1,-1), (0,1,1,-1,-1); S = 1 except the flux block             calCollisionTransportLinearEqlMRTGPU :535-590, Transport2DRK.py:313-347
           g_j += beta * indicator * w_j C cos(angle(e_j, -G)), j = 1..4                           calTransportWithInterfaceD2Q5 :976-1013
           reaction A + B -> C on tracers 0, 1, 2: g_ij += J_ij (-/+) k C_0 C_1                     calReactionTracersGPU :95-112
           free-flow outlet: row 0 copies the post-collision populations of row 1                    calFreeConcBoundary3 :461-474
           streaming on the tracer lattice's own table: periodic in x AND y whatever the flow boundaries are, half-way
           bounce back at solids, the rest population stays                                         fillNeighboringNodesTransport :51-75, calStreamingTransportGPU :139-192
           Inamuro inlet: top row, g_4 = C_in - (g_0 + g_1 + g_2 + g_3)                             calInamuroConstConcBoundary :682-698
           C = sum_j g_j
PINNED against tests/golden/tr2d_q5_*.npz (reference kernels, solids + reaction + inlet/outlet rows included).
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


class TracerDenseQ5:
    E = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]])
    OPP = [0, 2, 1, 4, 3]
    W = np.array([1. / 3., 1. / 6., 1. / 6., 1. / 6., 1. / 6.])

    def __init__(self, flow, dxx=(0.05,), dyy=(0.08,), dxy=(0.0,), dyx=(0.0,), beta=(0.6,), criterion=0.5,
                 reaction_rate=None, diff_j=None, inlet_conc=None, freeflow_outlet=False):
        assert flow.L.D == 2, "the reference's 5-velocity tracer lattice is two-dimensional"
        self.flow, self.criterion = flow, criterion
        self.beta = np.asarray(beta, float)
        self.nt = nt = self.beta.size
        M = np.ones((5, 5)); M[1] = [0, 1, -1, 0, 0]; M[2] = [0, 0, 0, 1, -1]; M[3] = [4, -1, -1, -1, -1]; M[4] = [0, 1, 1, -1, -1]
        self.M = M
        self.A = []
        for i in range(nt):
            S = np.eye(5)
            S[1, 1] = 0.5 + 3. * dxx[i]; S[2, 2] = 0.5 + 3. * dyy[i]; S[1, 2] = 3. * dxy[i]; S[2, 1] = 3. * dyx[i]
            self.A.append(-np.dot(np.linalg.inv(M), np.linalg.inv(S)))
        self.rate = reaction_rate
        if reaction_rate is not None:
            assert nt == 3, "the reference's reaction kernel is A + B -> C on three tracers"
            dj = np.asarray(diff_j, float)
            self.J = np.concatenate([dj[:, None], np.repeat((1. - dj)[:, None] / 4., 4, axis=1)], axis=1)
        self.inlet_conc = None if inlet_conc is None else np.asarray(inlet_conc, float)
        self.freeflow = freeflow_outlet

    def set_concentrations(self, conc):
        f = self.flow
        self.conc = np.asarray(conc, float).reshape((self.nt,) + f.shape) * f.dom
        self.g = self.W[None, :, None, None, None] * self.conc[:, None]
        f.stream_only()

    def step(self, n=1):
        f = self.flow; dom = f.dom; ef = self.E.astype(float)
        e3 = [(int(e[0]), int(e[1]), 0) for e in self.E]
        for _ in range(n):
            f.head()
            G = f.gradient()
            with np.errstate(invalid="ignore", divide="ignore"):
                value = np.where(f.rhoR > self.criterion, -0.0, -1.0)
                gn = np.sqrt((G * G).sum(0))
                big = gn > 1.0e-8
                ug = np.where(big, -G / np.where(big, gn, 1.), 0.)
                un = np.sqrt((ug * ug).sum(0))
                conc_lag = self.conc.copy()
                for i in range(self.nt):
                    C = conc_lag[i]
                    geq = np.stack([C * self.W[j] * (1. + 3. * (ef[j, 0] * f.u[0] + ef[j, 1] * f.u[1])) for j in range(5)])
                    diff = np.tensordot(self.M, self.g[i], axes=(1, 0)) - np.tensordot(self.M, geq, axes=(1, 0))
                    g = self.g[i] + np.tensordot(self.A[i], diff, axes=(1, 0))
                    for j in range(1, 5):
                        ok = un > 1.0e-8
                        cos = np.where(ok, (ef[j, 0] * ug[0] + ef[j, 1] * ug[1]) / np.where(ok, un, 1.), 0.)
                        g[j] = g[j] + self.beta[i] * value * (self.W[j] * C) * cos
                    if self.rate is not None:
                        src = self.rate * conc_lag[0] * conc_lag[1] * (1.0 if i == 2 else -1.0)
                        for j in range(5):
                            g[j] = g[j] + self.J[i, j] * src
                    if self.freeflow:
                        ok = dom[0, 0] & dom[0, 1]          # (the reference reads through index -1 where row 1 is solid)
                        g[:, 0, 0] = np.where(ok, g[:, 0, 1], g[:, 0, 0])
                    new = np.empty_like(g)
                    new[0] = g[0]
                    for j in range(1, 5):
                        src_fluid = shift(dom, tuple(-c for c in e3[j]))
                        new[j] = np.where(src_fluid, shift(g[j], tuple(-c for c in e3[j])), g[self.OPP[j]])
                    if self.inlet_conc is not None:
                        new[4, 0, -1] = self.W[4] * ((self.inlet_conc[i] - (new[0, 0, -1] + new[1, 0, -1] + new[2, 0, -1] + new[3, 0, -1])) / self.W[4])
                    self.g[i] = np.where(dom, new, 0.)
                    self.conc[i] = ((((self.g[i, 0] + self.g[i, 1]) + self.g[i, 2]) + self.g[i, 3]) + self.g[i, 4])
            f.body()
