#!/usr/bin/env python
"""Golden-vector generator for the solute-transport row: D2Q9 tracers riding on the D2Q9
colour-gradient CSF flow.

Neither transport driver of the reference can run as shipped (Transport2DRK.py:26 imports RKGPU2DBoundary, which fails;
runTransport2DMPMCRKNew launches RKGPU2D.calPhysicalVelocityRKGPU2DM, :1293, which exists nowhere).  This harness executes
the reference's own KERNELS verbatim (Numba CUDA simulator) in the order of `runTransport2DMPMCRKNew`
(RKCG2D/Transport2DRK.py:1180-1485), NumberSchemes = 9, on a closed (periodic, all-fluid) box:
    flow   calStreaming1GPU x2, calStreaming2GPU x2, calTotalFluidPDF, calMacroDensityRKGPU2D,          :1180-1296
           calPhysicalVelocityRKGPU2DNew1 (same argument list as the missing ...2DM), calPhaseFieldPhi,
           [flow output], calRKInitialGradient                                                           :1300-1328
    tracer calValueTransportDomain (criterion rho_R > 0.5)                                               :1341
           calCollisionQ9 (SRT) | calCollisionTransportLinearEqlMRTGPUD2Q9 (MRT)                         :1389-1404
           calTransportWithInterfaceD2Q9, Transport2D.calStreaming1GPU / 2GPU, calConcentrationGPU       :1406-1425
           [tracer output]
    flow   calForceTermInColorGradientNew2D, calRKCollision1TotalGPU2DMRTM, calPerturbationFromForce2DMRT,
           calRecoloringProcessM                                                                          :1439-1485
with the constants of Transport2DRK.__init__ (:307-391).  Two repairs, both forced:
  * the MRT tracer kernel is handed the 5-velocity arrays unitVX / unitVY (:1149-1150, 1397) and indexes them up to 8;
    the harness passes the 9-velocity arrays the SRT branch uses (:1390);
  * all-fluid boxes only: the tracer streaming tests `neighbour != -1` on the colour-gradient neighbour table, whose
    wetting solids are <= -2, i.e. next to a solid it writes through a negative index instead of bouncing back.

NumberSchemes = 5 (cases q5_*), the branch :1344-1384 of the same driver, in its order:
    tracer calValueTransportDomain, calCollisionTransportLinearEqlMRTGPU (MRT; the branch has no SRT collision),
           calTransportWithInterfaceD2Q5, [calReactionTracersGPU], [calFreeConcBoundary3], calStreamingTransportGPU /
           calStreamingTransport2GPU on the tracer lattice's own neighbour table (fillNeighboringNodesTransport, solids = -1,
           periodic in x and y whatever the flow boundaries are), [calInamuroConstConcBoundary], calConcentrationGPU
with the constants of Transport2DRK.__init__ (:313-366, 404-410).  Repairs, all forced: the driver never defines
deviceNeighboringNodesTR in this loop (the table of cal1DArrayForFluidDomain, :517-543, is used); betaTracerArray and
concBoundary have one element whatever the number of tracers (:311, :1161; one value per tracer is passed); the reaction
block is mis-indented in the source (:1358-1362).  Here solids are allowed: the tracer table marks them -1 and the flow part
runs the wetting kernels of the CSF loop (calColorValueOnSolid, updateColorGradientOnWettingNew, RKD2Q9.py:1396-1417).

Usage (build container only):  python tests/golden/gen_goldens_tr2d.py [case ...]   ->  tests/golden/tr2d_<case>.npz
"""
import contextlib
import io
import math
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_goldens_cg2d as base

import numpy as np
import AccelerateTransport2DRK as TR

RK = base.RK
RKD2Q9 = base.RKD2Q9

DEFAULTS = dict(base.DEFAULTS, relax="MRT", steps=40, tr_relax="SRT", tr_tau=0.8, dxx=0.05, dyy=0.08, dxy=0.01, dyx=0.02, beta_tr=0.6)

DEFAULTS.update(schemes=9, nt=1, reaction="no", rate=0.0, diffj=1. / 3., tr_inlet="none", tr_outlet="none", conc_in=1.0, geom="open")

CASES = {
    "droplet_srt": (16, 16, dict(tr_relax="SRT", tr_tau=0.8)),
    "droplet_mrt": (16, 16, dict(tr_relax="MRT")),      # flow MRT needs TauR == TauB (block-shared relaxation vector race, see gen_goldens_cg2d.py)
    "q5_droplet": (16, 16, dict(tr_relax="MRT", schemes=5)),
    "q5_block_react_inout": (16, 16, dict(tr_relax="MRT", schemes=5, nt=3, reaction="yes", rate=0.05, diffj=0.3, geom="block",
                                          tr_inlet="Dirichlet", tr_outlet="Freeflow", conc_in=0.7, steps=30)),
}


class RefTR(base.RefCG):
    def initializeDomainCondition(self):
        super().initializeDomainCondition()
        if getattr(self, "_rho_fields", None) is not None:
            # a generic colour field instead of a droplet: next to a solid corner the first arrival of a droplet's signal is an
            # exactly diagonal gradient, anti-parallel to n_s, where the WettingType-2 kernel divides by sin(acos(-1 + 2e-16))
            # (AcceleratedRKGPU2D.py:2455-2463) and the outcome hangs on the last bit
            rR, rB = self._rho_fields
            for i in range(self.yDomain):
                for j in range(self.xDomain):
                    if self.isDomain[i, j]:
                        self.fluidsRhoR[i, j] = rR[i, j]; self.fluidsRhoB[i, j] = rB[i, j]
                        self.fluidPDFR[i, j], self.fluidPDFB[i, j] = \
                            self._RKColorGradientLBM__initializeFluidPDF(rR[i, j], rB[i, j], 0.0, 0.0)

    def run_transport_kernels(self, p):
        from numba import cuda
        self.initializeDomainBorder()
        self.initializeDomainCondition()
        self.optimizeFluidandSolidArray()
        n_cs = n_wf = 0
        if self.wettingSolidNodes.size > 0:             # RKD2Q9.py:1233-1239
            assert p["schemes"] == 5, "the 9-velocity tracer streaming is not usable next to wetting solids"
            self.sortOutFluidNodesToSolid()
            n_wf = self.fluidNodesWithSolidGPU.size
            self.calVectorNormaltoSolid()
            n_cs = self.wettingSolidNodes.size
        d = cuda.to_device
        z = lambda a: d(np.zeros_like(a))
        rhoR, rhoB = d(self.optFluidRhoR), d(self.optFluidRhoB)
        fR, fB = d(self.optFluidPDFR), d(self.optFluidPDFB)
        fRn, fBn = z(self.optFluidPDFR), z(self.optFluidPDFB)
        vx, vy = d(self.optMacroVelocityX), d(self.optMacroVelocityY)
        phi, fT = z(self.optFluidRhoB), z(self.optFluidPDFR)
        Fx, Fy, Gx, Gy, K = (z(self.optFluidRhoB) for _ in range(5))
        solid_color = d(np.zeros(n_cs))
        if n_cs:
            nbS, nodesWS = d(self.neighboringWettingSolidNodes), d(self.fluidNodesWithSolidGPU)
            nsx, nsy = d(self.nsX), d(self.nsY)
        nodes, nb = d(self.fluidNodes), d(self.neighboringNodes)
        w, ex, ey = d(self.weightsCoeff), d(self.unitEX), d(self.unitEY)
        M, Mi, S = d(self.transformationM), d(self.invTransformationM), d(self.collisionS)
        n = self.fluidNodes.size
        grid = (int(self.xDimension / self.threadNum), math.ceil(n / self.xDimension)); block = (self.threadNum, 1)
        xd = self.xDimension
        # ---- transport constants, Transport2DRK.py:307-391 (NumberSchemes = 9, one tracer) ----
        nt = p["nt"]
        Q = p["schemes"]
        if Q == 5:
            # Transport2DRK.py:313-347, 404-410, 517-543
            wTR = np.array([1. / 3., 1. / 6., 1. / 6., 1. / 6., 1. / 6.])
            tM = np.ones((5, 5)); tM[1, 0] = 0; tM[1, 2] = -1.; tM[1, 3:] = 0.; tM[2, :3] = 0.; tM[2, 4] = -1.
            tM[3, 0] = 4.; tM[3, 1:] = -1.; tM[4, 0] = 0.; tM[4, 3:] = -1.
            relS5 = np.zeros((nt, 5, 5))
            for i in range(nt):
                relS5[i, 1, 1] = 0.5 + 3. * p["dxx"]; relS5[i, 2, 2] = 0.5 + 3. * p["dyy"]
                relS5[i, 1, 2] = 3. * p["dxy"]; relS5[i, 2, 1] = 3. * p["dyx"]
                relS5[i, 0, 0] = relS5[i, 3, 3] = relS5[i, 4, 4] = 1.0
            inv_rel5 = np.stack([-np.dot(np.linalg.inv(tM), np.linalg.inv(relS5[i])) for i in range(nt)])
            d_tM, d_inv_rel5 = d(tM), d(inv_rel5)
            unitVX, unitVY = d(np.array([0., 1., -1., 0., 0.])), d(np.array([0., 0., 0., 1., -1.]))
            diffJED = np.zeros((nt, 5)); diffJED[:, 0] = p["diffj"]; diffJED[:, 1:] = (1. - p["diffj"]) / 4.
            d_diffJED, d_rate = d(diffJED), d(np.array([p["rate"]]))
            idx2d = -np.ones((self.yDomain, self.xDomain), np.int64)
            idx2d.reshape(-1)[self.fluidNodes] = np.arange(n)
            nbTR = d(np.zeros(4 * n, np.int64))
            TR.fillNeighboringNodesTransport[grid, block](n, self.xDomain, self.yDomain, xd, nodes, d(idx2d), nbTR)
            conc_b = d(np.full(nt, p["conc_in"]))
        else:
            wTR = np.zeros(9); wTR[0] = 4. / 9.; wTR[1:5] = 1. / 9.; wTR[5:] = 1. / 36.
        relS = np.zeros((nt, 9, 9))
        for i in range(nt):
            for k in (0, 1, 2, 7, 8):
                relS[i, k, k] = 1.
            relS[i, 3, 3] = relS[i, 4, 4] = 0.5 + 3. * p["dxx"]
            relS[i, 5, 5] = relS[i, 6, 6] = 0.5 + 3. * p["dyy"]
            relS[i, 3, 5] = 3. * p["dxy"]; relS[i, 5, 3] = 3. * p["dyx"]
        inv_rel = np.stack([-np.dot(self.invTransformationM, np.linalg.inv(relS[i])) for i in range(nt)])
        tau_tr = d(np.full(nt, p["tr_tau"])); beta_tr = d(np.full(nt, p["beta_tr"]))
        d_wTR, d_inv_rel = d(wTR), d(inv_rel)
        conc0 = np.ascontiguousarray(self._tracer0.reshape(nt, -1)[:, self.fluidNodes])
        conc = d(conc0.copy()); pdf = d(np.ascontiguousarray(conc0[:, :, None] * wTR)); pdfn = d(np.ascontiguousarray(conc0[:, :, None] * wTR))
        value_td = d(np.ones(n))
        self.tracer_snapshots = []
        step = record = 0
        while step < self.timeSteps:
            step += 1
            RK.calStreaming1GPU[grid, block](n, xd, nodes, nb, fR, fRn)
            RK.calStreaming1GPU[grid, block](n, xd, nodes, nb, fB, fBn)
            RK.calStreaming2GPU[grid, block](n, xd, fRn, fR)
            RK.calStreaming2GPU[grid, block](n, xd, fBn, fB)
            RK.calTotalFluidPDF[grid, block](n, xd, fR, fB, fT)
            RK.calMacroDensityRKGPU2D[grid, block](n, xd, fR, fB, rhoR, rhoB)
            RK.calPhysicalVelocityRKGPU2DNew1[grid, block](n, xd, fT, rhoR, rhoB, vx, vy, Fx, Fy)
            RK.calPhaseFieldPhi[grid, block](n, xd, rhoR, rhoB, phi)
            if (step - 1) % self.timeInterval == 0:
                self.optFluidRhoR = rhoR.copy_to_host(); self.optFluidRhoB = rhoB.copy_to_host()
                self.optMacroVelocityX = vx.copy_to_host(); self.optMacroVelocityY = vy.copy_to_host()
                self.optFluidPDFR = fR.copy_to_host(); self.optFluidPDFB = fB.copy_to_host()
                self.convertOptTo2D()
                self.resultInHDF5(record)
            if n_cs:
                RK.calColorValueOnSolid[grid, block](n_cs, xd, nbS, w, phi, solid_color)
            RK.calRKInitialGradient[grid, block](n, xd, n_cs, nodes, nb, w, ex, ey, phi, solid_color, Gx, Gy)
            if n_cs:
                RK.updateColorGradientOnWettingNew[grid, block](n_wf, xd, self.cosTheta, self.sinTheta, nodesWS, nsx, nsy, Gx, Gy)
            TR.calValueTransportDomain[grid, block](n, xd, 0.5, value_td, rhoR)
            if Q == 5:
                TR.calCollisionTransportLinearEqlMRTGPU[grid, block](n, xd, nt, unitVX, unitVY, vx, vy, conc, pdf, d_tM, d_inv_rel5, d_wTR)
                TR.calTransportWithInterfaceD2Q5[grid, block](n, xd, nt, beta_tr, value_td, ex, ey, Gx, Gy, d_wTR, conc, pdf)
                if p["reaction"] == "yes":
                    TR.calReactionTracersGPU[grid, block](n, nt, xd, d_rate, d_diffJED, conc, pdf)
                if p["tr_outlet"] == "Freeflow":
                    TR.calFreeConcBoundary3[grid, block](n, nt, self.xDomain, xd, nodes, nbTR, conc, pdf)
                TR.calStreamingTransportGPU[grid, block](n, xd, nt, nbTR, pdf, pdfn)
                TR.calStreamingTransport2GPU[grid, block](n, nt, xd, pdfn, pdf)
                if p["tr_inlet"] == "Dirichlet":
                    TR.calInamuroConstConcBoundary[grid, block](n, xd, nt, self.yDomain, self.xDomain, nodes, nbTR, conc_b, d_wTR, pdf)
                TR.calConcentrationGPU[grid, block](n, nt, xd, 5, conc, pdf)
            elif p["tr_relax"] == "SRT":
                TR.calCollisionQ9[grid, block](n, xd, nt, ex, ey, vx, vy, tau_tr, conc, pdf, d_wTR)
            else:
                TR.calCollisionTransportLinearEqlMRTGPUD2Q9[grid, block](n, xd, nt, ex, ey, vx, vy, conc, pdf, M, d_inv_rel, d_wTR)
            if Q == 9:
                TR.calTransportWithInterfaceD2Q9[grid, block](n, xd, nt, beta_tr, value_td, ex, ey, Gx, Gy, d_wTR, conc, pdf)
                TR.calStreaming1GPU[grid, block](n, nt, xd, nodes, nb, pdf, pdfn)
                TR.calStreaming2GPU[grid, block](n, nt, xd, pdfn, pdf)
                TR.calConcentrationGPU[grid, block](n, nt, xd, 9, conc, pdf)
            if (step - 1) % self.timeInterval == 0:
                c2 = np.zeros((nt, self.yDomain, self.xDomain))
                c2.reshape(nt, -1)[:, self.fluidNodes] = conc.copy_to_host()
                self.tracer_snapshots.append(c2)
                record += 1
            RK.calForceTermInColorGradientNew2D[grid, block](n, xd, self.surfaceTension, nb, w, ex, ey, Gx, Gy, Fx, Fy, K)
            RK.calRKCollision1TotalGPU2DMRTM[grid, block](n, xd, self.tauCalculation, self.tauR, self.tauB, self.deltaValue, ex, ey, w,
                                                          vx, vy, rhoR, rhoB, phi, fT, M, Mi, S)
            RK.calPerturbationFromForce2DMRT[grid, block](n, xd, self.tauCalculation, self.tauR, self.tauB, self.deltaValue, w, ex, ey,
                                                          vx, vy, Fx, Fy, phi, fT, M, Mi, S, rhoR, rhoB)
            RK.calRecoloringProcessM[grid, block](n, xd, self.betaThickness, w, rhoR, rhoB, ex, ey, Gx, Gy, fR, fB, fT)


def run_case(name):
    nx, ny, over = CASES[name]
    p = dict(DEFAULTS); p.update(over); p.update(nx=nx, ny=ny)
    dom = base.geom_block(nx, ny) if p["geom"] == "block" else base.geom_open(nx, ny)
    RKD2Q9.defineGeometry = lambda x, y: (dom.copy(), ~dom)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "RKtwophasesetup2D.ini"), "w") as fh:
        fh.write(base.INI.format(**{k: v for k, v in p.items()}))
    with contextlib.redirect_stdout(io.StringIO()):
        sim = RefTR(tmp)
        sim.snapshots = []
        sim._red_mask = base.init_droplet(nx, ny, dom, p, cx=7.3, cy=8.6)      # off-centre: no symmetric cancellation nodes
        sim._minor = 0.0
        sim._rho_fields = None
        if p["geom"] == "block":
            r = 0.5 + 0.4 * (np.random.default_rng(7).random((ny, nx)) - 0.5)
            sim._rho_fields = (r * dom, (1.0 - r) * dom)
        yy, xx = np.mgrid[0:ny, 0:nx]
        sim._tracer0 = np.stack([np.where(xx >= nx // 2, 1.0, 0.2) if k != 1 else np.where(yy >= ny // 2, 0.9, 0.1) for k in range(p["nt"])])
        if p["nt"] == 3:
            sim._tracer0[2] *= 0.25
        sim._tracer0 = sim._tracer0 * dom
        sim.run_transport_kernels(p)
    snaps = sim.snapshots
    out = dict(ref_commit=base.REF_COMMIT, case=name, nx=nx, ny=ny, is_domain=dom,
               params_keys=np.array(sorted(p.keys())), params_vals=np.array([str(p[k]) for k in sorted(p.keys())]),
               red_mask=sim._red_mask, minor=sim._minor, tracer0=sim._tracer0,
               **({} if sim._rho_fields is None else dict(rhoR0=sim._rho_fields[0], rhoB0=sim._rho_fields[1])),
               rhoR=np.stack([s["rhoR"] for s in snaps]), rhoB=np.stack([s["rhoB"] for s in snaps]),
               ux=np.stack([s["ux"] for s in snaps]), uy=np.stack([s["uy"] for s in snaps]),
               conc=np.stack(sim.tracer_snapshots))
    path = os.path.join(HERE, "tr2d_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "->", path, "snapshots", len(snaps), "tracer mass first %.15g last %.15g" % (out["conc"][0].sum(), out["conc"][-1].sum()))


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
