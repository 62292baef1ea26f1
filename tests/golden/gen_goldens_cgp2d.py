#!/usr/bin/env python
"""Golden-vector generator for the D2Q9 colour-gradient path with the PERTURBATION surface-tension operator
(SURVEY.md section 8, row f-2).

The reference driver `RKColorGradientLBM.runRKColorGradient2DPerturbation` (/root/reference/RKCG2D/RKD2Q9.py:979-1223,
commit 3d84189) is work in progress and cannot run as shipped: it calls calPhysicalVelocityRKGPU2D with ten arguments
(:1128-1131) where the kernel takes eight (AcceleratedRKGPU2D.py:125), and hands calRecoloringProcess four never-written
arrays (:1218-1223).  This harness therefore executes the reference's own KERNELS verbatim (under Numba's CUDA
simulator), on the reference class's own constants, index arrays and initial state, in the driver's order:
    calStreaming1GPU x2, calStreaming2GPU x2          (:1048-1059)
    calTotalFluidPDF                                  (:1064)
    calMacroDensityRKGPU2D                            (:1121-1124)
    calPhysicalVelocityRKGPU2D                        (:1128, with the kernel's own 8-argument signature)
    [output point: convertOptTo2D + resultInHDF5]     (:1134-1150)
    calPhaseFieldPhi                                  (:1157)
    calRKCollision1GPU2DMRTNew                        (:1191-1199)
    calRKCollision23GPUNew                            (:1211-1219)
The one reading under which the driver is self-consistent is taken:
  * MRT.  The SRT branch collides fR and fB separately (calRKCollision1GPU2DSRTNew) but the recolouring that follows
    rebuilds both from the never-updated total population, i.e. the SRT collision is discarded.  Case `block_srt` wires the
    SRT kernel so that it counts (total population re-formed from its two outputs): the vector for the SRT relaxation the
    reference's 3-D ini asks for.
  * calRecoloringProcess (:1220) is left out: with its never-written inputs zeroed it degenerates to `fR += 0, fB += 0`.
  * closed boxes only: the open-boundary kernels of this driver treat fR / fB AFTER the total population was formed
    (:1063 vs :1065-1118), so their effect never reaches the collision.
  * `[BodyForce] isBodyForce = 'yes'`: the driver reads self.bodyFX / bodyFY unconditionally (:1193).

Usage (build container only):  python tests/golden/gen_goldens_cgp2d.py [case ...]   ->  tests/golden/cgp2d_<case>.npz
"""
import contextlib
import io
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_goldens_cg2d as base          # installs the import shim and the two documented deviations

import numpy as np

RKD2Q9 = base.RKD2Q9

INI = base.INI.replace("SurfaceTensionType = 'CSF'", "SurfaceTensionType = 'Perturbation'") \
              .replace("AkR = 1.4e-1\nAkB = 1.4e-1", "AkR = {akr}\nAkB = {akb}") \
              .replace("isBodyForce = 'no'\nbodyForceX = 0.0\nbodyForceY = 0.0",
                       "isBodyForce = 'yes'\nbodyForceX = {bfx}\nbodyForceY = {bfy}") \
              .replace("SolidColorDiff = 0.5", "SolidColorDiff = {solidphi}")
assert "{akr}" in INI and "{bfx}" in INI and "{solidphi}" in INI

DEFAULTS = dict(base.DEFAULTS, relax="MRT", akr=1.4e-2, akb=1.4e-2, bfx=0.0, bfy=0.0, solidphi=0.5, beta=0.7, steps=40)

CASES = {
    "droplet": (16, 16, base.geom_open, base.init_droplet, dict()),
    "droplet_taus_force": (16, 16, base.geom_open, base.init_droplet, dict(tauR=1.0, tauB=0.8, bfx=2.0e-5, bfy=-1.0e-5,
                                                                            akr=1.0e-2, akb=2.0e-2, beta=0.9)),
    "block_solidphi": (16, 16, base.geom_block,
                       lambda nx, ny, d, p: base.init_droplet(nx, ny, d, p, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                       dict(solidphi=0.4, tauR=0.9, tauB=1.1)),
    # SRT: reference kernel calRKCollision1GPU2DSRTNew wired so that its result is used (the driver drops it, see the docstring);
    # that kernel has no body-force term
    "block_srt": (16, 16, base.geom_block,
                  lambda nx, ny, d, p: base.init_droplet(nx, ny, d, p, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                  dict(relax="SRT", solidphi=0.4, tauR=1.0, tauB=0.8, steps=24)),
}


class RefCGP(base.RefCG):
    def run_perturbation_kernels(self):
        """the kernel sequence of runRKColorGradient2DPerturbation (MRT), see the module docstring"""
        import math
        from numba import cuda
        RK = base.RK
        self.initializeDomainBorder()
        self.initializeDomainCondition()
        self.optimizeFluidArray()
        d = cuda.to_device
        rhoR, rhoB = d(self.optFluidRhoR), d(self.optFluidRhoB)
        fR, fB = d(self.optFluidPDFR), d(self.optFluidPDFB)
        fRn, fBn = d(np.zeros_like(self.optFluidPDFR)), d(np.zeros_like(self.optFluidPDFB))
        vx, vy = d(self.optMacroVelocityX), d(self.optMacroVelocityY)
        phi = d(np.zeros_like(self.optFluidRhoB))
        fT = d(self.optFluidPDFB + self.optFluidPDFR)
        cgx, cgy = d(np.zeros(self.yDomain * self.xDomain)), d(np.zeros(self.yDomain * self.xDomain))
        nodes, nb = d(self.fluidNodes), d(self.neighboringNodes)
        w, cR, cB = d(self.weightsCoeff), d(self.constantCR), d(self.constantCB)
        ex, ey, scheme, bnew = d(self.unitEX), d(self.unitEY), d(self.gradientScheme), d(self.constantBNew)
        if self.relaxationType == "'MRT'":
            M, Mi, S = d(self.transformationM), d(self.invTransformationM), d(self.collisionS)
        n = self.fluidNodes.size
        grid = (int(self.xDimension / self.threadNum), math.ceil(n / self.xDimension)); block = (self.threadNum, 1)
        xd = self.xDimension
        step = record = 0
        while step < self.timeSteps:
            step += 1
            RK.calStreaming1GPU[grid, block](n, xd, nodes, nb, fR, fRn)
            RK.calStreaming1GPU[grid, block](n, xd, nodes, nb, fB, fBn)
            RK.calStreaming2GPU[grid, block](n, xd, fRn, fR)
            RK.calStreaming2GPU[grid, block](n, xd, fBn, fB)
            RK.calTotalFluidPDF[grid, block](n, xd, fR, fB, fT)
            RK.calMacroDensityRKGPU2D[grid, block](n, xd, fR, fB, rhoR, rhoB)
            RK.calPhysicalVelocityRKGPU2D[grid, block](n, xd, fR, fB, rhoR, rhoB, vx, vy)
            if (step - 1) % self.timeInterval == 0:
                self.optFluidRhoR = rhoR.copy_to_host(); self.optFluidRhoB = rhoB.copy_to_host()
                self.optMacroVelocityX = vx.copy_to_host(); self.optMacroVelocityY = vy.copy_to_host()
                self.optFluidPDFB = fB.copy_to_host(); self.optFluidPDFR = fR.copy_to_host()
                self.convertOptTo2D()
                self.resultInHDF5(record)
                record += 1
            RK.calPhaseFieldPhi[grid, block](n, xd, rhoR, rhoB, phi)
            if self.relaxationType == "'SRT'":
                # the reading under which the SRT branch is not discarded: the kernel's two per-colour collisions (:1166-1172)
                # are summed into the total population the recolouring rebuilds both colours from
                RK.calRKCollision1GPU2DSRTNew[grid, block](n, xd, self.deltaValue, self.tauR, self.tauB, ex, ey, cR, cB, w, vx, vy,
                                                           rhoR, rhoB, phi, fR, fB, fRn, fBn)
                RK.calTotalFluidPDF[grid, block](n, xd, fR, fB, fT)
            else:
                RK.calRKCollision1GPU2DMRTNew[grid, block](n, xd, self.deltaValue, self.tauR, self.tauB, self.bodyFX, self.bodyFY,
                                                           ex, ey, cR, cB, w, vx, vy, rhoR, rhoB, phi, fT, M, Mi, S)
            RK.calRKCollision23GPUNew[grid, block](n, xd, self.betaThickness, self.AkR, self.AkB, self.solidPhi, nodes, nb,
                                                   bnew, w, ex, ey, scheme, rhoR, rhoB, phi, cR, cB, fR, fB, cgx, cgy, fT)


def run_case(name):
    nx, ny, geom, init, over = CASES[name]
    p = dict(DEFAULTS); p.update(over); p.update(nx=nx, ny=ny)
    dom = geom(nx, ny)
    RKD2Q9.defineGeometry = lambda x, y: (dom.copy(), ~dom)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "RKtwophasesetup2D.ini"), "w") as fh:
        fh.write(INI.format(**p))
    with contextlib.redirect_stdout(io.StringIO()):
        sim = RefCGP(tmp)
        sim.snapshots = []
        sim._red_mask = init(nx, ny, dom, p)
        sim._minor = 0.0
        sim.run_perturbation_kernels()
    snaps = sim.snapshots
    out = dict(
        ref_commit=base.REF_COMMIT, case=name, nx=nx, ny=ny, is_domain=dom,
        params_keys=np.array(sorted(p.keys())), params_vals=np.array([str(p[k]) for k in sorted(p.keys())]),
        red_mask=sim._red_mask, minor=sim._minor,
        fluidNodes=sim.fluidNodes, neighboringNodes=sim.neighboringNodes,
        rhoR=np.stack([s["rhoR"] for s in snaps]), rhoB=np.stack([s["rhoB"] for s in snaps]),
        ux=np.stack([s["ux"] for s in snaps]), uy=np.stack([s["uy"] for s in snaps]),
        pdfR_first=snaps[0]["pdfR"], pdfB_first=snaps[0]["pdfB"],
        pdfR_last=snaps[-1]["pdfR"], pdfB_last=snaps[-1]["pdfB"])
    path = os.path.join(HERE, "cgp2d_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "->", path, "snapshots", len(snaps), "mass R %.15g B %.15g" % (out["rhoR"][-1].sum(), out["rhoB"][-1].sum()),
          "first mass R %.15g" % out["rhoR"][0].sum())


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
