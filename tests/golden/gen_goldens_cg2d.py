#!/usr/bin/env python
"""Golden-vector generator for the D2Q9 colour-gradient (CSF) path.

Runs the UNMODIFIED reference driver `RKColorGradientLBM.runRKColorGradient2DCSF`
(/root/reference/RKCG2D/RKD2Q9.py:1225-1490, commit 3d84189) with its own
kernels from AcceleratedRKGPU2D.py under Numba's CUDA simulator and records
what the reference itself would have written to HDF5 at every output interval
(RKD2Q9.py:1382-1393, 938-957): fluidsRhoR/B, physicalVX/VY, fluidPDFR/B, plus
the integer index structures built by RKD2Q9.py:657-892.

Only I/O is replaced (HDF5 writer, PNG plot, GPU check, `input()`), see
_ref_shim.py.  Two documented deviations, both forced by undefined behaviour
in the reference:
  * `cuda.device_array_like` returns ZEROED memory (the reference reads the
    uninitialised force arrays in its first velocity evaluation,
    RKD2Q9.py:1261-1262 vs :1362-1365);
  * `math.acos` is clamped to [-1, 1] for the simulator (Python raises on a
    rounding overshoot where the GPU returns NaN; both end in "no update",
    AcceleratedRKGPU2D.py:2451-2460).
MRT cases use TauR == TauB: with different taus the reference has a
block-shared-array data race (AcceleratedRKGPU2D.py:1953,1982).

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/gen_goldens_cg2d.py [case ...]
writes tests/golden/cg2d_<case>.npz
"""
import builtins
import contextlib
import io
import math
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: F401  (must precede the reference imports)

import numpy as np
import RKD2Q9
import AcceleratedRKGPU2D as RK
from numba import cuda

REF_COMMIT = "3d84189"


# --- deviations (see module docstring) -------------------------------------
def _zeroed_device_array_like(a):
    return cuda.to_device(np.zeros_like(np.asarray(a.copy_to_host() if hasattr(a, "copy_to_host") else a)))


class _ClampedMath(types.ModuleType):
    def __getattr__(self, k):
        return getattr(math, k)

    @staticmethod
    def acos(x):
        return math.acos(max(-1.0, min(1.0, x)))


RKD2Q9.cuda = types.SimpleNamespace(
    **{k: getattr(cuda, k) for k in ("to_device", "is_available")},
    device_array_like=_zeroed_device_array_like,
)
RK.math = _ClampedMath("math")
builtins.input = lambda *a, **k: ""

INI = """
[ImageSetup]
Existance = 'no'
[DomainSize]
xDomain = {nx}
yDomain = {ny}
numBufferingLayers = 2
ratioTopToBottom = 0.5
[SurfaceTension]
SurfaceTensionType = 'CSF'
SurfaceTensionValue = {sigma}
ContactAngle = {theta}
WettingType = {wetting}
[RKParameters]
AlphaR = 0.44444444
AlphaB = 0.44444444
BetaThickness = {beta}
AkR = 1.4e-1
AkB = 1.4e-1
DeltaValue = {delta}
[FluidParameters]
TauR = {tauR}
TauB = {tauB}
InitialRhoR = {rhoR}
InitialRhoB = {rhoB}
TauType = {tautype}
[BodyForce]
isBodyForce = 'no'
bodyForceX = 0.0
bodyForceY = 0.0
[SolidBoundarySetup]
SolidColorDiff = 0.5
[BoundaryCondition]
BoundaryTypeInlet = '{inlet}'
NeumannType = 'ZouHe'
VelocityYR = {vyr}
VelocityYB = {vyb}
densityBH = {dBH}
densityRH = {dRH}
BoundaryTypeOutlet = '{outlet}'
densityBL = {dBL}
densityRL = {dRL}
[GradientType]
Type = 'Isotropic'
[TimeSetup]
TimeSteps = {steps}
TimeInterval = 1
[Parallelism]
Parallel = 'yes'
xDimension = 128
ThreadsNum = 32
[RelaxationType]
Type = '{relax}'
[CyclesSetup]
IsCycle = 'no'
LastStep = 100
"""

DEFAULTS = dict(sigma=0.1, theta=60.0, wetting=2, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0,
                rhoR=1.0, rhoB=1.0, tautype=2, inlet="Periodic", outlet="Periodic",
                vyr=-1.0e-3, vyb=0.0, dBH=5e-8, dRH=1.00536, dBL=1.0, dRL=5e-8, relax="MRT",
                steps=40)


def geom_open(nx, ny):
    return np.ones((ny, nx), bool)


def geom_block(nx, ny):
    d = np.ones((ny, nx), bool)
    d[ny // 2 - 2:ny // 2 + 2, nx // 2 - 1:nx // 2 + 3] = False
    return d


def geom_walls(nx, ny):
    d = np.ones((ny, nx), bool)
    d[4:-4, 0] = False
    d[4:-4, -1] = False
    d[ny // 2, nx // 2] = False  # one obstacle inside the channel
    return d


def init_droplet(nx, ny, dom, p, cx=None, cy=None, r=4.5):
    cx = nx / 2 if cx is None else cx
    cy = ny / 2 if cy is None else cy
    yy, xx = np.mgrid[0:ny, 0:nx]
    red = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
    return red


def init_top(nx, ny, dom, p):
    yy, xx = np.mgrid[0:ny, 0:nx]
    return yy >= ny - 9


CASES = {
    # name: (nx, ny, geometry, init, overrides)
    "droplet_mrt": (16, 16, geom_open, init_droplet, dict(relax="MRT", tautype=2)),
    "droplet_srt": (16, 16, geom_open, init_droplet, dict(relax="SRT", tautype=1, tauR=1.0, tauB=0.8)),
    "block_w2_mrt": (16, 16, geom_block,
                     lambda nx, ny, d, p: init_droplet(nx, ny, d, p, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                     dict(relax="MRT", wetting=2, theta=60.0)),
    "block_w1_srt": (16, 16, geom_block,
                     lambda nx, ny, d, p: init_droplet(nx, ny, d, p, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                     dict(relax="SRT", wetting=1, theta=120.0, tautype=2, tauR=0.9, tauB=1.1)),
    "channel_neumann_dirichlet_mrt": (12, 26, geom_walls, init_top,
                                      dict(relax="MRT", wetting=2, theta=60.0, inlet="Neumann",
                                           outlet="Dirichlet", steps=40)),
    "channel_dirichlet_convective_srt": (12, 26, geom_walls, init_top,
                                         dict(relax="SRT", wetting=1, theta=45.0, inlet="Dirichlet",
                                              outlet="Convective", dBH=5e-8, dRH=1.002, tautype=1,
                                              tauR=1.0, tauB=0.7, steps=40)),
}


class RefCG(RKD2Q9.RKColorGradientLBM):
    """The reference class with only its I/O replaced."""

    def _RKColorGradientLBM__createHDF5File(self):
        pass

    def _RKColorGradientLBM__checkGPUAvailability(self):
        pass

    def plotDensityDistributionOPT(self, iStep):
        pass

    def resultInHDF5(self, iStep):
        self.snapshots.append(dict(
            rhoR=self.fluidsRhoR.copy(), rhoB=self.fluidsRhoB.copy(),
            ux=self.physicalVX.copy(), uy=self.physicalVY.copy(),
            pdfR=self.fluidPDFR.copy(), pdfB=self.fluidPDFB.copy()))

    def initializeDomainCondition(self):
        # reference allocation (RKD2Q9.py:451-456) + equilibrium fill (:561-585),
        # with the colour layout of the test case instead of the r<=16 droplet
        ny, nx = self.yDomain, self.xDomain
        self.fluidPDFR = np.zeros([ny, nx, 9]); self.fluidPDFB = np.zeros([ny, nx, 9])
        self.fluidsRhoR = np.zeros([ny, nx]); self.fluidsRhoB = np.zeros([ny, nx])
        self.physicalVX = np.zeros([ny, nx]); self.physicalVY = np.zeros([ny, nx])
        red = self._red_mask
        for i in range(ny):
            for j in range(nx):
                if self.isDomain[i, j]:
                    if red[i, j]:
                        self.fluidsRhoR[i, j] = self.initialRhoR; self.fluidsRhoB[i, j] = self._minor
                    else:
                        self.fluidsRhoB[i, j] = self.initialRhoB; self.fluidsRhoR[i, j] = self._minor
                    self.fluidPDFR[i, j], self.fluidPDFB[i, j] = \
                        self._RKColorGradientLBM__initializeFluidPDF(
                            self.fluidsRhoR[i, j], self.fluidsRhoB[i, j], 0.0, 0.0)


def run_case(name):
    nx, ny, geom, init, over = CASES[name]
    p = dict(DEFAULTS); p.update(over); p.update(nx=nx, ny=ny)
    dom = geom(nx, ny)
    RKD2Q9.defineGeometry = lambda x, y: (dom.copy(), ~dom)
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "RKtwophasesetup2D.ini"), "w") as fh:
        fh.write(INI.format(**p))
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        sim = RefCG(tmp)
        sim.snapshots = []
        sim._red_mask = init(nx, ny, dom, p)
        # the minority colour is exactly 0 in the reference init (RKD2Q9.py:474,486);
        # the open-boundary cases use the ini's 5e-8 trace density so that
        # rho_R + rho_B stays non-zero through the Zou-He splits.
        sim._minor = 0.0 if p["inlet"] == "Periodic" else 5e-8
        sim.runRKColorGradient2DCSF()
    snaps = sim.snapshots
    out = dict(
        ref_commit=REF_COMMIT, case=name, nx=nx, ny=ny, is_domain=dom,
        params_keys=np.array(sorted(p.keys())),
        params_vals=np.array([str(p[k]) for k in sorted(p.keys())]),
        red_mask=sim._red_mask, minor=sim._minor,
        fluidNodes=sim.fluidNodes, neighboringNodes=sim.neighboringNodes,
        wettingSolidNodes=sim.wettingSolidNodes,
        neighboringWettingSolidNodes=sim.neighboringWettingSolidNodes,
        rhoR=np.stack([s["rhoR"] for s in snaps]), rhoB=np.stack([s["rhoB"] for s in snaps]),
        ux=np.stack([s["ux"] for s in snaps]), uy=np.stack([s["uy"] for s in snaps]),
        pdfR_first=snaps[0]["pdfR"], pdfB_first=snaps[0]["pdfB"],
        pdfR_last=snaps[-1]["pdfR"], pdfB_last=snaps[-1]["pdfB"],
    )
    if sim.wettingSolidNodes.size > 0:
        out.update(fluidNodesWithSolidGPU=sim.fluidNodesWithSolidGPU,
                   fluidNodesWithSolidOriginal=sim.fluidNodesWithSolidOriginal,
                   nsX=sim.nsX, nsY=sim.nsY)
    path = os.path.join(HERE, "cg2d_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "->", path, "snapshots", len(snaps),
          "mass R %.15g B %.15g" % (out["rhoR"][-1].sum(), out["rhoB"][-1].sum()))


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
