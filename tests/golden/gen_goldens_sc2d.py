#!/usr/bin/env python
"""Golden-vector generator for the D2Q9 Shan-Chen paths (original SC and explicit-forcing SRT/MRT).

Runs the UNMODIFIED reference drivers `ShanChenD2Q9.runOptimizedLBM` (ShanChenD2Q9.py:1433-1629) and
`ShanChenD2Q9.runOptimizedEFLBM` (:1631-2087), reference commit 3d84189, with their own kernels from
OptimizedD2Q9GPU.py / ExplicitD2Q9GPU.py under Numba's CUDA simulator.

The drivers only write results every 80 (SC) / 1000 (EFS) iterations -- far too sparse for the simulator --
so the state is captured by INSTRUMENTATION instead: the module-level name `calPhysicalVelocity` (the last
kernel both loops launch in an iteration) is wrapped so that, after the real kernel has run, the device arrays
it was given (populations, densities, force, velocity) are copied out.  No reference code is changed.
I/O (HDF5, PNG, input()) is stubbed as in gen_goldens_cg2d.py (see _ref_shim.py).  The only deviation:
`cuda.device_array_like` returns zeroed memory (uninitialised in the reference, never read before written
on these paths except ForceX/Y in the EFS pre-loop, which the first force kernel overwrites).

Usage (build container only):  python tests/golden/gen_goldens_sc2d.py [case ...]  -> tests/golden/sc2d_<case>.npz
"""
import builtins
import contextlib
import io
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: F401

import numpy as np
import ShanChenD2Q9 as SC
from numba import cuda

REF_COMMIT = "3d84189"
builtins.input = lambda *a, **k: ""


def _zeroed_device_array_like(a):
    return cuda.to_device(np.zeros_like(np.asarray(a.copy_to_host() if hasattr(a, "copy_to_host") else a)))


SC.cuda = types.SimpleNamespace(**{k: getattr(cuda, k) for k in ("to_device", "is_available")},
                                device_array_like=_zeroed_device_array_like)

BASIC = """
[Scheme]
Type = 'SRT'
[Geometry]
length = 1.0
width = 1.0
nx = {nx}
ny = {ny}
[Time]
TimeLength = 1.0
TimeStep = 1.0
[InitialCondition]
VelocityXLB = 0.0
VelocityYLB = 0.0
[BodyForce]
gValue = 0.0
[FlowDomain]
xDomain = 0,{nx}
yDomain = 0,{ny}
"""
TWOPHASE = """
[PictureSetup]
Exist = 'no'
[SeparationBorder]
xGrid = {nx}
yGrid = {ny}
[FluidsTypes]
NumberOfFluids = 2
[InterType]
InteractionType = '{model}'
[Parallelism]
Parallel = 'yes'
xDimension = 128
ThreadsNum = 32
[RelaxationType]
Type = '{relax}'
[DuplicateDomain]
Option = 'no'
[DICycles]
Option = 'no'
"""
MODELINI = """
[FluidProperties]
InitialDensities = {rho0},{rho1}
BackgroundDensities = {bg0},{bg1}
FluidsTau = {tau0},{tau1}
[{section}]
InteractionFluid = {G}
InteractionSolid = {Gs0},{Gs1}
[ForceScheme]
ExplicitScheme = {scheme}
[BoundaryDefinition]
BoundaryTypeInlet = '{inlet}'
BoundaryMethod = 'ZouHe'
BoundaryTypeOutlet = '{outlet}'
[VelocityBoundary]
velocityX = 0.0,0.0
velocityY = {vy0},{vy1}
[PressureBoundary]
PressureInlet = 1.0,0.06
PressureOutlet = 1.0,0.06
[BodyForce]
Option = 'no'
[Time]
numberTimeStep = {steps}
"""

DEFAULTS = dict(relax="SRT", rho0=1.0, rho1=1.0, bg0=0.06, bg1=0.06, tau0=1.0, tau1=1.0, G=3.8, Gs0=-0.4, Gs1=0.4,
                inlet="Periodic", outlet="Periodic", vy0=0.0, vy1=-1.0e-3, steps=39, scheme=4)


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
    d[ny // 2, nx // 2] = False
    return d


def init_droplet(nx, ny, cx=None, cy=None, r=4.5):
    cx = nx / 2 if cx is None else cx
    cy = ny / 2 if cy is None else cy
    yy, xx = np.mgrid[0:ny, 0:nx]
    return (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r


def init_bottom(nx, ny):
    yy, xx = np.mgrid[0:ny, 0:nx]
    return yy < ny - 9            # fluid 0 below, fluid 1 on top (ShanChenD2Q9.py:759)


CASES = {
    # name: (model, nx, ny, geometry, fluid-0 region, overrides)
    "sc_droplet": ("ShanChen", 16, 16, geom_open, init_droplet, dict()),
    "sc_block": ("ShanChen", 16, 16, geom_block,
                 lambda nx, ny: init_droplet(nx, ny, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6), dict(tau0=1.0, tau1=0.8)),
    "sc_channel_neumann_convective": ("ShanChen", 12, 26, geom_walls, init_bottom,
                                      dict(inlet="Neumann", outlet="Convective", steps=39)),
    "efs_droplet_srt": ("EFS", 16, 16, geom_open, init_droplet, dict(G=0.2, Gs0=-0.14, Gs1=0.14, tau1=0.9)),
    "efs_block_mrt": ("EFS", 16, 16, geom_block,
                      lambda nx, ny: init_droplet(nx, ny, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                      dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, tau0=1.0, tau1=0.8)),
    "efs_channel_neumann_dirichlet_mrt": ("EFS", 12, 26, geom_walls, init_bottom,
                                          dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, inlet="Neumann",
                                               outlet="Dirichlet", vy1=-5.03e-4, steps=39)),
    "efs_channel_neumann_convective_srt": ("EFS", 12, 26, geom_walls, init_bottom,
                                           dict(relax="SRT", G=0.2, Gs0=-0.14, Gs1=0.14, inlet="Neumann",
                                                outlet="Convective", vy1=-5.03e-4, steps=39)),
    # higher-isotropy explicit forcing (24 / 36 neighbours): ExplicitD2Q9GPU.py:392-1372
    "efs_iso8_block_srt": ("EFS", 16, 16, geom_block,
                           lambda nx, ny: init_droplet(nx, ny, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                           dict(relax="SRT", G=0.2, Gs0=-0.14, Gs1=0.14, tau1=0.9, scheme=8)),
    "efs_iso8_droplet_mrt": ("EFS", 16, 16, geom_open, init_droplet, dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, scheme=8)),
    "efs_iso10_block_mrt": ("EFS", 16, 16, geom_block,
                            lambda nx, ny: init_droplet(nx, ny, cx=nx / 2 + 1, cy=ny / 2 + 4, r=3.6),
                            dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, tau1=0.8, scheme=10)),
    "efs_iso8_channel_neumann_dirichlet_srt": ("EFS", 12, 26, geom_walls, init_bottom,
                                               dict(relax="SRT", G=0.2, Gs0=-0.14, Gs1=0.14, inlet="Neumann",
                                                    outlet="Dirichlet", vy1=-5.03e-4, scheme=8)),
}


class RefSC(SC.ShanChenD2Q9):
    def _ShanChenD2Q9__createHDF5File(self):
        pass

    def plotDensityDistributionOPT(self, *a, **k):
        pass

    def plotPhysicalVelocity(self, *a, **k):
        pass

    def resultInHDF5(self, *a, **k):
        pass

    def initializeDomainCondition(self):
        # reference allocation (ShanChenD2Q9.py:740-745) + rest-equilibrium fill (:759-768) with the test layout
        ny, nx = self.ny, self.nx
        self.fluidPDF = np.zeros([self.typesFluids, ny, nx, 9])
        self.fluidsDensity = np.zeros([self.typesFluids, ny, nx])
        self.physicalVX = np.zeros([ny, nx]); self.physicalVY = np.zeros([ny, nx])
        self.forceX = np.zeros([self.typesFluids, ny, nx]); self.forceY = np.zeros([self.typesFluids, ny, nx])
        reg = self._region0
        for i in range(ny):
            for j in range(nx):
                if self.isDomain[i, j]:
                    if reg[i, j]:
                        d0, d1 = self.initialDensities[0], self.backgroundDensities[1]
                    else:
                        d0, d1 = self.backgroundDensities[0], self.initialDensities[1]
                    self.fluidsDensity[0, i, j] = d0; self.fluidPDF[0, i, j, :] = self.weightsCoeff * d0
                    self.fluidsDensity[1, i, j] = d1; self.fluidPDF[1, i, j, :] = self.weightsCoeff * d1


class _Tap:
    """wraps a Numba kernel: K[grid, block](*args) runs the real kernel, then records its device arrays"""

    def __init__(self, real, sink):
        self.real, self.sink = real, sink

    def __getitem__(self, cfg):
        launch = self.real[cfg]

        def run(*args):
            launch(*args)
            # calPhysicalVelocity(totalNodes, numFluids, xDim, fluidPDF, fluidRho, forceX, forceY, physVX, physVY)
            self.sink.append(dict(pdf=args[3].copy_to_host(), rho=args[4].copy_to_host(), fx=args[5].copy_to_host(),
                                  fy=args[6].copy_to_host(), ux=args[7].copy_to_host(), uy=args[8].copy_to_host()))
        return run


def run_case(name):
    model, nx, ny, geom, region, over = CASES[name]
    p = dict(DEFAULTS); p.update(over); p.update(nx=nx, ny=ny, model=model)
    p["section"] = "ShanChenParameters" if model == "ShanChen" else "EFSParameters"
    dom = geom(nx, ny)
    SC.defineGeometry = lambda x, y: (dom.copy(), ~dom)
    tmp = tempfile.mkdtemp()
    for fname, text in (("basicsetup.ini", BASIC), ("twophasesetup.ini", TWOPHASE),
                        ("shanchen2D.ini", MODELINI), ("efs2D.ini", MODELINI)):
        with open(os.path.join(tmp, fname), "w") as fh:
            fh.write(text.format(**p))
    taps = []
    real = SC.calPhysicalVelocity
    SC.calPhysicalVelocity = _Tap(real, taps)
    sink = io.StringIO()
    try:
        with contextlib.redirect_stdout(sink):
            sim = RefSC(tmp)
            sim._region0 = region(nx, ny)
            if model == "ShanChen":
                sim.runOptimizedLBM()
            else:
                sim.runOptimizedEFLBM()
    finally:
        SC.calPhysicalVelocity = real
    if model == "EFS":
        taps = taps[1::2]          # two launches per iteration (:1902 and :2016); the second is the output point
    fn = sim.fluidNodes

    def to2d(a):      # [nf, N] -> [nf, ny, nx];  [N] -> [ny, nx]; [nf, N, 9] -> [nf, ny, nx, 9]
        if a.ndim == 1:
            o = np.zeros(ny * nx); o[fn] = a; return o.reshape(ny, nx)
        if a.ndim == 2:
            o = np.zeros((a.shape[0], ny * nx)); o[:, fn] = a; return o.reshape(a.shape[0], ny, nx)
        o = np.zeros((a.shape[0], ny * nx, 9)); o[:, fn, :] = a; return o.reshape(a.shape[0], ny, nx, 9)

    out = dict(ref_commit=REF_COMMIT, case=name, model=model, nx=nx, ny=ny, is_domain=dom, region0=sim._region0,
               params_keys=np.array(sorted(p.keys())), params_vals=np.array([str(p[k]) for k in sorted(p.keys())]),
               fluidNodes=sim.fluidNodes, neighboringNodes=sim.neighboringNodes,
               rho=np.stack([to2d(t["rho"]) for t in taps]), ux=np.stack([to2d(t["ux"]) for t in taps]),
               uy=np.stack([to2d(t["uy"]) for t in taps]), fx=np.stack([to2d(t["fx"]) for t in taps]),
               fy=np.stack([to2d(t["fy"]) for t in taps]),
               pdf_first=to2d(taps[0]["pdf"]), pdf_last=to2d(taps[-1]["pdf"]))
    path = os.path.join(HERE, "sc2d_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "->", path, "snapshots", len(taps), "mass", out["rho"][-1].sum(axis=(1, 2)), flush=True)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
