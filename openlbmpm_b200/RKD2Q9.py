"""`RKColorGradientLBM` -- drop-in for the reference's colour-gradient host class (RKCG2D/RKD2Q9.py:23-1498).

Same constructor (`RKColorGradientLBM(pathIniFile)`, reads `RKtwophasesetup2D.ini`), same public attributes
(`isDomain`, `isSolid`, `fluidsRhoR/B`, `fluidPDFR/B`, `physicalVX/VY`, `fluidNodes`, `neighboringNodes`,
`wettingSolidNodes`, `fluidNodesWithSolidGPU`, `nsX/nsY`, ...) and the same entry points
(`runRKColorGradient2D`, `runRKColorGradient2DCSF`; plus `runModifiedRKColorGradient2D`, which the reference's
main.py:53 calls although the class never defined it).  The per-step Python loop that launched 17-21
Numba-CUDA kernels (RKD2Q9.py:1295-1490) is replaced by one `lbm_step(TimeInterval)` call into
liblbmpm.so; geometry compaction, neighbour tables and solid normals (RKD2Q9.py:657-892, Python loops) are
built on the GPU and exported bit-exactly.  There is no CPU fallback.
"""
import os
import sys
import time

import numpy as np

from . import _lib
from .inifile import Ini, IniError
from .results import ResultFile

INLET = {"periodic": _lib.BC_PERIODIC, "neumann": _lib.INLET_VELOCITY, "dirichlet": _lib.INLET_PRESSURE}
OUTLET = {"periodic": _lib.BC_PERIODIC, "convective": _lib.OUTLET_CONVECTIVE, "dirichlet": _lib.OUTLET_PRESSURE}


class RKColorGradientLBM:
    INI_NAME = "RKtwophasesetup2D.ini"
    LATTICE = 9

    def __init__(self, pathIniFile, verbose=True):
        self.pathIni = pathIniFile
        self.verbose = verbose
        ini = Ini(pathIniFile, self.INI_NAME)
        self.imageExist = ini.quoted("ImageSetup", "Existance", default="no")
        self._read_domain(ini)
        # surface tension / recolouring (RKD2Q9.py:60-140)
        self.surfaceTensionType = ini.quoted("SurfaceTension", "SurfaceTensionType", default=self._default_surface_tension(ini))
        if self.surfaceTensionType not in ("'CSF'", "'Perturbation'"):
            raise IniError("SurfaceTensionType must be 'CSF' or 'Perturbation'")
        self.surfaceTension = ini.number("SurfaceTension", "SurfaceTensionValue", "SurfaceTension", default=0.1)
        self.contactAngle = ini.number("SurfaceTension", "ContactAngle", default=90.0)
        self.cosTheta = np.cos(self.contactAngle / 180. * np.pi)
        self.sinTheta = np.sin(self.contactAngle / 180. * np.pi)
        self.wettingType = ini.integer("SurfaceTension", "WettingType", default=2)
        self.betaThickness = ini.number("RKParameters", "BetaThickness", default=0.7)
        self.deltaValue = ini.number("RKParameters", "DeltaValue", default=0.98)
        self.alphaR = ini.number("RKParameters", "AlphaR", default=4. / 9.)
        self.alphaB = ini.number("RKParameters", "AlphaB", default=4. / 9.)
        # perturbation operator (RKD2Q9.py:107-133, 182): surface strengths and the colour seen on solid neighbours
        self.AkR = ini.number("RKParameters", "AkR", default=0.0)
        self.AkB = ini.number("RKParameters", "AkB", default=0.0)
        self.solidPhi = ini.number("SolidBoundarySetup", "SolidColorDiff", default=self._default_solid_phi(ini))
        self.tauR = ini.number("FluidParameters", "TauR")
        self.tauB = ini.number("FluidParameters", "TauB")
        self.initialRhoR = ini.number("FluidParameters", "InitialRhoR")
        self.initialRhoB = ini.number("FluidParameters", "InitialRhoB")
        self.tauCalculation = ini.integer("FluidParameters", "TauType", default=2)
        self.isBodyForce = ini.quoted("BodyForce", "isBodyForce", default="no")
        self.bodyFX = ini.number("BodyForce", "bodyForceX", default=0.0)
        self.bodyFY = ini.number("BodyForce", "bodyForceY", default=0.0)
        self.bodyFZ = ini.number("BodyForce", "bodyForceZ", default=0.0)
        if self.isBodyForce == "'yes'" and self.surfaceTensionType == "'CSF'":
            raise IniError("A body force is read but never applied by the reference's CSF loop (RKD2Q9.py:1225-1490).")
        self._read_time(ini)
        self.Parallel = ini.quoted("Parallelism", "Parallel", default="yes")
        self.xDimension = ini.integer("Parallelism", "xDimension", default=128)    # kept for compatibility, unused
        self.threadNum = ini.integer("Parallelism", "ThreadsNum", default=32)
        self.numGPUs = ini.integer("Parallelism", "NumGPUs", default=1)
        self.relaxationType = ini.quoted("RelaxationType", "Type", default="MRT")
        self._read_boundaries(ini)
        self.isCycles = ini.quoted("CyclesSetup", "IsCycle", default="no")
        if self.isCycles == "'yes'":
            self.lastStep = ini.integer("CyclesSetup", "LastStep")
        # lattice constants (RKD2Q9.py:299-303)
        self.weightsCoeff = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
        self.unitEX = np.array([0., 1., 0., -1., 0., 1., -1., -1., 1.])
        self.unitEY = np.array([0., 0., 1., 0., -1., 1., 1., -1., -1.])
        self.engine = None
        self._results = None
        # slab decomposition along the flow axis: one process per GPU under `torchrun` (slab.from_environment), every
        # rank holds the whole host-side arrays, runs its own slab on its GPU and gathers the results
        self.slabs = None

    def _default_surface_tension(self, ini):
        return "CSF"

    def _default_solid_phi(self, ini):
        return 0.0

    # -- ini pieces (overridden by the 3-D class) ---------------------------------------------------
    def _read_domain(self, ini):
        self.xDomain = ini.integer("DomainSize", "xDomain")
        self.yDomain = ini.integer("DomainSize", "yDomain")
        self.numBufferingLayers = ini.integer("DomainSize", "numBufferingLayers", default=0)
        self.ratioTopToBottom = ini.number("DomainSize", "ratioTopToBottom", default=0.5)

    def _read_time(self, ini):
        self.timeSteps = ini.integer("TimeSetup", "TimeSteps")
        self.timeInterval = ini.integer("TimeSetup", "TimeInterval", default=max(1, self.timeSteps))

    def _read_boundaries(self, ini):
        self.boundaryTypeInlet = ini.quoted("BoundaryCondition", "BoundaryTypeInlet", default="Periodic")
        self.boundaryTypeOutlet = ini.quoted("BoundaryCondition", "BoundaryTypeOutlet", default="Periodic")
        self.velocityYR = self.velocityYB = 0.0
        self.densityRhoBH = self.densityRhoRH = self.densityRhoBL = self.densityRhoRL = 0.0
        if self.boundaryTypeInlet == "'Neumann'":
            self.neumannType = ini.quoted("BoundaryCondition", "NeumannType", default="ZouHe")
            self.velocityYR = ini.number("BoundaryCondition", "VelocityYR", default=0.0)
            self.velocityYB = ini.number("BoundaryCondition", "VelocityYB", default=0.0)
        elif self.boundaryTypeInlet == "'Dirichlet'":
            self.densityRhoBH = ini.number("BoundaryCondition", "densityBH")
            self.densityRhoRH = ini.number("BoundaryCondition", "densityRH")
        if self.boundaryTypeOutlet == "'Dirichlet'":
            self.densityRhoBL = ini.number("BoundaryCondition", "densityBL")
            self.densityRhoRL = ini.number("BoundaryCondition", "densityRL")

    def _say(self, *a):
        if self.verbose and (self.slabs is None or self.slabs.rank == 0):
            print(*a)

    # -- geometry and initial condition ---------------------------------------------------------------
    def _shape(self):
        return (self.yDomain, self.xDomain)

    def initializeDomainBorder(self):
        """RKD2Q9.py:417-443: `defineGeometry(xDomain, yDomain)` or ~/StructureImage/structure.png (0 = solid)"""
        if self.imageExist == "'yes'":
            self._process_image()
        else:
            try:
                from SimpleGeometryRK import defineGeometry       # a user's own module on sys.path wins
            except ImportError:
                from .SimpleGeometryRK import defineGeometry
            self.isDomain, self.isSolid = defineGeometry(self.xDomain, self.yDomain)
        self.isDomain = np.ascontiguousarray(self.isDomain, dtype=bool)
        self.isSolid = ~self.isDomain
        self.voidSpace = int(np.count_nonzero(self.isDomain))
        self._say('The number of vexls in void space is %g.' % self.voidSpace)
        self._say('The porosity of the layout is %f.' % (self.voidSpace / self.isDomain.size))

    def _process_image(self):
        """RKD2Q9.py:373-414: crop to the solid pixels (== 0), solid side columns, `int(2 n ratioTopToBottom)` void
        buffer rows in front of the image (low row indices, the outlet side) and the rest of the `2 n` behind it"""
        from . import imagegeo
        try:
            img = imagegeo.crop_to_solid(imagegeo.load_gray())
        except (FileNotFoundError, ValueError) as e:
            raise IniError(str(e))
        low = int(2 * self.numBufferingLayers * self.ratioTopToBottom)
        self.effectiveDomain = imagegeo.close_and_pad(img, low, 2 * self.numBufferingLayers - low)
        self.isDomain = imagegeo.to_domain(self.effectiveDomain)
        self.yDomain, self.xDomain = self.isDomain.shape
        self.originalXdim = self.xDomain
        self._say('Now the size of domain is %g and %g' % (self.yDomain, self.xDomain))

    def _equilibrium(self, rho, vel):
        """RKD2Q9.py:577-601: f_i = rho w_i (1 + 3 e.u + 4.5 (e.u)^2 - 1.5 u.u), dense `[..., Q]`"""
        e = [self.unitEX, self.unitEY] + ([self.unitEZ] if hasattr(self, "unitEZ") else [])
        eu = sum(e[a] * vel[a][..., None] for a in range(len(vel)))
        uu = sum(v * v for v in vel)[..., None]
        return rho[..., None] * self._weights() * (1. + (3. * eu + 4.5 * eu * eu - 1.5 * uu))

    def initializeDomainCondition(self):
        """RKD2Q9.py:445-559.  No image: droplet of R with radius 16 around the centre, B elsewhere, at rest (assign
        `self.initialRedRegion`, a boolean array, beforehand for another layout); image: R below the inlet buffer
        rows, B in them; `[CyclesSetup] IsCycle = 'yes'`: drainage-imbibition restart from the previous run's results
        in ~/LBMInitial (LBM_INITIAL_DIR)."""
        shape = self._shape()
        self.physicalVX = np.zeros(shape); self.physicalVY = np.zeros(shape)
        if len(shape) == 3:
            self.physicalVZ = np.zeros(shape)
        if self.isCycles == "'yes'":
            return self._initialize_from_previous_run()
        red = getattr(self, "initialRedRegion", None)
        if red is None:
            idx = np.indices(shape)
            if self.imageExist == "'yes'":
                red = idx[0] < shape[0] - self.numBufferingLayers
            else:
                centre = [int(n / 2) for n in shape]
                red = np.sqrt(sum((idx[a] - centre[a]) ** 2 for a in range(len(shape)))) <= 16.
        red = np.asarray(red, bool) & self.isDomain
        self.fluidsRhoR = np.where(red, self.initialRhoR, 0.0) * self.isDomain
        self.fluidsRhoB = np.where(red, 0.0, self.initialRhoB) * self.isDomain
        self.fluidPDFR = self.fluidsRhoR[..., None] * self._weights()
        self.fluidPDFB = self.fluidsRhoB[..., None] * self._weights()

    def _initialize_from_previous_run(self):
        """RKD2Q9.py:491-559"""
        from .results import initial_dir, read_arrays
        shape = self._shape()
        if len(shape) != 2:
            raise IniError("[CyclesSetup] IsCycle = 'yes' is a 2-D feature of the reference")
        try:
            if self.imageExist == "'no'":
                # :491-508 -- densities and velocity of record LastStep, fresh B in the top 20 rows, equilibrium populations
                n = self.lastStep
                names = ["/FluidMacro/FluidDensityRin%d" % n, "/FluidMacro/FluidDensityBin%d" % n,
                         "/FluidVelocity/FluidVelocityXAt%d" % n, "/FluidVelocity/FluidVelocityYAt%d" % n]
                d = read_arrays(initial_dir(), "SimulationResultsRK.h5", names)
                self.fluidsRhoR, self.fluidsRhoB, self.physicalVX, self.physicalVY = (np.array(d[k], float) for k in names)
                self.fluidsRhoR[-20:, :] = 0.; self.fluidsRhoB[-20:, :] = self.initialRhoB
                self.fluidsRhoR *= self.isDomain; self.fluidsRhoB *= self.isDomain
                vel = [self.physicalVX, self.physicalVY]
                self.fluidPDFR = self._equilibrium(self.fluidsRhoR, vel) * self.isDomain[..., None]
                self.fluidPDFB = self._equilibrium(self.fluidsRhoB, vel) * self.isDomain[..., None]
            else:
                # :533-559 -- cycleInitialRK.h5: state of the previous half cycle, the colours swapped in the inlet buffer rows
                names = ["/FluidMacro/FluidDensityR", "/FluidMacro/FluidDensityB", "/FluidPDF/FluidPDFR", "/FluidPDF/FluidPDFB",
                         "/FluidVelocity/FluidVelocityX", "/FluidVelocity/FluidVelocityY"]
                d = read_arrays(initial_dir(), "cycleInitialRK.h5", names)
                rR, rB, fR, fB, vx, vy = (np.array(d[k], float) for k in names)
                nb = self.numBufferingLayers
                self.fluidsRhoR, self.fluidsRhoB, self.fluidPDFR, self.fluidPDFB = rR.copy(), rB.copy(), fR.copy(), fB.copy()
                if nb > 0:
                    self.fluidsRhoR[-nb:], self.fluidsRhoB[-nb:] = rB[-nb:], rR[-nb:]
                    self.fluidPDFR[-nb:], self.fluidPDFB[-nb:] = fB[-nb:], fR[-nb:]
                self.physicalVX, self.physicalVY = vx, vy
        except FileNotFoundError as e:
            raise IniError("There is no file for initializing the domain: %s" % e)
        if self.fluidsRhoR.shape != shape:
            raise IniError("the restart file holds arrays of shape %s, the domain is %s" % (self.fluidsRhoR.shape, shape))

    def _weights(self):
        return self.weightsCoeff

    # -- engine ---------------------------------------------------------------------------------------
    def _make_engine(self):
        inlet = self.boundaryTypeInlet.strip("'").lower(); outlet = self.boundaryTypeOutlet.strip("'").lower()
        if inlet not in INLET or outlet not in OUTLET:
            raise IniError("Unknown boundary type %s / %s" % (self.boundaryTypeInlet, self.boundaryTypeOutlet))
        relax = _lib.RELAX_MRT if self.relaxationType == "'MRT'" else _lib.RELAX_SRT
        if self.slabs is None:
            from . import slab
            self.slabs = slab.from_environment()
        shape = self._shape()
        if self.slabs is not None:
            from . import slab
            lo, hi = slab.slab_bounds(shape[0], self.slabs.rank, self.slabs.world)
            self._slab = slice(lo, hi)
            shape = (hi - lo,) + tuple(shape[1:])
        else:
            self._slab = slice(None)
        self.engine = _lib.Engine(self.LATTICE, shape, model=_lib.MODEL_CG, relax=relax,
                                  device=self.slabs.device_index if self.slabs is not None else 0,
                                  sigma=self.surfaceTension, contact_angle_deg=self.contactAngle,
                                  wetting_type=self.wettingType, beta=self.betaThickness, delta=self.deltaValue,
                                  tauR=self.tauR, tauB=self.tauB, tau_type=self.tauCalculation,
                                  inlet=INLET[inlet], outlet=OUTLET[outlet], inlet_velocity=self._inlet_velocity(),
                                  rhoBH=self.densityRhoBH, rhoRH=self.densityRhoRH, rhoBL=self.densityRhoBL,
                                  rhoRL=self.densityRhoRL,
                                  surface_tension_type=_lib.ST_PERTURBATION if self.surfaceTensionType == "'Perturbation'" else _lib.ST_CSF,
                                  AkR=self.AkR, AkB=self.AkB, solid_phi=self.solidPhi,
                                  body_force=[self.bodyFX, self.bodyFY, self.bodyFZ if self.LATTICE == 19 else 0.0]
                                  if self.isBodyForce == "'yes'" else [0.0, 0.0, 0.0])
        if self.slabs is not None:
            self.engine.comm_init(self.slabs.rank, self.slabs.world, self.slabs.unique_id(self.engine))
        self.engine.set_geometry(self.isDomain[self._slab])

    def _inlet_velocity(self):
        return self.velocityYB + self.velocityYR            # RKD2Q9.py:1300

    def optimizeFluidandSolidArray(self):
        """RKD2Q9.py:657-736 (+ sortOutFluidNodesToSolid 741-760, calVectorNormaltoSolid 768-892): the
        compact index structures, built on the device and exported bit-exactly."""
        if self.engine is None:
            self._make_engine()
        if self.slabs is not None:
            # the compact node numbering of the reference is a property of the whole lattice; a slab only knows its own
            self._say("slab decomposition: the reference's compact index arrays are not exported")
            return
        idx = self.engine.export_indexing()
        self.fluidNodes = idx["fluidNodes"]
        self.neighboringNodes = idx["neighboringNodes"]
        self.wettingSolidNodes = idx["wettingSolidNodes"]
        self.neighboringWettingSolidNodes = idx["neighboringWettingSolidNodes"]
        self.fluidNodesWithSolidGPU = idx["fluidNodesWithSolidGPU"]
        self.fluidNodesWithSolidOriginal = idx["fluidNodesWithSolidOriginal"]
        self.nsX, self.nsY = idx["nsX"], idx["nsY"]
        if "nsZ" in idx:
            self.nsZ = idx["nsZ"]
        flat = self.fluidNodes
        Q = self.LATTICE
        self.optFluidPDFR = self.fluidPDFR.reshape(-1, Q)[flat]
        self.optFluidPDFB = self.fluidPDFB.reshape(-1, Q)[flat]
        self.optFluidRhoR = self.fluidsRhoR.ravel()[flat]
        self.optFluidRhoB = self.fluidsRhoB.ravel()[flat]

    sortOutFluidNodesToSolid = calVectorNormaltoSolid = lambda self: None    # folded into optimizeFluidandSolidArray

    def convertOptTo2D(self):
        """RKD2Q9.py:902-911: refresh the dense host arrays from the device state (output point of the loop)"""
        rho, u = self.engine.download_macros()
        pdf = self.engine.download_pdfs()
        if self.slabs is not None:          # every rank receives the whole lattice
            rho = [self.slabs.gather(a) for a in rho]; u = [self.slabs.gather(a) for a in u]
            pdf = [self.slabs.gather(a) for a in pdf]
        self.fluidsRhoR, self.fluidsRhoB = rho
        self.physicalVX, self.physicalVY = u[0], u[1]
        if len(u) == 3:
            self.physicalVZ = u[2]
        self.fluidPDFR, self.fluidPDFB = pdf

    def resultInHDF5(self, iStep):
        """RKD2Q9.py:938-957, same group / dataset names"""
        if self.slabs is not None and self.slabs.rank != 0:
            return                                  # rank 0 writes the gathered arrays
        if self._results is None:
            self._results = ResultFile("SimulationResultsRK.h5")
        arrays = {"/FluidMacro/FluidDensityRin%g" % iStep: self.fluidsRhoR,
                  "/FluidMacro/FluidDensityBin%g" % iStep: self.fluidsRhoB,
                  "/FluidPDF/FluidPDFBat%g" % iStep: self.fluidPDFB,
                  "/FluidPDF/FluidPDFRat%g" % iStep: self.fluidPDFR,
                  "/FluidVelocity/FluidVelocityXAt%g" % iStep: self.physicalVX,
                  "/FluidVelocity/FluidVelocityYAt%g" % iStep: self.physicalVY}
        if hasattr(self, "physicalVZ"):
            arrays["/FluidVelocity/FluidVelocityZAt%g" % iStep] = self.physicalVZ
        self._results.write(iStep, arrays)

    def _write_macro_record(self, iStep, rho, u):
        """record of the asynchronous output: the reference's dataset names, densities and velocities only"""
        if self._results is None:
            self._results = ResultFile("SimulationResultsRK.h5")
        arrays = {"/FluidMacro/FluidDensityRin%g" % iStep: rho[0], "/FluidMacro/FluidDensityBin%g" % iStep: rho[1]}
        for name, a in zip("XYZ", u):
            arrays["/FluidVelocity/FluidVelocity%sAt%g" % (name, iStep)] = a
        self._results.write(iStep, arrays)

    def saveCycleInitial(self):
        """Writes the current state as `cycleInitialRK.h5` into ~/LBMInitial (LBM_INITIAL_DIR): the file the image-based
        drainage-imbibition restart reads (RKD2Q9.py:533-559; upstream it is prepared by hand from a result file)."""
        from .results import initial_dir
        os.makedirs(initial_dir(), exist_ok=True)
        f = ResultFile("cycleInitialRK.h5", directory=initial_dir())
        f.write(0, {"/FluidMacro/FluidDensityR": self.fluidsRhoR, "/FluidMacro/FluidDensityB": self.fluidsRhoB,
                    "/FluidPDF/FluidPDFR": self.fluidPDFR, "/FluidPDF/FluidPDFB": self.fluidPDFB,
                    "/FluidVelocity/FluidVelocityX": self.physicalVX, "/FluidVelocity/FluidVelocityY": self.physicalVY},
                single_file=True)

    def plotDensityDistributionOPT(self, iStep):
        """RKD2Q9.py:959-975 (PNG snapshots; skipped when matplotlib is absent)"""
        if self.slabs is not None and self.slabs.rank != 0:
            return
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        from .results import results_dir
        for name, a in (("FluidRsDistributionAt%05d.png", self.fluidsRhoR), ("FluidBsDistributionAt%05d.png", self.fluidsRhoB)):
            img = a if a.ndim == 2 else a[:, a.shape[1] // 2, :]
            plt.imshow(img, origin="lower"); plt.colorbar()
            plt.savefig(os.path.join(results_dir(), name % iStep)); plt.close()

    # -- the run loop (RKD2Q9.py:1225-1490) ------------------------------------------------------------
    def runRKColorGradient2DCSF(self):
        self._say("Start to run R-K color gradient lattice Boltzmann method.")
        self.initializeDomainBorder()
        self.initializeDomainCondition()
        self._make_engine()
        self.optimizeFluidandSolidArray()
        sl = self._slab
        self.engine.upload_state([self.fluidPDFR[sl], self.fluidPDFB[sl]], [self.fluidsRhoR[sl], self.fluidsRhoB[sl]])
        iStep = 0
        recordStep = 0
        t0 = time.perf_counter()
        # `asyncOutput` (attribute, or LBM_ASYNC_OUTPUT=1): records of densities + velocities are copied and written
        # behind the step loop (results.AsyncMacroOutput) instead of the reference's blocking copy + append
        out = None
        if getattr(self, "asyncOutput", os.environ.get("LBM_ASYNC_OUTPUT") == "1") and self.slabs is None:
            from .results import AsyncMacroOutput
            out = AsyncMacroOutput(self.engine, self._write_macro_record)
        while iStep < self.timeSteps:
            if iStep % self.timeInterval == 0:                       # RKD2Q9.py:1382-1393
                if out is not None:
                    out.snapshot()
                else:
                    self.convertOptTo2D()
                    self.resultInHDF5(recordStep)
                    self.plotDensityDistributionOPT(recordStep)
                    m = self.engine.total_mass()
                    if self.slabs is not None:
                        m = self.slabs.sum(m)
                    self._say("step %d: mass R %.12g, mass B %.12g" % (iStep, m[0], m[1]))
                recordStep += 1
            n = min(self.timeInterval - iStep % self.timeInterval, self.timeSteps - iStep)
            self.engine.step(n)
            iStep += n
        self.engine.synchronize()
        if out is not None:
            out.close()
        dt = time.perf_counter() - t0
        self.convertOptTo2D()
        self._say("%d steps, %.3f s, %.1f MLUPS (output included)" % (self.timeSteps, dt, self.voidSpace * self.timeSteps / dt / 1e6))

    def runRKColorGradient2DPerturbation(self):
        """RKD2Q9.py:979-1223: same host loop, the engine runs the perturbation operator (LBM_ST_PERTURBATION).  The first
        record is the state after the streaming the reference's loop starts with (:1048-1059)."""
        if self.surfaceTensionType != "'Perturbation'":
            raise IniError("SurfaceTensionType is not 'Perturbation'")
        self.runRKColorGradient2DCSF()

    def runRKColorGradient2D(self):
        """RKD2Q9.py:1495-1498"""
        if self.surfaceTensionType == "'CSF'":
            self.runRKColorGradient2DCSF()
        elif self.surfaceTensionType == "'Perturbation'":
            self.runRKColorGradient2DPerturbation()

    runModifiedRKColorGradient2D = runRKColorGradient2D     # the name main.py:53 calls
