"""`ShanChenD2Q9` -- drop-in for the reference's Shan-Chen host class (ShanChen2D/ShanChenD2Q9.py:38-2094).

Constructor `ShanChenD2Q9(pathIniFile)` reads twophasesetup.ini + shanchen2D.ini | efs2D.ini (basicsetup.ini,
which the reference demands but does not ship, is optional); entry points `runTypeSCmodel`,
`runOptimizedLBM`, `runOptimizedEFLBM`; public arrays `isDomain`, `isSolid`, `fluidsDensity[nf, ny, nx]`,
`fluidPDF[nf, ny, nx, 9]`, `physicalVX/VY`, `fluidNodes`, `neighboringNodes`, `optFluidRho`, `optFluidPDF`.
The per-step loops (ShanChenD2Q9.py:1492-1629, 1852-2087; 15-20 kernel launches each) run inside liblbmpm.so."""
import os
import time

import numpy as np

from . import _lib
from .inifile import Ini, IniError
from .results import ResultFile


class ShanChenD2Q9:
    LATTICE = 9

    def __init__(self, pathIniFile, verbose=True):
        self.path = pathIniFile
        self.verbose = verbose
        ini = Ini(pathIniFile, "twophasesetup.ini")
        self.PictureExistance = ini.quoted("PictureSetup", "Exist", default="no")
        self.nx = self.borderX = ini.integer("SeparationBorder", "xGrid")
        self.ny = self.borderY = ini.integer("SeparationBorder", "yGrid")
        self._read_extra_dimensions(ini)
        self.typesFluids = ini.integer("FluidsTypes", "NumberOfFluids", default=2)
        if not 1 <= self.typesFluids <= 4:
            raise IniError("1..4 fluids are supported")
        self.interactionType = ini.quoted("InterType", "InteractionType", default="ShanChen")
        self.Parallel = ini.quoted("Parallelism", "Parallel", default="yes")
        self.xDimension = ini.integer("Parallelism", "xDimension", default=128)
        self.threadNum = ini.integer("Parallelism", "ThreadsNum", default=32)
        self.relaxationType = ini.quoted("RelaxationType", "Type", default="SRT")
        if self.relaxationType == "'TRT'":
            raise IniError("TRT is read by the reference but never launched by a live driver")
        self.duplicateDomain = ini.quoted("DuplicateDomain", "Option", default="no")
        self.isCycles = ini.quoted("DICycles", "Option", default="no")
        if self.isCycles == "'yes'":
            self.lastStep = ini.integer("DICycles", "LastStep")
        self._set_lattice()
        efs = self.interactionType == "'EFS'"
        self._read_model(self._model_ini(pathIniFile, efs), "EFSParameters" if efs else "ShanChenParameters")
        self.engine = None
        self._results = None
        self.slabs = None        # slab decomposition under `torchrun` (slab.from_environment), see RKD2Q9.py

    # -- lattice-specific pieces (ShanChenD3Q19 overrides them) ------------------------------------
    def _read_extra_dimensions(self, ini):
        pass

    def _set_lattice(self):
        self.unitEX = np.array([0., 1., 0., -1., 0., 1., -1., -1., 1.])
        self.unitEY = np.array([0., 0., 1., 0., -1., 1., 1., -1., -1.])
        self.weightsCoeff = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)

    def _model_ini(self, path, efs):
        return Ini(path, "efs2D.ini" if efs else "shanchen2D.ini")

    def _shape(self):
        return (self.ny, self.nx)

    def _define_geometry(self):
        try:
            from SimpleGeometry import defineGeometry
        except ImportError:
            from .SimpleGeometry import defineGeometry
        return defineGeometry(self.nx, self.ny)

    def _say(self, *a):
        if self.verbose and (self.slabs is None or self.slabs.rank == 0):
            print(*a)

    def _read_model(self, ini, section):
        """ShanChenD2Q9.py:164-321 / 323-499"""
        nf = self.typesFluids
        self.initialDensities = np.array(ini.numbers("FluidProperties", "InitialDensities"))
        self.backgroundDensities = np.array(ini.numbers("FluidProperties", "BackgroundDensities"))
        self.tau = np.array(ini.numbers("FluidProperties", "FluidsTau"))
        if not (self.initialDensities.size == self.backgroundDensities.size == self.tau.size == nf):
            raise IniError("The number of fluids does not match the number of densities / taus.")
        pairs = ini.numbers(section, "InteractionFluid")
        self.interCoeff = np.zeros((nf, nf))
        k = 0
        for i in range(nf - 1):
            for j in range(i + 1, nf):
                self.interCoeff[i, j] = self.interCoeff[j, i] = pairs[k]
                k += 1
        self.interactionSolid = np.array(ini.numbers(section, "InteractionSolid"))
        if self.interactionSolid.size != nf:
            raise IniError("The number of fluids does not match the number of interaction coeff with solid.")
        self.explicitScheme = ini.integer("ForceScheme", "ExplicitScheme", default=4)
        if self.interactionType == "'EFS'" and self.explicitScheme not in (4, 8, 10):
            raise IniError("ExplicitScheme must be 4, 8 or 10")
        self.boundaryTypeInlet = ini.quoted("BoundaryDefinition", "BoundaryTypeInlet", default="Periodic")
        self.boundaryMethod = ini.quoted("BoundaryDefinition", "BoundaryMethod", default="ZouHe")
        self.boundaryTypeOutlet = ini.quoted("BoundaryDefinition", "BoundaryTypeOutlet", default="Periodic")
        self.velocityYInlet = np.zeros(nf); self.velocityXInlet = np.zeros(nf)
        if self.boundaryTypeInlet == "'Neumann'":
            if self.boundaryMethod != "'ZouHe'":
                raise IniError("BoundaryMethod 'Chang' is not built; use 'ZouHe'")
            self.velocityXInlet = np.array(ini.numbers("VelocityBoundary", "velocityX"))
            self.velocityYInlet = np.array(ini.numbers("VelocityBoundary", "velocityY"))
        elif self.boundaryTypeInlet == "'Dirichlet'":
            raise IniError("The reference's pressure inlet reads self.specificRho1Upper, which is never set "
                           "(ShanChenD2Q9.py:1498): that path cannot run upstream either")
        if (self.interactionType == "'EFS'" and self.explicitScheme == 10 and
                (self.boundaryTypeInlet != "'Periodic'" or self.boundaryTypeOutlet != "'Periodic'")):
            raise IniError("the reference's explicit-forcing loop has open boundaries for ExplicitScheme 4 and 8 only "
                           "(ShanChenD2Q9.py:1933-2014)")
        self.numTimeStep = ini.integer("Time", "numberTimeStep")

    # -- geometry / initial condition --------------------------------------------------------------
    def _process_image(self):
        """ShanChenD2Q9.py:542-585: crop to the solid pixels, optional mirror tiling ([DuplicateDomain] Option = 'yes'; the
        two counts the reference asks for with input() come from `self.duplicateX / duplicateY`, LBM_DUPLICATE_X / _Y or
        the prompt), solid side columns, 20 void buffer rows at either end"""
        import os
        from . import imagegeo
        if self.LATTICE != 9:
            raise IniError("image input is a 2-D feature of the reference")
        try:
            img = imagegeo.crop_to_solid(imagegeo.load_gray())
        except (FileNotFoundError, ValueError) as e:
            raise IniError(str(e))
        if self.duplicateDomain == "'yes'":
            def count(attr, env, prompt):
                v = getattr(self, attr, None) or os.environ.get(env)
                return int(v if v else input(prompt))
            img = imagegeo.expand_image_domain(img, count("duplicateX", "LBM_DUPLICATE_X", "Number of duplication in x direction: "),
                                               count("duplicateY", "LBM_DUPLICATE_Y", "Number of duplication in y direction: "))
        self.effectiveDomain = imagegeo.close_and_pad(img, 20, 20)
        self.ny, self.nx = self.effectiveDomain.shape
        self.originalXdim = self.nx
        self._say('Now the size of domain is %g and %g' % (self.ny, self.nx))
        dom = imagegeo.to_domain(self.effectiveDomain)
        return dom, ~dom

    def initializeDomainBorder(self):
        self.isDomain, self.isSolid = self._process_image() if self.PictureExistance == "'yes'" else self._define_geometry()
        self.isDomain = np.ascontiguousarray(self.isDomain, dtype=bool)
        self.isSolid = ~self.isDomain
        self.voidSpace = int(np.count_nonzero(self.isDomain))
        self._say('The porosity of the layout is %f.' % (self.voidSpace / self.isDomain.size))

    def initializeDomainCondition(self):
        """ShanChenD2Q9.py:734-768: fluid 0 below row ny-10, fluid 1 above; assign `self.initialRegion0`
        (boolean [ny, nx]) beforehand for another layout"""
        reg = getattr(self, "initialRegion0", None)
        shape = self._shape()
        image = self.PictureExistance == "'yes'"
        if reg is None:
            reg = np.indices(shape)[0] < shape[0] - (20 if image else 10)      # :757 / :771
        nf = self.typesFluids
        self.fluidsDensity = np.zeros((nf,) + shape)
        if image and self.isCycles == "'yes'":
            return self._initialize_from_previous_run()
        for k in range(nf):
            inside = self.initialDensities[k] if k == 0 else self.backgroundDensities[k]
            outside = self.backgroundDensities[k] if k == 0 else self.initialDensities[k]
            self.fluidsDensity[k] = np.where(reg, inside, outside) * self.isDomain
        self.fluidPDF = self.fluidsDensity[..., None] * self.weightsCoeff
        self.physicalVX = np.zeros(shape); self.physicalVY = np.zeros(shape)
        if self.LATTICE == 19:
            self.physicalVZ = np.zeros(shape)

    def _initialize_from_previous_run(self):
        """ShanChenD2Q9.py:788-817: the fluids of the previous run keep their densities of record LastStep (background
        in the 30 inlet rows), the NEW fluid enters through those rows; populations at rest"""
        from .results import initial_dir, read_arrays
        nf, shape = self.typesFluids, self._shape()
        names = ["/FluidMacro/FluidDensityType%gin%d" % (k, self.lastStep) for k in range(nf - 1)]
        try:
            d = read_arrays(initial_dir(), "SimulationResults.h5", names)
        except FileNotFoundError as e:
            raise IniError("There is no file for initializing the domain: %s" % e)
        for k in range(nf - 1):
            old = np.array(d[names[k]], float)
            if old.shape != shape:
                raise IniError("the restart file holds arrays of shape %s, the domain is %s" % (old.shape, shape))
            self.fluidsDensity[k, :-30] = old[:-30]
            self.fluidsDensity[k, -30:] = self.backgroundDensities[k]
        rows = np.indices(shape)[0]
        self.fluidsDensity[-1] = np.where(rows < shape[0] - 30, self.backgroundDensities[-1], self.initialDensities[-1])
        self.fluidsDensity *= self.isDomain
        self.fluidPDF = self.fluidsDensity[..., None] * self.weightsCoeff
        self.physicalVX = np.zeros(shape); self.physicalVY = np.zeros(shape)

    def _make_engine(self, model):
        inlet = {"'Periodic'": _lib.BC_PERIODIC, "'Neumann'": _lib.INLET_VELOCITY}.get(self.boundaryTypeInlet)
        outlet = {"'Periodic'": _lib.BC_PERIODIC, "'Convective'": _lib.OUTLET_CONVECTIVE,
                  "'Dirichlet'": _lib.OUTLET_PRESSURE}.get(self.boundaryTypeOutlet)
        if inlet is None or outlet is None:
            raise IniError("Unknown boundary type %s / %s" % (self.boundaryTypeInlet, self.boundaryTypeOutlet))
        if model == _lib.MODEL_SC and outlet == _lib.OUTLET_PRESSURE:
            raise IniError("the original Shan-Chen loop has no pressure outlet (ShanChenD2Q9.py:1603-1621)")
        G = np.zeros((4, 4)); G[:self.typesFluids, :self.typesFluids] = self.interCoeff
        from . import slab
        if self.slabs is None:
            self.slabs = slab.from_environment()
        shape = self._shape()
        self._slab = slice(None)
        if self.slabs is not None:
            lo, hi = slab.slab_bounds(shape[0], self.slabs.rank, self.slabs.world)
            self._slab = slice(lo, hi)
            shape = (hi - lo,) + tuple(shape[1:])
        self.engine = _lib.Engine(self.LATTICE, shape, model=model,
                                  device=self.slabs.device_index if self.slabs is not None else 0,
                                  relax=_lib.RELAX_MRT if self.relaxationType == "'MRT'" else _lib.RELAX_SRT,
                                  n_components=self.typesFluids, inlet=inlet, outlet=outlet, sc_tau=self.tau,
                                  sc_G=G.ravel(), sc_Gsolid=self.interactionSolid, sc_inlet_velocity=self.velocityYInlet,
                                  sc_rho_out=[1.0, 0.02],      # hard-coded upstream: OptimizedD2Q9GPU.py:560-561
                                  sc_isotropy=self.explicitScheme if model == _lib.MODEL_EFS else 4)
        if self.slabs is not None:
            self.engine.comm_init(self.slabs.rank, self.slabs.world, self.slabs.unique_id(self.engine))
        self.engine.set_geometry(self.isDomain[self._slab])

    def optimizeFluidArray(self):
        """ShanChenD2Q9.py:587-659: compact node list and neighbour table (solids are -1 in this class)"""
        if self.slabs is not None:
            return          # the compact numbering belongs to the whole lattice; a slab only knows its own nodes
        idx = self.engine.export_indexing()
        self.fluidNodes = idx["fluidNodes"]
        nb = idx["neighboringNodes"].copy()
        nb[nb < 0] = -1
        self.neighboringNodes = nb
        self.optFluidRho = self.fluidsDensity.reshape(self.typesFluids, -1)[:, self.fluidNodes]
        self.optFluidPDF = self.fluidPDF.reshape(self.typesFluids, -1, self.LATTICE)[:, self.fluidNodes]

    def convertOptTo2D(self):
        rho, u = self.engine.download_macros()
        pdf = self.engine.download_pdfs()
        if self.slabs is not None:          # every rank receives the whole lattice
            rho = [self.slabs.gather(a) for a in rho]; u = [self.slabs.gather(a) for a in u]
            pdf = [self.slabs.gather(a) for a in pdf]
        self.fluidsDensity = np.stack(rho)
        self.physicalVX, self.physicalVY = u[0], u[1]
        if self.LATTICE == 19:
            self.physicalVZ = u[2]
        self.fluidPDF = np.stack(pdf)

    def resultInHDF5(self, iStep):
        """ShanChenD2Q9.py:940-955"""
        if self.slabs is not None and self.slabs.rank != 0:
            return                                  # rank 0 writes the gathered arrays
        if self._results is None:
            self._results = ResultFile("SimulationResults.h5", groups=("FluidMacro", "FluidVelocity"))
        arrays = {"/FluidMacro/FluidDensityType%gin%g" % (k, iStep): self.fluidsDensity[k] for k in range(self.typesFluids)}
        arrays["/FluidVelocity/FluidVelocityXAt%g" % iStep] = self.physicalVX
        arrays["/FluidVelocity/FluidVelocityYAt%g" % iStep] = self.physicalVY
        if self.LATTICE == 19:
            arrays["/FluidVelocity/FluidVelocityZAt%g" % iStep] = self.physicalVZ
        self._results.write(iStep, arrays)

    def _write_macro_record(self, iStep, rho, u):
        if self._results is None:
            self._results = ResultFile("SimulationResults.h5", groups=("FluidMacro", "FluidVelocity"))
        arrays = {"/FluidMacro/FluidDensityType%gin%g" % (k, iStep): rho[k] for k in range(self.typesFluids)}
        for name, a in zip("XYZ", u):
            arrays["/FluidVelocity/FluidVelocity%sAt%g" % (name, iStep)] = a
        self._results.write(iStep, arrays)

    def _run(self, model, interval):
        self.initializeDomainBorder()
        self.initializeDomainCondition()
        self._make_engine(model)
        self.optimizeFluidArray()
        sl = self._slab
        self.engine.upload_state([f[sl] for f in self.fluidPDF], [r[sl] for r in self.fluidsDensity])
        step = record = 0
        total = self.numTimeStep + 1                 # both loops run numTimeStep + 1 iterations
        t0 = time.perf_counter()
        out = None          # `asyncOutput` / LBM_ASYNC_OUTPUT=1: records are copied and written behind the step loop
        if getattr(self, "asyncOutput", os.environ.get("LBM_ASYNC_OUTPUT") == "1") and self.slabs is None:
            from .results import AsyncMacroOutput
            out = AsyncMacroOutput(self.engine, self._write_macro_record)
        while step < total:
            if step % interval == 0:
                if out is not None:
                    out.snapshot()
                else:
                    self.convertOptTo2D()
                    self.resultInHDF5(record)
                    m = self.engine.total_mass()
                    self._say("step %d: masses %s" % (step, self.slabs.sum(m) if self.slabs is not None else m))
                record += 1
            n = min(interval - step % interval, total - step)
            self.engine.step(n)
            step += n
        self.engine.synchronize()
        if out is not None:
            out.close()
        dt = time.perf_counter() - t0
        self.convertOptTo2D()
        self._say("%d steps, %.3f s, %.1f MLUPS (output included)" % (total, dt, self.voidSpace * total / dt / 1e6))

    def runOptimizedLBM(self):
        """ShanChenD2Q9.py:1433-1629 (results every 80 iterations, :1561)"""
        self._run(_lib.MODEL_SC, 80)

    def runOptimizedEFLBM(self):
        """ShanChenD2Q9.py:1631-2087 (results every 1000 iterations, :2029)"""
        self._run(_lib.MODEL_EFS, 1000)

    def runTypeSCmodel(self):
        """ShanChenD2Q9.py:2089-2094"""
        if self.interactionType == "'ShanChen'":
            self.runOptimizedLBM()
        elif self.interactionType == "'EFS'":
            self.runOptimizedEFLBM()
