"""`Transport2DRK` -- solute tracers riding on the colour-gradient flow, the class the reference's main.py drives for
"transport" + "CG" (`Transport2DRK(ini).runTransport2DMPMCRK()`, main.py:66-68; RKCG2D/Transport2DRK.py).

Upstream neither of its drivers runs as shipped.  This class keeps the constructor contract
(`RKtwophasesetup2D.ini` for the flow + `transportsetup.ini`: [SystemType] Option = 'MPMC', NumberSchemes = 9 | 5, Reaction;
[TransportParameters] NumberTracers, Tau, BetaInterface, DiffusionJ; [TransportMRT] DiffusionX / Y / XY / YX; [RelaxationType]
Relaxation; [InitialCondition] TracerConc; 5-velocity branch: [Reaction] ReactionRate, [BoundaryCondition] InletType /
ConcentrationInlet / OutletType) and runs the tracer phase of `runTransport2DMPMCRKNew` (Transport2DRK.py:1341-1425) inside
liblbmpm.so (lbm_tracer_*): 9-velocity tracers on closed boxes, 5-velocity tracers (MRT, reaction A + B -> C, Inamuro inlet
row, free-flow outlet row) on closed boxes and open channels.
Public arrays: `tracerConc[numTracers, ny, nx]` next to the flow class's `fluidsRhoR/B`, `physicalVX/VY`."""
import time

import numpy as np

from . import _lib
from .inifile import Ini, IniError
from .results import ResultFile
from .RKD2Q9 import RKColorGradientLBM


class Transport2DRK(RKColorGradientLBM):
    def __init__(self, pathIniFile, verbose=True):
        ini = Ini(pathIniFile, "transportsetup.ini")
        self.systemType = ini.quoted("SystemType", "Option", default="MPMC")
        if self.systemType != "'MPMC'":
            raise IniError("only [SystemType] Option = 'MPMC' (tracers on the two-phase flow) is built")
        super().__init__(pathIniFile, verbose=verbose)
        self.reaction = ini.quoted("SystemType", "Reaction", default="no")
        self.numSchemes = ini.integer("SystemType", "NumberSchemes", default=9)
        if self.numSchemes not in (5, 9):
            raise IniError("NumberSchemes must be 5 or 9")
        if self.reaction == "'yes'" and self.numSchemes != 5:
            raise IniError("reactions between tracers act on the 5-velocity scheme of the reference only")
        self.numTracers = nt = ini.integer("TransportParameters", "NumberTracers", default=1)
        if not 1 <= nt <= 4:
            raise IniError("1..4 tracers are supported")
        if self.reaction == "'yes'" and nt != 3:
            raise IniError("the reaction of the reference is A + B -> C on three tracers (calReactionTracersGPU)")

        def per_tracer(section, key, default):
            vals = ini.numbers(section, key, default=default)
            if len(vals) == 1:
                vals = vals * nt
            if len(vals) != nt:
                raise IniError("The number of values of [%s] %s does not match the number of tracers in the domain." % (section, key))
            return np.array(vals, float)
        self.transportTau = per_tracer("TransportParameters", "Tau", "1.0")
        self.betaTracerArray = per_tracer("TransportParameters", "BetaInterface", "0.0")
        self.betaTracer = float(self.betaTracerArray[0])
        self.relaxationTypeTR = ini.quoted("RelaxationType", "Relaxation", default="SRT")
        self.diffusionX = per_tracer("TransportMRT", "DiffusionX", "0.1666666666666667")
        self.diffusionY = per_tracer("TransportMRT", "DiffusionY", "0.1666666666666667")
        self.diffusionXY = per_tracer("TransportMRT", "DiffusionXY", "0.0")
        self.diffusionYX = per_tracer("TransportMRT", "DiffusionYX", "0.0")
        self.initialTracerConc = per_tracer("InitialCondition", "TracerConc", "1.0")
        self.criteriaFluidRho = 0.5                               # Transport2DRK.py:1167
        if self.surfaceTensionType != "'CSF'":
            raise IniError("the tracers ride on the CSF flow (runTransport2DMPMCRKNew)")
        if self.numSchemes == 9:
            self.weightsCoeffTR = self.weightsCoeff.copy()
            if self.boundaryTypeInlet != "'Periodic'" or self.boundaryTypeOutlet != "'Periodic'":
                raise IniError("the 9-velocity tracer branch of the reference has no inlet / outlet treatment: closed boxes only")
        else:
            # Transport2DRK.py:313-347 (lattice, weights), :82-89 (reaction), :124-127 (J_0), :147-192 (tracer boundary rows)
            self.weightsCoeffTR = np.array([1. / 3., 1. / 6., 1. / 6., 1. / 6., 1. / 6.])
            self.unitVX = np.array([0., 1., -1., 0., 0.]); self.unitVY = np.array([0., 0., 0., 1., -1.])
            if self.relaxationTypeTR != "'MRT'":
                raise IniError("the 5-velocity branch of the reference collides with MRT only ([RelaxationType] Relaxation = 'MRT')")
            self.diffJ = per_tracer("TransportParameters", "DiffusionJ", "0.3333333333333333")
            self.reactionRate = np.array(ini.numbers("Reaction", "ReactionRate", default="0.0"), float) if self.reaction == "'yes'" else np.zeros(0)
            self.typeTracerInletBoundary = ini.quoted("BoundaryCondition", "InletType", default="Periodic")
            self.typeTracerOutletBoundary = ini.quoted("BoundaryCondition", "OutletType", default="Periodic")
            self.concTracerIn = per_tracer("BoundaryCondition", "ConcentrationInlet", "1.0")
        self._tracer_results = None

    def initializeTransportDomain(self):
        """Transport2DRK.py:399-470: concentration TracerConc below the inlet buffer rows, populations w_j C; assign
        `self.initialTracerRegion` (boolean [ny, nx]) beforehand for another layout"""
        shape = self._shape()
        reg = getattr(self, "initialTracerRegion", None)
        if reg is None:
            reg = np.indices(shape)[0] <= shape[0] - self.numBufferingLayers
        self.tracerConc = np.stack([np.where(reg, c, 0.0) for c in self.initialTracerConc]) * self.isDomain
        self.tracerDistr = self.tracerConc[..., None] * self.weightsCoeffTR

    def _make_engine(self):
        super()._make_engine()
        extra = {}
        if self.numSchemes == 5:
            extra = dict(n_schemes=5, reaction=self.reaction == "'yes'", reaction_rate=float(self.reactionRate[0]) if self.reactionRate.size else 0.0,
                         diff_j=self.diffJ, inlet_conc=self.concTracerIn,
                         inlet_type=_lib.TR_INLET_DIRICHLET if self.typeTracerInletBoundary == "'Dirichlet'" else _lib.TR_NONE,
                         # the reference's loop tests this spelling (Transport2DRK.py:1363; its parser mentions 'FreeFlow', :191)
                         outlet_type=_lib.TR_OUTLET_FREEFLOW if self.typeTracerOutletBoundary in ("'Freeflow'", "'FreeFlow'") else _lib.TR_NONE)
        self.engine.tracer_setup(n_tracers=self.numTracers, relax=_lib.RELAX_MRT if self.relaxationTypeTR == "'MRT'" else _lib.RELAX_SRT,
                                 tau=self.transportTau, dxx=self.diffusionX, dyy=self.diffusionY, dxy=self.diffusionXY,
                                 dyx=self.diffusionYX, beta=self.betaTracerArray, criterion=self.criteriaFluidRho, **extra)

    def saveConcentrationHDF5(self, index):
        """Transport2DRK.py:651-661"""
        if self.slabs is not None and self.slabs.rank != 0:
            return
        if self._tracer_results is None:
            self._tracer_results = ResultFile("ConcentrationResults.h5", groups=("TransportMacro",))
        self._tracer_results.write(index, {"/TransportMacro/TracerConcType%gin%g" % (i, index): self.tracerConc[i]
                                           for i in range(self.numTracers)})

    def runTransport2DMPMCRKNew(self):
        """Transport2DRK.py:1059-1485: flow and tracers advance together; records of both every TimeInterval steps"""
        self._say("Start to run R-K color gradient lattice Boltzmann method.")
        self.initializeDomainBorder()
        self.initializeDomainCondition()
        self._make_engine()
        self.optimizeFluidandSolidArray()
        self.initializeTransportDomain()
        sl = self._slab
        self.engine.upload_state([self.fluidPDFR[sl], self.fluidPDFB[sl]], [self.fluidsRhoR[sl], self.fluidsRhoB[sl]])
        self.engine.tracer_init(*[c[sl] for c in self.tracerConc])
        iStep = recordStep = 0
        t0 = time.perf_counter()
        while iStep < self.timeSteps:
            if iStep % self.timeInterval == 0:
                self.convertOptTo2D()
                self.resultInHDF5(recordStep)
                conc = self.engine.tracer_download()
                if self.slabs is not None:
                    conc = [self.slabs.gather(c) for c in conc]
                self.tracerConc = np.stack(conc)
                self.saveConcentrationHDF5(recordStep)
                recordStep += 1
                self._say("step %d: tracer totals %s" % (iStep, self.tracerConc.sum(axis=tuple(range(1, self.tracerConc.ndim)))))
            n = min(self.timeInterval - iStep % self.timeInterval, self.timeSteps - iStep)
            self.engine.step(n)
            iStep += n
        self.engine.synchronize()
        self.convertOptTo2D()
        conc = self.engine.tracer_download()
        self.tracerConc = np.stack([self.slabs.gather(c) for c in conc] if self.slabs is not None else conc)
        dt = time.perf_counter() - t0
        self._say("%d steps, %.3f s, %.1f MLUPS (flow + %d tracers, output included)" % (self.timeSteps, dt, self.voidSpace * self.timeSteps / dt / 1e6, self.numTracers))

    runTransport2DMPMCRK = runTransport2DMPMCRKNew        # the name main.py:68 calls
