"""`RKColorGradient3D` -- the D3Q19 colour-gradient class the reference's main.py imports
(`from RKColorGradientD3Q19 import RKColorGradient3D`, main.py:22,78-81) but never shipped.  Input contract:
IniFiles/RKtwophasesetup3D.ini (xDomain, yDomain, zDomain; flow along z: velocityZR/ZB, densityBL/RL;
[TimeSteps] TimeSteps).  The physics is the generalisation of the 2-D CSF path fixed in DESIGN.md
(d'Humieres D3Q19 basis, K = n_a n_b d_a n_b - d_a n_a, Akai-2018 wetting in the plane of (n_s, n*)).
Arrays are `[zDomain, yDomain, xDomain]`; populations `[z, y, x, 19]`."""
import numpy as np

from .RKD2Q9 import RKColorGradientLBM


class RKColorGradient3D(RKColorGradientLBM):
    INI_NAME = "RKtwophasesetup3D.ini"
    LATTICE = 19

    def _read_domain(self, ini):
        self.xDomain = ini.integer("DomainSize", "xDomain")
        self.yDomain = ini.integer("DomainSize", "yDomain")
        self.zDomain = ini.integer("DomainSize", "zDomain")
        self.numBufferingLayers = ini.integer("DomainSize", "numBufferingLayers", default=0)
        self.ratioTopToBottom = ini.number("DomainSize", "ratioTopToBottom", default=0.5)

    def _default_surface_tension(self, ini):
        """The reference's 3-D ini has no [SurfaceTension] section: it parameterises the perturbation operator
        ([RKParameters] AkR / AkB, [BoundariesSetup] SolidRhoR / SolidRhoB; RKtwophasesetup3D.ini:9-25)."""
        if not ini.has_section("SurfaceTension") and ini.number("RKParameters", "AkR", default=0.0) > 0.0:
            return "Perturbation"
        return "CSF"

    def _default_solid_phi(self, ini):
        """fictitious colour densities on the solid -> the colour its neighbours see"""
        r = ini.number("BoundariesSetup", "SolidRhoR", default=0.0); b = ini.number("BoundariesSetup", "SolidRhoB", default=0.0)
        return (r - b) / (r + b) if r + b > 0.0 else 0.0

    def _read_time(self, ini):
        sec = "TimeSteps" if ini.has_section("TimeSteps") else "TimeSetup"
        self.timeSteps = ini.integer(sec, "TimeSteps")
        self.timeInterval = ini.integer(sec, "TimeInterval", default=max(1, self.timeSteps))

    def _read_boundaries(self, ini):
        self.boundaryTypeInlet = ini.quoted("BoundaryCondition", "BoundaryTypeInlet", default="Periodic")
        self.boundaryTypeOutlet = ini.quoted("BoundaryCondition", "BoundaryTypeOutlet", default="Periodic")
        self.velocityZR = ini.number("BoundaryCondition", "velocityZR", default=0.0)
        self.velocityZB = ini.number("BoundaryCondition", "velocityZB", default=0.0)
        self.densityRhoBH = ini.number("BoundaryCondition", "densityBH", default=0.0)
        self.densityRhoRH = ini.number("BoundaryCondition", "densityRH", default=0.0)
        self.densityRhoBL = ini.number("BoundaryCondition", "densityBL", default=0.0)
        self.densityRhoRL = ini.number("BoundaryCondition", "densityRL", default=0.0)

    def _inlet_velocity(self):
        return self.velocityZB + self.velocityZR

    def _shape(self):
        return (self.zDomain, self.yDomain, self.xDomain)

    def _weights(self):
        return np.array([1. / 3.] + [1. / 18.] * 6 + [1. / 36.] * 12)

    def initializeDomainBorder(self):
        try:
            from SimpleGeometryRK import defineGeometry3D
        except ImportError:
            from .SimpleGeometryRK import defineGeometry3D
        self.isDomain, self.isSolid = defineGeometry3D(self.xDomain, self.yDomain, self.zDomain)
        self.isDomain = np.ascontiguousarray(self.isDomain, dtype=bool)
        self.isSolid = ~self.isDomain
        self.voidSpace = int(np.count_nonzero(self.isDomain))
        self._say('The number of vexls in void space is %g.' % self.voidSpace)

    def runRKColorGradient3D(self):
        self.runRKColorGradient2DCSF()
