"""`ShanChenD3Q19` -- the 3-D Shan-Chen class the reference's main.py imports and drives
(`from ShanChenD3Q19 import ShanChenD3Q19`, main.py:17,73-77: `.runEFS4LBM3DGPU()`, `.runOriginalSC3DGPU()`)
but never shipped.  It is the D3Q19 instantiation of the SAME lattice-generic operators that reproduce the
reference's D2Q9 Shan-Chen vectors (sc_ops.cuh; specification in oracle/sc_dense.py: interaction weights w_q
(original) / 3 w_q (explicit forcing), solid weights w_q, d'Humieres MRT basis, isotropy 4; open boundaries along z:
per-component Zou-He velocity inlet, Zou-He pressure / convective outlet in their Hecht-Harting form).
Input contract, mirroring the 2-D class: twophasesetup.ini ([SeparationBorder] xGrid, yGrid, zGrid) plus
efs3D.ini | shanchen3D.ini (falling back to efs2D.ini | shanchen2D.ini, same sections); geometry from
`SimpleGeometry.defineGeometry3D(x, y, z)`.  Arrays are `[zGrid, yGrid, xGrid]`, populations `[nf, z, y, x, 19]`."""
import os

import numpy as np

from . import _lib
from .inifile import Ini, IniError
from .ShanChenD2Q9 import ShanChenD2Q9


class ShanChenD3Q19(ShanChenD2Q9):
    LATTICE = 19

    def _read_extra_dimensions(self, ini):
        self.nz = self.borderZ = ini.integer("SeparationBorder", "zGrid")

    def _set_lattice(self):
        e = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                      [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1],
                      [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1]], float)
        self.unitEX, self.unitEY, self.unitEZ = e[:, 0].copy(), e[:, 1].copy(), e[:, 2].copy()
        self.weightsCoeff = np.array([1. / 3.] + [1. / 18.] * 6 + [1. / 36.] * 12)

    def _model_ini(self, path, efs):
        for name in (("efs3D.ini", "efs2D.ini") if efs else ("shanchen3D.ini", "shanchen2D.ini")):
            if os.path.exists(os.path.join(path, name)):
                return Ini(path, name)
        raise IniError("no efs3D.ini / shanchen3D.ini (or their 2-D namesakes) in %s" % path)

    def _shape(self):
        return (self.nz, self.ny, self.nx)

    def _define_geometry(self):
        try:
            from SimpleGeometry import defineGeometry3D
        except ImportError:
            from .SimpleGeometry import defineGeometry3D
        return defineGeometry3D(self.nx, self.ny, self.nz)

    def _read_model(self, ini, section):
        super()._read_model(ini, section)
        if self.boundaryTypeInlet == "'Neumann'":       # the flow axis is z: [VelocityBoundary] velocityZ (velocityY accepted)
            self.velocityZInlet = np.array(ini.numbers("VelocityBoundary", "velocityZ", "velocityY"))
            self.velocityYInlet = self.velocityZInlet   # what the engine receives as the inlet velocity per component

    def _make_engine(self, model):
        if model == _lib.MODEL_EFS and self.explicitScheme != 4:
            raise IniError("D3Q19 explicit forcing: ExplicitScheme 4 (the higher-isotropy neighbour tables of the reference are 2-D)")
        super()._make_engine(model)

    def runEFS4LBM3DGPU(self):
        """main.py:77 -- explicit-forcing Shan-Chen (isotropy 4), SRT or MRT by [RelaxationType]"""
        self._run(_lib.MODEL_EFS, 1000)

    def runOriginalSC3DGPU(self):
        """main.py:76 -- original Shan-Chen"""
        self._run(_lib.MODEL_SC, 80)
