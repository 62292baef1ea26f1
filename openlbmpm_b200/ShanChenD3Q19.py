"""`ShanChenD3Q19` -- named by the reference's main.py:17,73-77 but absent from its tree (there is no 3-D
Shan-Chen code upstream to be on par with).  Import-safe placeholder; SURVEY.md 8(f-1) lists it as "next"."""


class ShanChenD3Q19:
    def __init__(self, pathIniFile):
        self.path = pathIniFile

    def _missing(self):
        raise NotImplementedError("D3Q19 Shan-Chen is not part of the reference tree; the colour-gradient D3Q19 path "
                                  "(RKColorGradient3D) is the 3-D model of this build")

    def runEFS4LBM3DGPU(self):
        self._missing()

    def runOriginalSC3DGPU(self):
        self._missing()
