"""`Transport2D` (tracers on the Shan-Chen flow, reference main.py:15,56-62: ShanChen2D/Transport2D.py) is outside this
build; `Transport2DRK` (tracers on the colour-gradient flow) lives in Transport2DRK.py."""
from .Transport2DRK import Transport2DRK  # noqa: F401


class Transport2D:
    def __init__(self, pathIniFile):
        raise NotImplementedError("solute transport on the Shan-Chen flow is outside this build (SURVEY.md section 8, f-3); "
                                  "Transport2DRK runs tracers on the colour-gradient flow")
