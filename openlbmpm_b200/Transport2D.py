"""Import-safe placeholder for the solute-transport drivers (`Transport2D`, `Transport2DRK`; reference
main.py:15-16): out of scope of the collision + streaming hot path."""


class Transport2D:
    def __init__(self, pathIniFile):
        raise NotImplementedError("solute transport is outside this build's hot-path scope (SURVEY.md section 8, f-3)")


Transport2DRK = Transport2D
