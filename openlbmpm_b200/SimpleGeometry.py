"""User-editable geometry hook with the reference's contract (ShanChen2D/SimpleGeometry.py:11-27):
`defineGeometry(xDomain, yDomain) -> (isDomain, isSolid)`, boolean `[yDomain, xDomain]` arrays.
Default layout as shipped by the reference: side walls at x = 0 and x = xDomain-1 except in the 10 rows
next to the inlet and the outlet."""
import numpy as np


def defineGeometry(xDomain, yDomain):
    isSolid = np.zeros((yDomain, xDomain), dtype=bool)
    if yDomain > 20:
        isSolid[10:yDomain - 10, [0, xDomain - 1]] = True
    return ~isSolid, isSolid


def defineGeometry3D(xDomain, yDomain, zDomain):
    """3-D counterpart (the reference ships none): all void, `[zDomain, yDomain, xDomain]`."""
    isSolid = np.zeros((zDomain, yDomain, xDomain), dtype=bool)
    return ~isSolid, isSolid
