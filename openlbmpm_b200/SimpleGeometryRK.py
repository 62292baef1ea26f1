"""`SimpleGeometryRK` is imported by the reference (RKCG2D/RKD2Q9.py:21) but is not in its tree; it has the
same contract as ShanChen2D/SimpleGeometry.py."""
from .SimpleGeometry import defineGeometry, defineGeometry3D  # noqa: F401
