"""openlbmpm_b200 -- B200-native collision + streaming hot path of openLBMPM behind the reference's class
surface: Python host code -> ctypes -> liblbmpm.so (hand-written sm_100a kernels).  See DESIGN.md."""
from ._lib import Engine, LbmError  # noqa: F401
