"""Import shim used ONLY by the golden-vector generators in this directory.

It lets the reference's own modules (mounted read-only at /root/reference,
commit 3d84189) be imported in the build container so that their kernels and
host methods can be executed verbatim:

* kernels run under Numba's CUDA simulator (NUMBA_ENABLE_CUDASIM=1),
* modules that no longer exist upstream (numba.autojit, scipy.arange/empty/sqrt,
  matplotlib, tables, SimpleGeometryRK) are replaced by inert stand-ins so
  that `import RKD2Q9` / `import ShanChenD2Q9` succeed.  None of the stand-ins
  takes part in any arithmetic that ends up in a golden vector.

Nothing here is imported by the product, the tests or bench.py: the reference
tree does not exist on the GPU box.  The generators write .npz fixtures which
ARE committed.
"""
import os
import sys
import types

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
REF = os.environ.get("OPENLBMPM_REFERENCE", "/root/reference")
for sub in ("RKCG2D", "ShanChen2D"):
    p = os.path.join(REF, sub)
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import scipy as sp
import numba


def _identity_decorator(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda f: f


def install():
    # numba.autojit was removed; `@jit` on methods taking self must be a no-op too
    numba.autojit = _identity_decorator
    numba.jit = _identity_decorator
    for name in ("arange", "empty", "sqrt", "zeros", "ones", "array"):
        if not hasattr(sp, name):
            setattr(sp, name, getattr(np, name))
    if not hasattr(np, "bool"):
        np.bool = np.bool_
    for modname in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "tables"):
        if modname not in sys.modules:
            m = types.ModuleType(modname)
            m.use = lambda *a, **k: None
            sys.modules[modname] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if "SimpleGeometryRK" not in sys.modules:
        m = types.ModuleType("SimpleGeometryRK")
        m.defineGeometry = lambda x, y: (np.ones((y, x), bool), np.zeros((y, x), bool))
        sys.modules["SimpleGeometryRK"] = m


install()
