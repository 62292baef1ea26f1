"""CPU ORACLE (test infrastructure, NOT product code) -- ctypes wrapper of oracle/cg_c.c, the C + OpenMP
restatement of the reference's colour-gradient loop (RKD2Q9.py:1295-1490).  Lattice tables, node classes and
solid normals come from oracle/cg_dense.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm
may import this module."""
import ctypes
import os
import subprocess

import numpy as np

from . import cg_dense

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_cg.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "cg_c.c")):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = ctypes.CDLL(LIB)
        _lib.cgc_create.restype = ctypes.c_void_p
        _lib.cgc_create.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 10 + [ctypes.c_double] * 6 + [ctypes.c_int] * 4
        for name, args in (("cgc_destroy", [ctypes.c_void_p]), ("cgc_set_densities", [ctypes.c_void_p] * 3),
                           ("cgc_step", [ctypes.c_void_p, ctypes.c_int]), ("cgc_head", [ctypes.c_void_p]),
                           ("cgc_get", [ctypes.c_void_p] * 6)):
            getattr(_lib, name).restype = None
            getattr(_lib, name).argtypes = args
    return _lib


class CGC:
    def __init__(self, lattice, is_domain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98,
                 tauR=1.0, tauB=1.0, tautype=2, relax="MRT", threads=0):
        lib = load()
        L = self.L = cg_dense.d2q9() if lattice == 9 else cg_dense.d3q19()
        geo = cg_dense.CGDense(L, is_domain, theta_deg=theta_deg)      # node classes + solid normals only
        self.shape = geo.shape
        nz, ny, nx = geo.shape
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        keep = [np.ascontiguousarray(L.e, np.int64), np.ascontiguousarray(L.opp, np.int64),
                np.ascontiguousarray(L.w), np.ascontiguousarray(L.M), np.ascontiguousarray(L.Mi),
                np.ascontiguousarray(np.where(np.isnan(L.S_fixed), -1.0, L.S_fixed)),
                np.ascontiguousarray(geo.dom, np.uint8), np.ascontiguousarray(geo.wet_solid, np.uint8),
                np.ascontiguousarray(geo.near_solid, np.uint8), np.ascontiguousarray(geo.ns)]
        self._h = lib.cgc_create(L.D, L.Q, nx, ny, nz, *[p(a) for a in keep], sigma, theta_deg, beta, delta, tauR, tauB,
                                 tautype, wetting, 0 if relax == "SRT" else 1, threads)
        self.lib = lib

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.cgc_destroy(self._h)
            self._h = None

    def set_densities(self, rhoR, rhoB):
        r = np.ascontiguousarray(np.asarray(rhoR, float).reshape(self.shape))
        b = np.ascontiguousarray(np.asarray(rhoB, float).reshape(self.shape))
        self.lib.cgc_set_densities(self._h, r.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))

    def step(self, n=1):
        self.lib.cgc_step(self._h, int(n))

    def head(self):
        self.lib.cgc_head(self._h)

    def get(self, pdfs=False):
        N = int(np.prod(self.shape)); Q = self.L.Q
        rR = np.empty(self.shape); rB = np.empty(self.shape); u = np.empty((3,) + self.shape)
        fR = np.empty(self.shape + (Q,)) if pdfs else None
        fB = np.empty(self.shape + (Q,)) if pdfs else None
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
        self.lib.cgc_get(self._h, p(rR), p(rB), p(u), p(fR), p(fB))
        return dict(rhoR=rR, rhoB=rB, u=u, fR=fR, fB=fB)
