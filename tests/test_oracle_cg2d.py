"""Pins oracle/cg2d.py against golden vectors produced by the reference itself
(tests/golden/gen_goldens_cg2d.py: RKD2Q9.py driver + AcceleratedRKGPU2D.py
kernels run verbatim under Numba's CUDA simulator, reference commit 3d84189)."""
import glob
import os

import numpy as np
import pytest

from oracle import cg2d

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cg2d_*.npz")))


def load(path):
    g = np.load(path, allow_pickle=False)
    p = dict(zip(g["params_keys"].tolist(), g["params_vals"].tolist()))
    return g, p


def make_oracle(g, p):
    sim = cg2d.CG2D(g["is_domain"], sigma=float(p["sigma"]), theta_deg=float(p["theta"]),
                    wetting=int(p["wetting"]), beta=float(p["beta"]), delta=float(p["delta"]),
                    tauR=float(p["tauR"]), tauB=float(p["tauB"]), tautype=int(p["tautype"]),
                    relax=p["relax"], inlet=p["inlet"], outlet=p["outlet"],
                    vy_inlet=float(p["vyb"]) + float(p["vyr"]), dBH=float(p["dBH"]), dRH=float(p["dRH"]),
                    dBL=float(p["dBL"]), dRL=float(p["dRL"]))
    red = g["red_mask"]; dom = g["is_domain"]; minor = float(g["minor"])
    rhoR = np.where(dom, np.where(red, float(p["rhoR"]), minor), 0.0)
    rhoB = np.where(dom, np.where(red, minor, float(p["rhoB"])), 0.0)
    sim.set_densities(rhoR, rhoB)
    return sim


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_indexing_bit_exact(path):
    g, p = load(path)
    idx = cg2d.build_indexing(g["is_domain"])
    for k in ("fluidNodes", "neighboringNodes", "wettingSolidNodes", "neighboringWettingSolidNodes"):
        assert np.array_equal(idx[k], g[k]), k
    if "nsX" in g.files:
        for k in ("fluidNodesWithSolidGPU", "fluidNodesWithSolidOriginal"):
            assert np.array_equal(idx[k], g[k]), k
        np.testing.assert_allclose(idx["nsX"], g["nsX"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(idx["nsY"], g["nsY"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_trajectory_matches_reference(path):
    g, p = load(path)
    sim = make_oracle(g, p)
    nsnap = g["rhoR"].shape[0]
    for s in range(nsnap):
        sim.head()                      # the reference snapshots after BC + velocity (RKD2Q9.py:1382)
        d = sim.to_dense()
        for k in ("rhoR", "rhoB", "ux", "uy"):
            np.testing.assert_allclose(d[k], g[k][s], rtol=0, atol=2e-13, err_msg="%s snapshot %d" % (k, s))
        if s == 0:
            np.testing.assert_allclose(d["pdfR"], g["pdfR_first"], rtol=0, atol=2e-13)
        if s == nsnap - 1:
            np.testing.assert_allclose(d["pdfR"], g["pdfR_last"], rtol=0, atol=2e-13)
            np.testing.assert_allclose(d["pdfB"], g["pdfB_last"], rtol=0, atol=2e-13)
        sim.body()
