"""oracle/cg_c.c (C + OpenMP restatement, the CPU baseline of bench.py) must agree with oracle/cg_dense.py,
which is pinned to the reference's golden vectors (test_oracle_dense.py)."""
import numpy as np
import pytest

from oracle import cg_c, cg_dense


def _compare(lattice, dom, rhoR, rhoB, steps, **par):
    L = cg_dense.d2q9() if lattice == 9 else cg_dense.d3q19()
    a = cg_dense.CGDense(L, dom, **par)
    a.set_densities(rhoR, rhoB)
    b = cg_c.CGC(lattice, dom, threads=4, **par)
    b.set_densities(np.where(dom, rhoR, 0.0), np.where(dom, rhoB, 0.0))
    a.step(steps); b.step(steps)
    a.head(); b.head()
    out = b.get(pdfs=True)
    np.testing.assert_allclose(out["rhoR"], a.rhoR, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["rhoB"], a.rhoB, rtol=0, atol=1e-12)
    for d in range(L.D):
        np.testing.assert_allclose(out["u"][d], a.u[d], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["fR"], np.moveaxis(a.fR, 0, -1), rtol=0, atol=1e-12)


@pytest.mark.parametrize("relax", ["MRT", "SRT"])
def test_d3q19_periodic(relax):
    rng = np.random.default_rng(1)
    n = (8, 10, 12)
    r = 0.5 + 0.3 * (rng.random(n) - 0.5)
    _compare(19, np.ones(n, bool), r, 1 - r, 5, relax=relax)


def test_d3q19_sphere_wetting():
    n = 12
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    c = (n - 1) / 2
    dom = ((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2) > 2.6 ** 2
    red = z < n // 2
    _compare(19, dom, np.where(red, 1.0, 0.0), np.where(red, 0.0, 1.0), 5, theta_deg=70.0)


@pytest.mark.parametrize("wetting,relax", [(2, "MRT"), (1, "SRT")])
def test_d2q9_block(wetting, relax):
    rng = np.random.default_rng(2)
    n = (20, 18)
    dom = np.ones(n, bool); dom[8:12, 6:11] = False
    r = 0.5 + 0.4 * (rng.random(n) - 0.5)
    _compare(9, dom, r, 1 - r, 6, wetting=wetting, relax=relax, theta_deg=50.0, tauR=0.9, tauB=1.2, tautype=1)
