"""oracle/cg_dense.py (lattice-generic, dense, used for D3Q19) instantiated for D2Q9 must reproduce
the reference's golden vectors; its D3Q19 instantiation must satisfy the physical invariants."""
import glob
import os

import numpy as np
import pytest

from oracle import cg2d, cg_dense

HERE = os.path.dirname(__file__)
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "cg2d_*.npz")))


def _load(path):
    g = np.load(path, allow_pickle=False)
    return g, dict(zip(g["params_keys"].tolist(), g["params_vals"].tolist()))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[5:-4] for p in GOLD])
def test_dense_d2q9_matches_reference(path):
    g, p = _load(path)
    sim = cg_dense.CGDense(cg_dense.d2q9(), g["is_domain"], sigma=float(p["sigma"]),
                           theta_deg=float(p["theta"]), wetting=int(p["wetting"]), beta=float(p["beta"]),
                           delta=float(p["delta"]), tauR=float(p["tauR"]), tauB=float(p["tauB"]),
                           tautype=int(p["tautype"]), relax=p["relax"], inlet=p["inlet"], outlet=p["outlet"],
                           v_inlet=float(p["vyb"]) + float(p["vyr"]), dBH=float(p["dBH"]), dRH=float(p["dRH"]),
                           dBL=float(p["dBL"]), dRL=float(p["dRL"]))
    red, dom, minor = g["red_mask"], g["is_domain"], float(g["minor"])
    sim.set_densities(np.where(red, float(p["rhoR"]), minor), np.where(red, minor, float(p["rhoB"])))
    import cases
    for s in range(cases.well_conditioned_snapshots(g, p) if "channel" in path else g["rhoR"].shape[0]):
        sim.head()
        for k, a in (("rhoR", sim.rhoR[0]), ("rhoB", sim.rhoB[0]), ("ux", sim.u[0, 0]), ("uy", sim.u[1, 0])):
            np.testing.assert_allclose(a, g[k][s], rtol=0, atol=1e-9 if "channel" in path else 5e-13, err_msg="%s snapshot %d" % (k, s))
        sim.body()


def test_d3q19_basis():
    L = cg_dense.d3q19()
    assert np.allclose(L.M @ L.M.T, np.diag(np.diag(L.M @ L.M.T)))        # orthogonal rows
    assert np.array_equal(L.e[L.opp], -L.e)
    assert abs(L.w.sum() - 1) < 1e-15
    # second moment of the weights = c_s^2 * identity
    m2 = np.einsum("i,ia,ib->ab", L.w, L.e.astype(float), L.e.astype(float))
    assert np.allclose(m2, np.eye(3) / 3)


def test_d3q19_slab_equals_d2q9():
    """A z-invariant D3Q19 run stays z-invariant and conserves each colour exactly-ish."""
    rng = np.random.default_rng(3)
    nz, ny, nx = 4, 12, 12
    dom = np.ones((nz, ny, nx), bool)
    base = 0.5 + 0.2 * (rng.random((ny, nx)) - 0.5)
    rhoR = np.broadcast_to(base, (nz, ny, nx)).copy()
    sim = cg_dense.CGDense(cg_dense.d3q19(), dom)
    sim.set_densities(rhoR, 1.0 - rhoR)
    m0 = sim.rhoR.sum(), sim.rhoB.sum()
    sim.step(6)
    assert np.abs(sim.rhoR - sim.rhoR[0:1]).max() < 1e-14
    assert abs(sim.rhoR.sum() - m0[0]) < 1e-10 and abs(sim.rhoB.sum() - m0[1]) < 1e-10


def test_d3q19_solid_sphere_mass_and_symmetry():
    n = 14
    z, y, x = np.mgrid[0:n, 0:n, 0:n]
    c = (n - 1) / 2
    dom = ((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2) > 2.6 ** 2
    red = (z < n // 2)
    sim = cg_dense.CGDense(cg_dense.d3q19(), dom, theta_deg=70.0)
    sim.set_densities(np.where(red, 1.0, 0.0), np.where(red, 0.0, 1.0))
    m0 = sim.rhoR.sum(), sim.rhoB.sum()
    sim.step(5)
    assert abs(sim.rhoR.sum() - m0[0]) < 1e-10 and abs(sim.rhoB.sum() - m0[1]) < 1e-10
    # mirror symmetry x <-> -x of the set-up is kept
    assert np.abs(sim.rhoR - sim.rhoR[:, :, ::-1]).max() < 1e-12
    assert np.abs(sim.rhoR - sim.rhoR.transpose(0, 2, 1)).max() < 1e-12
    assert np.isfinite(sim.rhoR).all()
