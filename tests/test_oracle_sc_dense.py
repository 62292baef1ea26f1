"""oracle/sc_dense.py (lattice-generic Shan-Chen step, used for D3Q19) instantiated for D2Q9 must reproduce the
reference's golden vectors for the closed-box cases; its D3Q19 instantiation must satisfy the invariants."""
import numpy as np
import pytest

import cases
from oracle import sc_dense

CLOSED = [p for p in cases.GOLD_SC2D if "iso" not in p]        # isotropy 8 / 10 are 2-D neighbour tables (oracle/sc2d.py)


@pytest.mark.parametrize("path", CLOSED, ids=[cases.gold_id(p) for p in CLOSED])
def test_d2q9_matches_reference(path):
    g, p = cases.load_gold(path)
    sim = sc_dense.SCDense(sc_dense.d2q9(), g["is_domain"], model=str(g["model"]), relax=p["relax"],
                           tau=(float(p["tau0"]), float(p["tau1"])), G=float(p["G"]), Gs=(float(p["Gs0"]), float(p["Gs1"])),
                           inlet=p["inlet"], outlet=p["outlet"], v_in=(float(p["vy0"]), float(p["vy1"])))
    reg = g["region0"]
    sim.set_densities(np.stack([np.where(reg, float(p["rho0"]), float(p["bg0"])),
                                np.where(reg, float(p["bg1"]), float(p["rho1"]))]))
    for s in range(g["rho"].shape[0]):
        sim.step(1)
        np.testing.assert_allclose(sim.rho[:, 0], g["rho"][s], rtol=0, atol=5e-13, err_msg="rho snapshot %d" % s)
        np.testing.assert_allclose(sim.uph[0, 0], g["ux"][s], rtol=0, atol=5e-13, err_msg="ux snapshot %d" % s)
        np.testing.assert_allclose(sim.uph[1, 0], g["uy"][s], rtol=0, atol=5e-13, err_msg="uy snapshot %d" % s)


@pytest.mark.parametrize("model,relax", [("ShanChen", "SRT"), ("EFS", "SRT"), ("EFS", "MRT")])
def test_d3q19_mass_and_z_invariance(model, relax):
    rng = np.random.default_rng(4)
    nz, ny, nx = 4, 10, 12
    dom = np.ones((nz, ny, nx), bool); dom[:, 4:6, 5:8] = False
    base = 0.5 + 0.3 * rng.random((ny, nx))
    rho = np.stack([np.broadcast_to(base, (nz, ny, nx)), np.broadcast_to(1.1 - base, (nz, ny, nx))]) * dom
    G = 3.0 if model == "ShanChen" else 0.2
    sim = sc_dense.SCDense(sc_dense.d3q19(), dom, model=model, relax=relax, tau=(1.0, 0.9), G=G)
    sim.set_densities(rho)
    m0 = sim.rho.sum(axis=(1, 2, 3))
    sim.step(6)
    assert np.allclose(sim.rho.sum(axis=(1, 2, 3)), m0, rtol=1e-12)
    assert np.abs(sim.rho - sim.rho[:, 0:1]).max() < 1e-13        # a z-invariant state stays z-invariant
    assert np.isfinite(sim.rho).all()
