"""Pins oracle/sc2d.py against golden vectors produced by the reference's own Shan-Chen drivers
(tests/golden/gen_goldens_sc2d.py, reference commit 3d84189)."""
import numpy as np
import pytest

import cases
from oracle import sc2d


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_trajectory_matches_reference(path):
    g, p = cases.load_gold(path)
    dom = g["is_domain"]
    sim = sc2d.SC2D(dom, model=str(g["model"]), relax=p["relax"], tau=(float(p["tau0"]), float(p["tau1"])),
                    G=float(p["G"]), Gs=(float(p["Gs0"]), float(p["Gs1"])), inlet=p["inlet"], outlet=p["outlet"],
                    vy=(float(p["vy0"]), float(p["vy1"])), scheme=int(p.get("scheme", 4)))
    reg = g["region0"]
    sim.set_densities(np.stack([np.where(reg, float(p["rho0"]), float(p["bg0"])),
                                np.where(reg, float(p["bg1"]), float(p["rho1"]))]))
    for s in range(g["rho"].shape[0]):
        sim.step(1)
        np.testing.assert_allclose(sim.rho, g["rho"][s], rtol=0, atol=2e-13, err_msg="rho snapshot %d" % s)
        np.testing.assert_allclose(sim.uph[0], g["ux"][s], rtol=0, atol=2e-13, err_msg="ux snapshot %d" % s)
        np.testing.assert_allclose(sim.uph[1], g["uy"][s], rtol=0, atol=2e-13, err_msg="uy snapshot %d" % s)
        if str(g["model"]) == "ShanChen":      # EFS: the tap sits before the force of this iteration is evaluated
            np.testing.assert_allclose(sim.F[:, 0], g["fx"][s], rtol=0, atol=2e-13, err_msg="Fx snapshot %d" % s)
    np.testing.assert_allclose(sim.f, g["pdf_last"], rtol=0, atol=2e-13)
