"""oracle/tr_dense.py (solute tracers riding on the colour-gradient CSF flow): its D2Q9
instantiation must reproduce the golden vectors produced by the reference's own kernels
(tests/golden/gen_goldens_tr2d.py); D3Q19 must conserve the tracer and keep a uniform concentration uniform."""
import numpy as np
import pytest

import cases
from oracle import cg_dense, tr_dense


@pytest.mark.parametrize("path", cases.GOLD_TR2D, ids=[cases.os.path.basename(p)[5:-4] for p in cases.GOLD_TR2D])
def test_d2q9_matches_reference_kernels(path):
    g, p = cases.load_gold(path)
    dom, red, minor = g["is_domain"], g["red_mask"], float(g["minor"])
    sim = cg_dense.CGDense(cg_dense.d2q9(), dom, sigma=float(p["sigma"]), theta_deg=float(p["theta"]), wetting=int(p["wetting"]),
                           beta=float(p["beta"]), delta=float(p["delta"]), tauR=float(p["tauR"]), tauB=float(p["tauB"]),
                           tautype=int(p["tautype"]), relax="MRT")
    sim.set_densities(*cases.gold_initial_densities(g, p))
    nt = int(p.get("nt", 1))
    per = lambda key: (float(p[key]),) * nt
    if int(p.get("schemes", 9)) == 5:
        tr = tr_dense.TracerDenseQ5(sim, dxx=per("dxx"), dyy=per("dyy"), dxy=per("dxy"), dyx=per("dyx"), beta=per("beta_tr"),
                                    reaction_rate=float(p["rate"]) if p["reaction"] == "yes" else None, diff_j=per("diffj"),
                                    inlet_conc=per("conc_in") if p["tr_inlet"] == "Dirichlet" else None,
                                    freeflow_outlet=p["tr_outlet"] == "Freeflow")
    else:
        tr = tr_dense.TracerDense(sim, relax=p["tr_relax"], tau=per("tr_tau"), dxx=per("dxx"), dyy=per("dyy"), dxy=per("dxy"),
                                  dyx=per("dyx"), beta=per("beta_tr"))
    tr.set_concentrations(g["tracer0"])
    for s in range(g["rhoR"].shape[0]):
        tr.step(1)          # iteration s: flow head, gradient, tracer phase, flow collision
        for i in range(nt):
            np.testing.assert_allclose(tr.conc[i, 0], g["conc"][s][i], rtol=0, atol=5e-13, err_msg="tracer %d snapshot %d" % (i, s))
    assert len(cases.GOLD_TR2D) >= 2


def test_flow_of_the_transport_loop_matches_reference_kernels():
    """the flow side of the same golden files: densities and velocity at the flow output point of every iteration"""
    g, p = cases.load_gold(cases.GOLD_TR2D[0])
    dom, red, minor = g["is_domain"], g["red_mask"], float(g["minor"])
    sim = cg_dense.CGDense(cg_dense.d2q9(), dom, sigma=float(p["sigma"]), theta_deg=float(p["theta"]), wetting=int(p["wetting"]),
                           beta=float(p["beta"]), delta=float(p["delta"]), tauR=float(p["tauR"]), tauB=float(p["tauB"]),
                           tautype=int(p["tautype"]), relax="MRT")
    sim.set_densities(np.where(red, float(p["rhoR"]), minor), np.where(red, minor, float(p["rhoB"])))
    sim.stream_only()
    for s in range(g["rhoR"].shape[0]):
        sim.head()
        np.testing.assert_allclose(sim.rhoR[0], g["rhoR"][s], rtol=0, atol=5e-13, err_msg="rhoR snapshot %d" % s)
        np.testing.assert_allclose(sim.u[1, 0], g["uy"][s], rtol=0, atol=5e-13, err_msg="uy snapshot %d" % s)
        sim.body()


def test_d3q19_tracer_conservation_and_uniform_state():
    n = (8, 10, 12)
    rng = np.random.default_rng(2)
    dom = cases.sphere_geometry(n, 2.4)
    rhoR = 0.5 + 0.4 * (rng.random(n) - 0.5)
    sim = cg_dense.CGDense(cg_dense.d3q19(), dom, relax="MRT")
    sim.set_densities(rhoR, 1.0 - rhoR)
    tr = tr_dense.TracerDense(sim, relax="SRT", tau=(0.9, 0.7), beta=(0.5, 0.0))
    conc = np.stack([0.3 + rng.random(n), np.full(n, 0.7)])
    tr.set_concentrations(conc)
    m0 = tr.conc.sum(axis=(1, 2, 3))
    tr.step(8)
    np.testing.assert_allclose(tr.conc.sum(axis=(1, 2, 3)), m0, rtol=1e-12)
    # beta = 0: a uniform concentration is only moved by the (weakly compressible) flow: stays within 1e-2 of uniform
    assert np.abs(tr.conc[1][dom] - 0.7).max() < 2e-2
