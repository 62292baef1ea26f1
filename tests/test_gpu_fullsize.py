"""GPU tier at the sizes of BASELINE.json's configs: comparisons with the oracle where it finishes in seconds,
size-independent properties (mass balance, symmetry, finiteness) elsewhere."""
import numpy as np
import pytest

from openlbmpm_b200 import _lib
from oracle import sc2d

pytestmark = pytest.mark.gpu


def discs(shape, seed=7, porosity=0.6, rmin=6, rmax=14, buffer_rows=40):
    """cfg 3 generator: union of discs until the target porosity is reached; void buffer rows"""
    rng = np.random.default_rng(seed)
    ny, nx = shape
    dom = np.ones(shape, bool)
    core = slice(buffer_rows, ny - buffer_rows)
    k = 0
    while True:
        if k % 25 == 0 and dom[core].mean() <= porosity:
            return dom
        k += 1
        r = rng.uniform(rmin, rmax); cx = rng.uniform(0, nx); cy = rng.uniform(buffer_rows + r, ny - buffer_rows - r)
        y0, y1 = max(0, int(cy - r) - 1), min(ny, int(cy + r) + 2)
        x0, x1 = max(0, int(cx - r) - 1), min(nx, int(cx + r) + 2)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dom[y0:y1, x0:x1] &= ((xx - cx) ** 2 + (yy - cy) ** 2) > r * r


def test_cfg1_shanchen_128_droplet_vs_oracle():
    """D2Q9 original Shan-Chen, 128 x 128 periodic droplet (IniFiles/shanchen2D.ini values)"""
    n = (128, 128)
    yy, xx = np.mgrid[0:128, 0:128]
    reg = (xx - 64) ** 2 + (yy - 64) ** 2 <= 20 ** 2
    rho = np.stack([np.where(reg, 1.0, 0.06), np.where(reg, 0.06, 1.0)])
    sim = sc2d.SC2D(np.ones(n, bool), model="ShanChen", tau=(1.0, 1.0), G=3.8, Gs=(-0.4, 0.4))
    sim.set_densities(rho)
    eng = _lib.Engine(9, n, model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, n_components=2, sc_tau=[1.0, 1.0],
                      sc_G=[0, 3.8, 0, 0, 3.8, 0], sc_Gsolid=[-0.4, 0.4])
    eng.set_geometry(np.ones(n, bool))
    eng.init_equilibrium(rho[0], rho[1])
    m0 = eng.total_mass()
    for _ in range(2):
        eng.step(50); sim.step(50)
        r, u = eng.download_macros()
        np.testing.assert_allclose(r[0], sim.rho[0], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r[1], sim.rho[1], rtol=0, atol=1e-8)
        np.testing.assert_allclose(u[1], sim.uph[1], rtol=0, atol=1e-8)
    # the rest of the 1000 steps: the segregating droplet amplifies rounding differences exponentially (1.6e-7 after
    # 250 steps, 3e-3 after 500 against the NumPy oracle), so only invariants are checked from here on
    eng.step(900)
    r, u = eng.download_macros()
    assert np.isfinite(r[0]).all() and np.isfinite(u[0]).all()
    assert np.allclose(eng.total_mass(), m0, rtol=1e-12)
    assert r[0][64, 64] > 0.5 and r[0][0, 0] < 0.5          # the droplet of fluid 0 is still there
    eng.close()


def test_cfg2_colour_gradient_512_capillary_intrusion_properties():
    """D2Q9 CSF-MRT 512 x 512, side walls, contact angle 60, velocity inlet, pressure outlet (RKtwophasesetup2D.ini)"""
    ny = nx = 512
    dom = np.ones((ny, nx), bool)
    dom[10:-10, 0] = False; dom[10:-10, -1] = False
    red = np.indices((ny, nx))[0] >= ny - 20
    v = -1.0e-4
    eng = _lib.Engine(9, (ny, nx), relax=_lib.RELAX_MRT, sigma=0.1, contact_angle_deg=60.0, wetting_type=2, beta=0.7,
                      delta=0.98, tauR=1.0, tauB=1.0, tau_type=2, inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE,
                      inlet_velocity=v, rhoBL=1.0, rhoRL=5e-8)
    eng.set_geometry(dom)
    eng.init_equilibrium(np.where(red, 1.0, 5e-8) * dom, np.where(red, 5e-8, 1.0) * dom)
    m0 = eng.total_mass()
    steps = 400
    eng.step(steps)
    rho, u = eng.download_macros()
    m1 = eng.total_mass()
    assert all(np.isfinite(a).all() for a in rho + u)
    # mirror symmetry about the vertical centre line (geometry and initial condition are symmetric)
    assert np.abs(rho[0] - rho[0][:, ::-1]).max() < 1e-7
    # red enters through the inlet at |v| rho per node and step (inlet row is 512 nodes wide)
    gained = m1[0] - m0[0]
    assert 0.7 * abs(v) * nx * steps < gained < 1.3 * abs(v) * nx * steps, gained
    eng.close()


def test_cfg3_explicit_forcing_1024_porous_vs_oracle_crop_and_properties():
    """D2Q9 EFS-MRT with half-way bounce back in a disc pack: 256 x 256 against the oracle, 1024 x 1024 properties"""
    par = dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, n_components=2, sc_tau=[1.0, 1.0], sc_G=[0, 0.2, 0, 0, 0.2, 0],
               sc_Gsolid=[-0.14, 0.14], inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE,
               sc_inlet_velocity=[0.0, -5.03e-4], sc_rho_out=[1.0, 0.02])
    for shape, steps, check in (((256, 256), 60, True), ((1024, 1024), 200, False)):
        dom = discs(shape, buffer_rows=40 if shape[0] > 300 else 24)
        reg = np.indices(shape)[0] < shape[0] - 10
        rho = np.stack([np.where(reg, 1.0, 0.02), np.where(reg, 0.02, 1.0)]) * dom
        eng = _lib.Engine(9, shape, **par)
        eng.set_geometry(dom)
        eng.init_equilibrium(rho[0], rho[1])
        eng.step(steps)
        r, u = eng.download_macros()
        assert all(np.isfinite(a).all() for a in r + u)
        assert (r[0][~dom] == 0).all()
        if check:
            sim = sc2d.SC2D(dom, model="EFS", relax="MRT", tau=(1.0, 1.0), G=0.2, Gs=(-0.14, 0.14), inlet="Neumann",
                            outlet="Dirichlet", vy=(0.0, -5.03e-4))
            sim.set_densities(rho)
            sim.step(steps)
            np.testing.assert_allclose(r[0], sim.rho[0], rtol=0, atol=1e-9)
            np.testing.assert_allclose(r[1], sim.rho[1], rtol=0, atol=1e-9)
            np.testing.assert_allclose(u[1], sim.uph[1], rtol=0, atol=1e-9)
        eng.close()


def test_cfg4_d3q19_256_spinodal_mass_and_determinism():
    """D3Q19 CSF-MRT 256^3 periodic spinodal start: colour masses conserved, two runs bit-identical"""
    n = (256, 256, 256)
    out = []
    for _ in range(2):
        eng = _lib.Engine(19, n, relax=_lib.RELAX_MRT, sigma=0.1, beta=0.7)
        eng.set_geometry(np.ones(n, np.uint8))
        eng.init_spinodal_device(0.01, 20260117)
        m0 = eng.total_mass()
        eng.step(100)
        m1 = eng.total_mass()
        assert np.allclose(m0, m1, rtol=1e-12), (m0, m1)
        rho, _ = eng.download_macros()
        out.append(rho[0])
        eng.close()
    assert np.array_equal(out[0], out[1])
    assert np.isfinite(out[0]).all() and out[0].std() > 0.0
