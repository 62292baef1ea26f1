"""Size-independent properties on RANDOM masks (CPU tier, host test hook): every closed-box model conserves the mass of
each component to rounding on arbitrary solid layouts (half-way bounce back), stays finite, and the factored fast path
agrees with the reference-ordered kernels on the same ragged geometry."""
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
import cases
from openlbmpm_b200 import _lib

LIB = None


def lib():
    global LIB
    LIB = LIB or hostcheck_build.build()
    return LIB


def random_case(seed, lattice, porosity):
    rng = np.random.default_rng(seed)
    shape = tuple(int(rng.integers(3, 9)) for _ in range(3 if lattice == 19 else 2))
    dom = rng.random(shape) < porosity
    if not dom.any():
        dom.flat[0] = True
    r = (0.2 + 0.6 * rng.random(shape)) * dom
    return shape, dom, r, (1.0 - r) * dom


MODELS = ["cg_fast", "cg_general", "cg_srt", "cgp", "sc", "efs_srt", "efs_mrt"]


def make_engine(model, lattice, shape, theta):
    L = lib()
    if model.startswith("cg_"):
        return _lib.Engine(lattice, shape, lib_path=L, relax=_lib.RELAX_SRT if model == "cg_srt" else _lib.RELAX_MRT,
                           flags=1 if model != "cg_fast" else 0, sigma=0.05, contact_angle_deg=theta, tauR=1.0, tauB=0.8)
    if model == "cgp":
        return _lib.Engine(lattice, shape, lib_path=L, surface_tension_type=_lib.ST_PERTURBATION, AkR=5e-3, AkB=8e-3,
                           solid_phi=0.4, tauR=1.0, tauB=0.8)
    return _lib.Engine(lattice, shape, lib_path=L, model=_lib.MODEL_SC if model == "sc" else _lib.MODEL_EFS,
                       relax=_lib.RELAX_MRT if model == "efs_mrt" else _lib.RELAX_SRT, n_components=2, sc_tau=[1.0, 0.9],
                       sc_G=[0, 0.9 if model == "sc" else 0.15, 0, 0, 0.9 if model == "sc" else 0.15, 0], sc_Gsolid=[-0.1, 0.1])


@settings(max_examples=80, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10 ** 6), lattice=st.sampled_from([9, 19]), porosity=st.sampled_from([0.35, 0.6, 0.85, 1.0]),
       model=st.sampled_from(MODELS), theta=st.sampled_from([35.0, 90.0, 140.0]))
def test_mass_is_conserved_on_random_masks(seed, lattice, porosity, model, theta):
    shape, dom, rR, rB = random_case(seed, lattice, porosity)
    eng = make_engine(model, lattice, shape, theta)
    eng.set_geometry(dom)
    eng.init_equilibrium(rR, rB)
    m0 = eng.total_mass()
    eng.step(6)
    rho, u = eng.download_macros()
    m1 = eng.total_mass()
    eng.close()
    assert np.isfinite(rho[0]).all() and np.isfinite(rho[1]).all() and all(np.isfinite(a).all() for a in u)
    np.testing.assert_allclose(m1, m0, rtol=1e-11, atol=1e-12)
    assert not rho[0][~dom].any() and not rho[1][~dom].any()          # nothing leaks into the solid


@settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10 ** 6), lattice=st.sampled_from([9, 19]), porosity=st.sampled_from([0.5, 0.8, 1.0]))
def test_fast_path_equals_reference_ordered_kernels_on_random_masks(seed, lattice, porosity):
    shape, dom, rR, rB = random_case(seed, lattice, porosity)
    out = []
    for model in ("cg_fast", "cg_general"):
        eng = make_engine(model, lattice, shape, 70.0)
        eng.set_geometry(dom); eng.init_equilibrium(rR, rB); eng.step(5)
        rho, u = eng.download_macros()
        out.append(np.stack(rho + u))
        eng.close()
    # 1e-16 agreement is the rule; on ragged masks a near-solid node occasionally has its colour gradient almost
    # parallel to the solid normal, where the Akai correction amplifies rounding by 1 / sin(theta') (seed 200, D2Q9,
    # porosity 0.5: one jump to 2e-10 at step 5 that decays afterwards) -- the bound is that of the ill-conditioned case
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=1e-8)


@settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10 ** 6), schemes=st.sampled_from([9, 5]), porosity=st.sampled_from([0.5, 0.8, 1.0]), react=st.booleans())
def test_tracers_are_conserved_on_random_masks(seed, schemes, porosity, react):
    """closed box, any solid layout: collision, interface term and half-way bounce back conserve every tracer; with the
    reaction A + B -> C of the 5-velocity branch, A + C and B + C are conserved instead"""
    rng = np.random.default_rng(seed)
    shape = (int(rng.integers(4, 10)), int(rng.integers(4, 10)))
    dom = rng.random(shape) < porosity
    if dom.sum() < 2:
        dom[:] = True
    r = (0.2 + 0.6 * rng.random(shape)) * dom
    react = react and schemes == 5
    nt = 3 if react else 2
    conc = [(0.1 + rng.random(shape)) * dom for _ in range(nt)]
    eng = _lib.Engine(9, shape, lib_path=lib(), sigma=0.05, contact_angle_deg=70.0, tauR=1.0, tauB=0.8)
    eng.tracer_setup(n_tracers=nt, relax=_lib.RELAX_MRT, dxx=(0.05, 0.1, 0.07), dyy=(0.08, 0.1, 0.07), dxy=(0.01, 0.0, 0.0),
                     dyx=(0.02, 0.0, 0.0), beta=(0.6, 0.3, 0.1), n_schemes=schemes, reaction=react, reaction_rate=0.05,
                     diff_j=(0.3, 1. / 3., 0.4))
    eng.set_geometry(dom)
    eng.init_equilibrium(r, (1.0 - r) * dom)
    eng.tracer_init(*conc)
    m0 = np.array([c.sum() for c in eng.tracer_download()])
    eng.step(7)
    out = eng.tracer_download()
    eng.close()
    m1 = np.array([c.sum() for c in out])
    assert all(np.isfinite(c).all() and not c[~dom].any() for c in out)
    if react:
        np.testing.assert_allclose([m1[0] + m1[2], m1[1] + m1[2]], [m0[0] + m0[2], m0[1] + m0[2]], rtol=1e-12)
        assert m1[2] > m0[2]
    else:
        np.testing.assert_allclose(m1, m0, rtol=1e-12)
