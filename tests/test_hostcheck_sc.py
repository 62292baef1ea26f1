"""CPU tier for the Shan-Chen models: operator code built for the host (tests/hostcheck) vs the reference's
golden vectors (tests/golden/sc2d_*.npz, produced by gen_goldens_sc2d.py from the reference drivers)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
import cases


@pytest.fixture(scope="module")
def lib():
    return hostcheck_build.build()


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_trajectory_vs_reference(path, lib):
    cases.check_sc_vs_gold(path, lib)


def test_chunked(lib):
    cases.check_sc_vs_gold(cases.GOLD_SC2D[0], lib, chunk=13)


# ---- D3Q19 (`ShanChenD3Q19`: named by the reference's main.py, absent from its tree): same operators, 19 velocities,
# checked against the lattice-generic oracle whose D2Q9 instantiation is pinned to the reference's vectors ----
@pytest.mark.parametrize("model,relax", [("ShanChen", "SRT"), ("EFS", "SRT"), ("EFS", "MRT")])
@pytest.mark.parametrize("solid", [False, True])
def test_d3q19_vs_dense_oracle(model, relax, solid, lib):
    m, m_ref = cases.case_sc_d3q19(lib, model, relax, solid=solid)
    assert abs(m - m_ref).max() < 1e-9


def test_d2q9_dense_case_through_the_generic_operators(lib):
    import numpy as np
    rng = np.random.default_rng(3)
    dom = np.ones((20, 24), bool); dom[8:11, 5:15] = False
    r0 = 0.6 + 0.3 * (rng.random(dom.shape) - 0.5)
    for model, relax in (("ShanChen", "SRT"), ("EFS", "MRT")):
        cases.run_sc_dense_case(9, dom, [r0, 1.1 - r0], 10, lib, model=model, relax=relax, G=3.0 if model == "ShanChen" else 0.2)


@pytest.mark.parametrize("model,relax,outlet", [("ShanChen", "SRT", "Convective"), ("EFS", "MRT", "Dirichlet"), ("EFS", "SRT", "Convective"),
                                                ("EFS", "SRT", "Dirichlet")])
def test_d3q19_open_boundaries_vs_dense_oracle(model, relax, outlet, lib):
    cases.case_sc_d3q19_open(lib, model, relax, outlet)


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_trajectory_vs_reference_persistent_kernel(path, lib):
    """LBM_FLAG_PERSISTENT: the iterations of a call inside one cooperative kernel (host threads here, csrc/cta_emu.h)"""
    from openlbmpm_b200 import _lib
    cases.check_sc_vs_gold(path, lib, chunk=13, flags=_lib.FLAG_PERSISTENT)
