"""CPU tier for the two-pass form of the Shan-Chen loops (csrc/sc_fast.cuh): on the host build (no FMA contraction) it must be
BIT-EQUAL to the reference-ordered operators (LBM_FLAG_GENERIC_KERNELS), which are pinned to the reference's golden vectors --
closed and open boxes, solids across the periodic faces, D2Q9 and D3Q19, odd and even numbers of fused iterations -- and it
must follow the golden trajectories itself when a call spans many iterations."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
import cases
from openlbmpm_b200 import _lib
from test_hostcheck_slabs import geometry, run


@pytest.fixture(scope="module")
def lib():
    return hostcheck_build.build()


SC = dict(model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, sc_G=[0, 0.9, 0, 0, 0.9, 0])
EFS_SRT = dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, sc_G=[0, 0.15, 0, 0, 0.15, 0])
EFS_MRT = dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, sc_G=[0, 0.15, 0, 0, 0.15, 0])
IN = dict(inlet=_lib.INLET_VELOCITY, sc_inlet_velocity=[0.0, -5.0e-4], sc_rho_out=[1.0, 0.02])
VARIANTS = [
    ("original Shan-Chen", SC),
    ("explicit forcing SRT", EFS_SRT),
    ("explicit forcing MRT", EFS_MRT),
    ("original Shan-Chen, velocity inlet + convective outlet", dict(SC, outlet=_lib.OUTLET_CONVECTIVE, **IN)),
    ("original Shan-Chen, velocity inlet only", dict(SC, **IN)),
    ("explicit forcing MRT, velocity inlet + pressure outlet", dict(EFS_MRT, outlet=_lib.OUTLET_PRESSURE, **IN)),
    ("explicit forcing SRT, pressure outlet only", dict(EFS_SRT, outlet=_lib.OUTLET_PRESSURE, sc_rho_out=[1.0, 0.02])),
]


@pytest.mark.parametrize("lattice,shape", [(9, (24, 12)), (19, (12, 6, 8))])
@pytest.mark.parametrize("solid", [False, True])
@pytest.mark.parametrize("name,kw", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_two_pass_form_bit_equal_to_reference_ordered_operators(lattice, shape, solid, name, kw, lib):
    open_ends = kw.get("inlet", 0) != 0 or kw.get("outlet", 0) != 0
    dom = geometry(shape, solid, open_ends)
    rng = np.random.default_rng(5)
    base = 0.5 + 0.3 * (rng.random(shape) - 0.5)
    rho = [base, 1.0 - base]
    par = dict(n_components=2, sc_tau=[1.0, 0.9], sc_Gsolid=[-0.1, 0.1])
    par.update(kw)
    steps = [1, 3, 4, 2, 9]          # 3: one fused iteration; 4: two; 9: seven (odd: the buffers end up swapped)
    ref = run(lib, lattice, dom, rho, steps, 1, flags=_lib.FLAG_GENERIC_KERNELS, **par)
    got = run(lib, lattice, dom, rho, steps, 1, **par)
    for k in range(len(steps)):
        assert np.array_equal(got[0][k], ref[0][k]), "macroscopic fields differ after chunk %d (max %.3e)" % (
            k, np.abs(got[0][k] - ref[0][k]).max())
    assert np.array_equal(got[1], ref[1]), "populations differ"


def test_two_pass_form_is_what_runs(lib):
    """launch count of a 10-iteration call on a closed box: 1 reference-ordered iteration (2 launches) + first collision (1) +
    8 fused iterations (2 each) + tail (1) for the original model"""
    shape = (16, 12)
    dom = np.ones(shape, bool)
    e = _lib.Engine(9, shape, lib_path=lib, n_components=2, sc_tau=[1.0, 0.9], sc_Gsolid=[0.0, 0.0], **SC)
    e.set_geometry(dom)
    e.init_equilibrium(np.full(shape, 0.6), np.full(shape, 0.4))
    e.step(1)
    e.step(10)
    assert e.timing()["launches"] == 2 + 1 + 8 * 2 + 1
    e.close()


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
@pytest.mark.parametrize("chunk", [7, 20])
def test_trajectory_vs_reference_in_long_calls(path, chunk, lib):
    cases.check_sc_vs_gold(path, lib, chunk=chunk)
