"""The tiled sm_100a kernels of cg_fast.cu (column-marching CTAs, shared-memory phi / normal / scalar windows, pull masks) on
the CPU tier: the host test hook runs them with one host thread per CUDA thread of a CTA and a pthread barrier for
__syncthreads (csrc/cta_emu.h).  Same cases and oracle as the GPU tier (tests/test_gpu_cg.py); asynchronous copies are
immediate and the TMA variants are not instantiated, so this checks the kernels' logic and barrier protocol (removing one
__syncthreads makes these tests fail), not their speed."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
import cases
from openlbmpm_b200 import _lib

pytestmark = pytest.mark.timeout(600)


@pytest.fixture(scope="module")
def lib():
    return hostcheck_build.build()


def launches_per_step(lib, shape, dom=None, **kw):
    eng = _lib.Engine(19, shape, lib_path=lib, **kw)
    eng.set_geometry(np.ones(shape, bool) if dom is None else dom)
    r = np.random.default_rng(0).random(shape) * 0.2 + 0.4
    eng.init_equilibrium(r, 1 - r)
    eng.step(2); eng.step(4)
    n = eng.timing()["launches"] / 4
    eng.close()
    return n


def test_the_tiled_kernels_are_the_ones_that_run(lib):
    assert launches_per_step(lib, (10, 16, 32)) == 4                 # ghost wrap, tiled density, ghost wrap, tiled collision
    assert launches_per_step(lib, (10, 16, 32), flags=2) == 5        # ... untiled: + the gradient operator
    assert launches_per_step(lib, (10, 12, 30)) == 5                 # extents that are no multiple of the tile: untiled


@pytest.mark.parametrize("relax", ["MRT", "SRT"])
def test_tiled_periodic_vs_oracle(relax, lib):
    cases.case_d3q19_periodic(lib, n=(10, 16, 32), steps=8, relax=relax)


def test_tiled_two_tiles_and_chunks_vs_oracle(lib):
    cases.case_d3q19_periodic(lib, n=(70, 16, 64), steps=4)          # 2 tiles in x, 3 z-chunks (32 + 32 + 6 planes)


def test_tiled_sphere_wetting_vs_oracle(lib):
    cases.case_d3q19_sphere(lib, n=(12, 16, 32), steps=8)             # <solids> variants: pull masks, wetting in the normal tile
    cases.case_d3q19_sphere(lib, n=(12, 16, 32), steps=6, centre=[0.3, 0.2, 0.4])     # solid across the periodic faces / tile edges


def test_tiled_density_or_collision_alone(lib):
    cases.case_d3q19_sphere(lib, n=(12, 16, 32), steps=6, flags=4)    # tiled collision, untiled density


@pytest.mark.parametrize("inlet,outlet", [("Neumann", "Convective"), ("Dirichlet", "Dirichlet")])
def test_tiled_open_boundaries_vs_oracle(inlet, outlet, lib):
    cases.case_d3q19_open_boundaries(lib, inlet, outlet, n=(40, 16, 32), steps=8)


def test_tiled_slabs_bit_equal(lib):
    """P = 2, 3 slabs over the in-process ring with the tiled kernels on every slab"""
    import test_hostcheck_slabs as S
    S.compare(lib, 19, (24, 8, 32), [1, 2, 4], contact_angle_deg=70.0)
    S.compare(lib, 19, (24, 8, 32), [2, 4], solid=False)
    S.compare(lib, 19, (48, 8, 32), [1, 3], worlds=(2, 3), **dict(S.OPEN, contact_angle_deg=60.0))
