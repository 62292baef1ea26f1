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
    # untiled (flag, or extents that are no multiple of the tile): density, gradient, collision -- a single slab wraps the
    # flow axis by index arithmetic, the two ghost-plane copies only come back with LBM_FLAG_GHOST_PLANES
    assert launches_per_step(lib, (10, 16, 32), flags=2) == 3
    assert launches_per_step(lib, (10, 12, 30)) == 3
    assert launches_per_step(lib, (10, 12, 30), flags=_lib.FLAG_GHOST_PLANES) == 5


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


def test_index_wrap_equals_ghost_plane_copies(lib):
    """one slab: wrapping the flow axis by index arithmetic (default when the tiled kernels are not in play) is bit-equal
    to copying the periodic ghost planes every step, for every model"""
    rng = np.random.default_rng(6)
    for lattice, shape in ((9, (14, 18)), (19, (9, 6, 10))):
        dom = np.ones(shape, bool); dom[(slice(4, 6),) + (slice(2, 5),) * (len(shape) - 1)] = False
        dom[(0,) + (slice(0, 2),) * (len(shape) - 1)] = False                    # a solid on the periodic seam
        r = 0.5 + 0.3 * (rng.random(shape) - 0.5)
        for kw in (dict(), dict(flags=1), dict(surface_tension_type=_lib.ST_PERTURBATION, AkR=8e-3, AkB=1e-2, solid_phi=0.3),
                   dict(model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, n_components=2, sc_tau=[1.0, 0.9], sc_G=[0, 0.9, 0, 0, 0.9, 0], sc_Gsolid=[-0.1, 0.1]),
                   dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, n_components=2, sc_tau=[1.0, 0.9], sc_G=[0, 0.15, 0, 0, 0.15, 0], sc_Gsolid=[-0.1, 0.1],
                        sc_isotropy=10 if lattice == 9 else 4)):
            out = []
            for extra in (0, _lib.FLAG_GHOST_PLANES):
                par = dict(kw); par["flags"] = par.get("flags", 0) | extra
                eng = _lib.Engine(lattice, shape, lib_path=lib, **par)
                eng.set_geometry(dom); eng.init_equilibrium(r * dom, (1 - r) * dom); eng.step(7)
                rho, u = eng.download_macros()
                out.append(np.stack(rho + u))
                eng.close()
            assert np.array_equal(out[0], out[1]), (lattice, kw)


def test_persistent_kernel_equals_launched_form(lib):
    """LBM_FLAG_PERSISTENT: all steps of an lbm_step call in one cooperative kernel with grid-wide barriers between the
    phases (here: 3 CTAs x 32 host threads, csrc/cta_emu.h) -- bit-equal to one launch per phase"""
    cases.check_persistent_kernel(lib)


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_colour_gradient_trajectories_persistent_kernel(path, lib):
    """the reference's own vectors through the persistent kernel (19 steps per call)"""
    cases.check_trajectory_vs_gold(path, lib, chunk=19, flags=_lib.FLAG_PERSISTENT)


def test_d2q9_tile_kernels_equal_the_operators_they_replace(lib):
    """the D2Q9 tile kernels on host threads (cta_emu.h): bit-equal to the one-thread-per-node fast path, equal to the oracle"""
    cases.check_d2q9_tile_kernels(lib, only=("WettingType 1", "velocity inlet", "convective"))     # host threads are slow: 3 of the 5 cases


def test_perturbation_tiled_kernels_on_slabs(lib):
    """the perturbation-operator model with the tiled density pass and cgp_collide_tiled_d3q19 on every slab: P = 2, 3 bit-equal to
    one slab, closed box with solids and the open channel of the reference's 3-D ini"""
    import test_hostcheck_slabs as S
    pert = dict(surface_tension_type=_lib.ST_PERTURBATION, AkR=7e-3, AkB=9e-3, solid_phi=0.7, beta=1.0)
    S.compare(lib, 19, (24, 8, 32), [1, 2, 4], **pert)
    S.compare(lib, 19, (48, 8, 32), [1, 3], worlds=(2, 3), relax=_lib.RELAX_SRT,
              inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE, inlet_velocity=-1e-3, rhoBL=1.0, rhoRL=1e-8, **pert)
