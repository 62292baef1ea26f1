"""GPU tier: the parity tests proper.  liblbmpm.so (sm_100a kernels) is driven through the C ABI and
compared with the reference's golden vectors (tests/golden, produced by the reference itself) and with
the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_indexing_bit_exact(path):
    cases.check_indexing_vs_gold(path, None)


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_trajectory_vs_reference(path):
    cases.check_trajectory_vs_gold(path, None)


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_trajectory_vs_reference_general_kernels(path):
    cases.check_trajectory_vs_gold(path, None, flags=1)


def test_chunked_steps_equal_single_steps():
    cases.check_trajectory_vs_gold(cases.GOLD_CG2D[0], None, chunk=7)


@pytest.mark.parametrize("relax", ["MRT", "SRT"])
def test_d3q19_periodic_vs_oracle(relax):
    m, mo = cases.case_d3q19_periodic(None, relax=relax)
    assert np.allclose(m, mo, rtol=0, atol=1e-9)


@pytest.mark.parametrize("relax", ["MRT", "SRT"])
def test_d3q19_periodic_vs_oracle_general_kernels(relax):
    cases.case_d3q19_periodic(None, relax=relax, flags=1)


def test_d3q19_sphere_wetting_vs_oracle():
    cases.case_d3q19_sphere(None)


def test_d3q19_sphere_wetting_vs_oracle_general_kernels():
    cases.case_d3q19_sphere(None, flags=1)


def test_d3q19_larger_box_vs_oracle():
    cases.case_d3q19_periodic(None, n=(24, 40, 72), steps=10)


def test_d2q9_obstacles_vs_oracle():
    cases.case_d2q9_random(None, n=(96, 130), steps=30)


def test_no_state_is_an_error():
    from openlbmpm_b200 import _lib
    eng = _lib.Engine(9, (8, 8))
    with pytest.raises(_lib.LbmError):
        eng.step(1)
    eng.set_geometry(np.ones((8, 8), bool))
    with pytest.raises(_lib.LbmError):
        eng.step(1)
    eng.close()
