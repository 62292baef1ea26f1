"""GPU tier, run last (file name): the opt-in persistent form of the one-thread-per-node fast path (LBM_FLAG_PERSISTENT,
cooperative launch with grid-wide barriers).  Added after the round-1 GPU budget was spent: verified so far on host threads
(tests/test_hostcheck_tiled.py); this is its first run on the hardware."""
import pytest

import cases

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]      # a cooperative kernel that cannot make progress must not hang the tier


def test_persistent_kernel_equals_launched_form():
    cases.check_persistent_kernel(None)


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_shan_chen_trajectories_persistent_kernel(path):
    from openlbmpm_b200 import _lib
    cases.check_sc_vs_gold(path, None, chunk=13, flags=_lib.FLAG_PERSISTENT)


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_colour_gradient_trajectories_persistent_kernel(path):
    from openlbmpm_b200 import _lib
    cases.check_trajectory_vs_gold(path, None, chunk=19, flags=_lib.FLAG_PERSISTENT)
