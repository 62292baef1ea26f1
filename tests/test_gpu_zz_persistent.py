"""GPU tier, run last (file name): the opt-in persistent form of the one-thread-per-node fast path (LBM_FLAG_PERSISTENT,
cooperative launch with grid-wide barriers).  Added after the round-1 GPU budget was spent: verified so far on host threads
(tests/test_hostcheck_tiled.py); this is its first run on the hardware."""
import pytest

import cases

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]      # a cooperative kernel that cannot make progress must not hang the tier


def test_persistent_kernel_equals_launched_form():
    cases.check_persistent_kernel(None)
