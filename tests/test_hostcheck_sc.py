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
