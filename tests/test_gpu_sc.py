"""GPU tier for the Shan-Chen models (original SC, explicit forcing SRT/MRT, open boundaries) against the
reference's golden vectors."""
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_trajectory_vs_reference(path):
    cases.check_sc_vs_gold(path, None)


def test_chunked():
    cases.check_sc_vs_gold(cases.GOLD_SC2D[0], None, chunk=13)


@pytest.mark.parametrize("path", cases.GOLD_SC2D, ids=[cases.gold_id(p) for p in cases.GOLD_SC2D])
def test_trajectory_vs_reference_graph_replay(path):
    cases.check_sc_vs_gold(path, None, chunk=20)
