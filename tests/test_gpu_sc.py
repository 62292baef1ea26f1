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


# ---- D3Q19 Shan-Chen (`ShanChenD3Q19`): the same operators with 19 velocities vs the lattice-generic oracle ----
@pytest.mark.parametrize("model,relax", [("ShanChen", "SRT"), ("EFS", "SRT"), ("EFS", "MRT")])
@pytest.mark.parametrize("solid", [False, True])
def test_d3q19_vs_dense_oracle(model, relax, solid):
    m, m_ref = cases.case_sc_d3q19(None, model, relax, solid=solid)
    assert abs(m - m_ref).max() < 1e-9


def test_d3q19_larger_box_graph_replay():
    """24 x 20 x 36 nodes, 30 steps in chunks of 1 / 2 / 27 (the last one replays the captured CUDA graph)"""
    m, m_ref = cases.case_sc_d3q19(None, "EFS", "MRT", n=(24, 20, 36), steps=30)
    assert abs(m - m_ref).max() < 1e-8


@pytest.mark.parametrize("model,relax,outlet", [("ShanChen", "SRT", "Convective"), ("EFS", "MRT", "Dirichlet"), ("EFS", "SRT", "Convective")])
def test_d3q19_open_boundaries_vs_dense_oracle(model, relax, outlet):
    cases.case_sc_d3q19_open(None, model, relax, outlet)
    cases.case_sc_d3q19_open(None, model, relax, outlet, n=(30, 12, 16), steps=14)
