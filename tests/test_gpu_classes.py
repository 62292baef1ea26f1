"""GPU tier of the class surface: the reference-named classes and main.py on liblbmpm.so."""
import os
import sys

import numpy as np
import pytest

import test_host_classes as T

pytestmark = pytest.mark.gpu


@pytest.fixture()
def results(monkeypatch, tmp_path):
    monkeypatch.setenv("LBM_RESULTS_DIR", str(tmp_path / "results"))
    return tmp_path


def test_cg2d_class(results):
    T.test_cg2d_class_runs_with_reference_style_ini(results)


@pytest.mark.parametrize("which", ["sc", "efs"])
def test_shanchen_class(results, which):
    T.test_shanchen_class(results, which)


def test_cg3d_class_and_main(results, monkeypatch):
    T.test_cg3d_class_and_main(results, monkeypatch)


def test_asynchronous_output_equals_blocking_download():
    """lbm_download_macros_async (device snapshot + copy on a second stream into page-locked buffers) while the next steps
    are already queued: same numbers as the blocking lbm_download_macros at the same step"""
    import cases
    cases.check_async_output(None, (40, 48, 64))
