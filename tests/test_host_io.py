"""Host-side I/O and pre-processing of the reference's classes: structure image ->
geometry (crop, mirror tiling, side walls, buffer rows), result files, drainage-imbibition restarts.  CPU tier: the
classes are pointed at the host test hook instead of liblbmpm.so."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build

from openlbmpm_b200 import _lib, imagegeo, results

REF_INI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ini")


@pytest.fixture()
def hostlib(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", hostcheck_build.build())
    monkeypatch.setenv("LBM_RESULTS_DIR", str(tmp_path / "results"))
    monkeypatch.setenv("LBM_INITIAL_DIR", str(tmp_path / "initial"))
    return tmp_path


def structure_png(path, shape=(30, 22), margin=3):
    """white canvas with a frame of unused margin, a few black grains inside"""
    from PIL import Image
    img = np.full((shape[0] + 2 * margin, shape[1] + 2 * margin), 255, np.uint8)
    core = img[margin:-margin, margin:-margin]
    core[0, 4:9] = 0; core[-1, 10:15] = 0; core[5:9, 0] = 0; core[12:15, -1] = 0      # the bounding box of the solid
    core[8:12, 6:11] = 0; core[18:23, 12:18] = 0; core[24:27, 3:7] = 0
    Image.fromarray(img).save(path)
    return core.copy()


def test_mirror_tiling_equals_the_reference_loop():
    """imagegeo.expand_image_domain vs a literal restatement of ShanChenD2Q9.py:514-540"""
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2, (5, 7)).astype(float) * 255
    for xn, yn in ((1, 1), (2, 1), (3, 2), (4, 3)):
        rx, ry = np.fliplr(a), np.flipud(a)
        ryx = np.fliplr(ry)
        new = None
        for i in range(yn):
            row = a if i % 2 == 0 else ry
            for j in range(1, xn):
                row = np.hstack((row, (rx if i % 2 == 0 else ryx) if j % 2 != 0 else (a if i % 2 == 0 else ry)))
            new = row if i == 0 else np.vstack((new, row))
        assert np.array_equal(imagegeo.expand_image_domain(a, xn, yn), new)


def test_image_to_geometry_colour_gradient_class(hostlib, monkeypatch):
    core = structure_png(str(hostlib / "structure.png"))
    monkeypatch.setenv("LBM_STRUCTURE_IMAGE", str(hostlib / "structure.png"))
    d = hostlib / "ini"; d.mkdir()
    txt = open(os.path.join(REF_INI, "cg2d", "RKtwophasesetup2D.ini")).read().replace("Existance = 'no'", "Existance = 'yes'")
    (d / "RKtwophasesetup2D.ini").write_text(txt)
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    sim = RKColorGradientLBM(str(d), verbose=False)
    nb, ratio = sim.numBufferingLayers, sim.ratioTopToBottom
    assert nb > 0
    sim.initializeDomainBorder()
    low = int(2 * nb * ratio)
    assert sim.isDomain.shape == (core.shape[0] + 2 * nb, core.shape[1])
    assert sim.isDomain[:low].all() and sim.isDomain[-(2 * nb - low):].all()                     # void buffer rows
    body = sim.isDomain[low:low + core.shape[0]]
    assert not body[:, 0].any() and not body[:, -1].any()                                        # solid side columns
    assert np.array_equal(body[:, 1:-1], core[:, 1:-1] != 0)
    sim.initializeDomainCondition()
    rows = np.indices(sim.isDomain.shape)[0]
    assert np.array_equal(sim.fluidsRhoR > 0, sim.isDomain & (rows < sim.yDomain - nb))          # RKD2Q9.py:511-520
    assert np.array_equal(sim.fluidsRhoB > 0, sim.isDomain & (rows >= sim.yDomain - nb))
    sim.timeSteps, sim.timeInterval = 6, 3
    sim.runRKColorGradient2D()
    assert np.isfinite(sim.fluidsRhoR).all() and sim.fluidsRhoR.shape == sim.isDomain.shape


def test_result_files_and_drainage_imbibition_restart(hostlib, monkeypatch):
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    base = open(os.path.join(REF_INI, "cg2d", "RKtwophasesetup2D.ini")).read()
    first = RKColorGradientLBM(os.path.join(REF_INI, "cg2d"), verbose=False)
    first.runRKColorGradient2D()
    n_rec = (first.timeSteps - 1) // first.timeInterval
    names = ["/FluidMacro/FluidDensityRin%d" % n_rec, "/FluidVelocity/FluidVelocityYAt%d" % n_rec, "/FluidPDF/FluidPDFBat%d" % n_rec]
    rec = results.read_arrays(str(hostlib / "results"), "SimulationResultsRK.h5", names)           # the reference's dataset names
    assert rec[names[0]].shape == first.isDomain.shape and rec[names[2]].shape == first.isDomain.shape + (9,)
    # second run: IsCycle = 'yes' restarts from record LastStep of ~/LBMInitial/SimulationResultsRK.h5 (RKD2Q9.py:491-508)
    monkeypatch.setenv("LBM_INITIAL_DIR", str(hostlib / "results"))
    d = hostlib / "ini2"; d.mkdir()
    (d / "RKtwophasesetup2D.ini").write_text(base.replace("IsCycle = 'no'", "IsCycle = 'yes'\nLastStep = %d" % n_rec))
    second = RKColorGradientLBM(str(d), verbose=False)
    assert second.isCycles == "'yes'" and second.lastStep == n_rec
    second.initializeDomainBorder(); second.initializeDomainCondition()
    np.testing.assert_array_equal(second.fluidsRhoR[:-20], rec[names[0]][:-20] * second.isDomain[:-20])
    assert (second.fluidsRhoR[-20:] == 0).all() and (second.fluidsRhoB[-20:][second.isDomain[-20:]] == second.initialRhoB).all()
    np.testing.assert_allclose(second.fluidPDFR.sum(-1), second.fluidsRhoR, atol=1e-14)            # equilibrium populations
    second.timeSteps = 4
    second.runRKColorGradient2D()
    assert np.isfinite(second.fluidsRhoB).all()


def test_image_cycle_restart_swaps_the_colours_in_the_inlet_rows(hostlib, monkeypatch):
    structure_png(str(hostlib / "structure.png"))
    monkeypatch.setenv("LBM_STRUCTURE_IMAGE", str(hostlib / "structure.png"))
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    txt = open(os.path.join(REF_INI, "cg2d", "RKtwophasesetup2D.ini")).read().replace("Existance = 'no'", "Existance = 'yes'")
    d = hostlib / "a"; d.mkdir(); (d / "RKtwophasesetup2D.ini").write_text(txt)
    a = RKColorGradientLBM(str(d), verbose=False)
    a.timeSteps, a.timeInterval = 5, 5
    a.runRKColorGradient2D()
    a.saveCycleInitial()
    d2 = hostlib / "b"; d2.mkdir(); (d2 / "RKtwophasesetup2D.ini").write_text(txt.replace("IsCycle = 'no'", "IsCycle = 'yes'\nLastStep = 0"))
    b = RKColorGradientLBM(str(d2), verbose=False)
    b.initializeDomainBorder(); b.initializeDomainCondition()
    nb = b.numBufferingLayers
    np.testing.assert_array_equal(b.fluidsRhoR[:-nb], a.fluidsRhoR[:-nb]); np.testing.assert_array_equal(b.fluidsRhoR[-nb:], a.fluidsRhoB[-nb:])
    np.testing.assert_array_equal(b.fluidPDFB[-nb:], a.fluidPDFR[-nb:]); np.testing.assert_array_equal(b.physicalVY, a.physicalVY)


def test_shan_chen_image_duplicate_and_cycles(hostlib, monkeypatch):
    core = structure_png(str(hostlib / "structure.png"))
    monkeypatch.setenv("LBM_STRUCTURE_IMAGE", str(hostlib / "structure.png"))
    monkeypatch.setenv("LBM_DUPLICATE_X", "2"); monkeypatch.setenv("LBM_DUPLICATE_Y", "2")
    from openlbmpm_b200.ShanChenD2Q9 import ShanChenD2Q9
    src = os.path.join(REF_INI, "sc")
    d = hostlib / "sc"; d.mkdir()
    two = open(os.path.join(src, "twophasesetup.ini")).read()
    (d / "shanchen2D.ini").write_text(open(os.path.join(src, "shanchen2D.ini")).read())
    import re
    two = two.replace("Exist = 'no'", "Exist = 'yes'")
    two = re.sub(r"(\[DuplicateDomain\]\s*Option = )'no'", r"\1'yes'", two)
    (d / "twophasesetup.ini").write_text(two)
    sim = ShanChenD2Q9(str(d), verbose=False)
    sim.numTimeStep = 3
    sim.runTypeSCmodel()
    assert sim.isDomain.shape == (2 * core.shape[0] + 40, 2 * core.shape[1])
    assert sim.isDomain[:20].all() and sim.isDomain[-20:].all() and not sim.isDomain[20:-20, 0].any()
    tile = imagegeo.expand_image_domain(core.astype(float), 2, 2) != 0
    assert np.array_equal(sim.isDomain[20:-20, 1:-1], tile[:, 1:-1])
    assert np.isfinite(sim.fluidsDensity).all()
    # next cycle: a third fluid... the reference keeps nf fluids; the last one is the NEW fluid entering through the inlet rows
    monkeypatch.setenv("LBM_INITIAL_DIR", str(hostlib / "results"))
    two2 = re.sub(r"(\[DICycles\]\s*Option = )'no'", r"\1'yes'\nLastStep = 0", two)
    d2 = hostlib / "sc2"; d2.mkdir(); (d2 / "twophasesetup.ini").write_text(two2); (d2 / "shanchen2D.ini").write_text((d / "shanchen2D.ini").read_text())
    nxt = ShanChenD2Q9(str(d2), verbose=False)
    nxt.initializeDomainBorder(); nxt.initializeDomainCondition()
    rec = results.read_arrays(str(hostlib / "results"), "SimulationResults.h5", ["/FluidMacro/FluidDensityType0in0"])
    np.testing.assert_array_equal(nxt.fluidsDensity[0, :-30], rec["/FluidMacro/FluidDensityType0in0"][:-30] * nxt.isDomain[:-30])
    assert (nxt.fluidsDensity[1, -30:][nxt.isDomain[-30:]] == nxt.initialDensities[1]).all()


def test_asynchronous_output_writes_the_same_records(hostlib, monkeypatch):
    """LBM_ASYNC_OUTPUT=1: the records written behind the step loop equal those of the blocking path"""
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    from openlbmpm_b200.ShanChenD2Q9 import ShanChenD2Q9
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LBM_ASYNC_OUTPUT", mode)
        monkeypatch.setenv("LBM_RESULTS_DIR", str(hostlib / ("res" + mode)))
        cg = RKColorGradientLBM(os.path.join(REF_INI, "cg2d"), verbose=False)
        cg.runRKColorGradient2D()
        sc = ShanChenD2Q9(os.path.join(REF_INI, "efs"), verbose=False)
        sc.runTypeSCmodel()
        nrec = (cg.timeSteps - 1) // cg.timeInterval + 1
        names = ["/FluidMacro/FluidDensityRin%d" % r for r in range(nrec)] + ["/FluidVelocity/FluidVelocityYAt%d" % r for r in range(nrec)]
        runs[mode] = (results.read_arrays(str(hostlib / ("res" + mode)), "SimulationResultsRK.h5", names),
                      results.read_arrays(str(hostlib / ("res" + mode)), "SimulationResults.h5",
                                          ["/FluidMacro/FluidDensityType1in0", "/FluidVelocity/FluidVelocityXAt0"]),
                      cg.fluidsRhoR.copy())
    for a, b in zip(runs["0"][:2], runs["1"][:2]):
        assert a.keys() == b.keys()
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_array_equal(runs["0"][2], runs["1"][2])


def test_async_download_in_lock_step_with_blocking_download(hostlib):
    import cases
    cases.check_async_output(hostcheck_build.build())
