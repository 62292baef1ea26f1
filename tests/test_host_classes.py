"""Host-side mirror of the reference's class surface (RKColorGradientLBM, ShanChenD2Q9, RKColorGradient3D, main.py):
ini parsing, attribute names, index export, result files.  CPU tier: the classes are pointed at the host test
hook (tests/hostcheck) instead of liblbmpm.so; the GPU tier (test_gpu_classes.py) runs them unmodified."""
import glob
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build

from openlbmpm_b200 import _lib
from oracle import cg2d

REF_INI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ini")


@pytest.fixture()
def hostlib(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", hostcheck_build.build())
    monkeypatch.setenv("LBM_RESULTS_DIR", str(tmp_path / "results"))
    return tmp_path


def run_cg2d(hostlib):
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    sim = RKColorGradientLBM(os.path.join(REF_INI, "cg2d"), verbose=False)
    sim.runModifiedRKColorGradient2D()
    return sim


def test_cg2d_class_runs_with_reference_style_ini(hostlib):
    sim = run_cg2d(hostlib)
    assert sim.relaxationType == "'MRT'" and sim.boundaryTypeInlet == "'Neumann'"
    assert sim.surfaceTension == 0.1 and sim.wettingType == 2       # `SurfaceTension` spelling of the shipped ini
    assert sim.fluidsRhoR.shape == (sim.yDomain, sim.xDomain) and sim.fluidPDFR.shape == (sim.yDomain, sim.xDomain, 9)
    assert np.isfinite(sim.fluidsRhoR).all() and np.isfinite(sim.physicalVY).all()
    idx = cg2d.build_indexing(sim.isDomain)
    for k in ("fluidNodes", "neighboringNodes", "wettingSolidNodes", "fluidNodesWithSolidGPU"):
        assert np.array_equal(getattr(sim, k), idx[k]), k
    np.testing.assert_allclose(sim.nsX, idx["nsX"], atol=1e-15)
    files = glob.glob(str(hostlib / "results" / "SimulationResultsRK*"))
    assert files, "no result file written"
    # trajectory equals the oracle driven with the same parameters
    ref = cg2d.CG2D(sim.isDomain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0,
                    tautype=2, relax="MRT", inlet="Neumann", outlet="Dirichlet", vy_inlet=-1e-4, dBL=1.0, dRL=5e-8)
    yy, xx = np.indices(sim.isDomain.shape)
    red = (np.sqrt((yy - sim.yDomain // 2) ** 2 + (xx - sim.xDomain // 2) ** 2) <= 16.) & sim.isDomain
    ref.set_densities(np.where(red, 1.0, 0.0) * sim.isDomain, np.where(red, 0.0, 1.0) * sim.isDomain)
    ref.step(sim.timeSteps)
    ref.head()
    d = ref.to_dense()
    np.testing.assert_allclose(sim.fluidsRhoR, d["rhoR"], atol=1e-9)
    np.testing.assert_allclose(sim.physicalVY, d["uy"], atol=1e-9)


@pytest.mark.parametrize("which", ["sc", "efs"])
def test_shanchen_class(hostlib, which):
    from openlbmpm_b200.ShanChenD2Q9 import ShanChenD2Q9
    sim = ShanChenD2Q9(os.path.join(REF_INI, which), verbose=False)
    sim.runTypeSCmodel()
    assert sim.fluidsDensity.shape == (2, sim.ny, sim.nx) and sim.fluidPDF.shape == (2, sim.ny, sim.nx, 9)
    assert sim.neighboringNodes.min() == -1 and sim.fluidNodes.size == sim.voidSpace
    assert np.isfinite(sim.fluidsDensity).all()
    m = sim.fluidsDensity.sum(axis=(1, 2))
    assert m[0] > 0 and m[1] > 0


def test_cg3d_class_and_main(hostlib, monkeypatch):
    import main
    monkeypatch.chdir(os.path.join(REF_INI, "cg3d"))
    assert main.main(["3D", "flow", "CG", "--ini", os.path.join(REF_INI, "cg3d")]) == 0
    assert main.main(["2D", "transport", "SC"]) == 2


def test_sc3d_class_and_main(hostlib):
    """`ShanChenD3Q19(ini).runEFS4LBM3DGPU()` (main.py:73-77): the class the reference names but does not ship"""
    import main
    from openlbmpm_b200.ShanChenD3Q19 import ShanChenD3Q19
    from oracle import sc_dense
    sim = ShanChenD3Q19(os.path.join(REF_INI, "efs3d"), verbose=False)
    sim.runEFS4LBM3DGPU()
    assert sim.fluidsDensity.shape == (2, 24, 10, 12) and sim.fluidPDF.shape == (2, 24, 10, 12, 19)
    assert sim.neighboringNodes.size == 18 * sim.fluidNodes.size and sim.physicalVZ.shape == (24, 10, 12)
    ref = sc_dense.SCDense(sc_dense.d3q19(), sim.isDomain, model="EFS", relax="MRT", tau=(1., 1.), G=0.2, Gs=(-0.14, 0.14))
    reg = np.indices((24, 10, 12))[0] < 14
    ref.set_densities(np.stack([np.where(reg, 1.0, 0.02), np.where(reg, 0.02, 1.0)]))
    ref.step(sim.numTimeStep + 1)               # the reference's loops run numTimeStep + 1 iterations
    np.testing.assert_allclose(sim.fluidsDensity, ref.rho, atol=1e-10)
    np.testing.assert_allclose(sim.physicalVZ, ref.uph[2], atol=1e-10)
    assert main.main(["3D", "flow", "SC", "--ini", os.path.join(REF_INI, "efs3d")]) == 0


def test_missing_key_exits_like_the_reference(hostlib, tmp_path):
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    d = tmp_path / "bad"; d.mkdir()
    (d / "RKtwophasesetup2D.ini").write_text("[DomainSize]\nxDomain = 8\n")
    with pytest.raises(SystemExit):
        RKColorGradientLBM(str(d))


def test_perturbation_classes(hostlib):
    """SurfaceTensionType 'Perturbation' (RKD2Q9.py:979-1223; the flavour the reference's 3-D ini parameterises): the
    classes drive the engine's perturbation operator; trajectories equal the oracle run with the ini's numbers"""
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    from openlbmpm_b200.RKColorGradientD3Q19 import RKColorGradient3D
    from oracle import cgp_dense
    sim = RKColorGradientLBM(os.path.join(REF_INI, "cgp2d"), verbose=False)
    yy, xx = np.indices((22, 18))
    sim.initialRedRegion = np.sqrt((yy - 12.3) ** 2 + (xx - 8.6) ** 2) <= 5.
    sim.runRKColorGradient2D()
    ref = cgp_dense.CGPDense(cgp_dense.d2q9(), sim.isDomain, beta=0.7, AkR=1.4e-2, AkB=1.0e-2, tauR=1.0, tauB=0.8,
                             solid_phi=0.4, body_force=(1.0e-5, -2.0e-5))
    red = sim.initialRedRegion & sim.isDomain
    ref.set_densities(np.where(red, 1.0, 0.0) * sim.isDomain, np.where(red, 0.0, 1.0) * sim.isDomain)
    ref.step(sim.timeSteps)
    assert 0.05 < sim.fluidsRhoR.mean() < 0.5
    np.testing.assert_allclose(sim.fluidsRhoR, ref.rhoR[0], atol=1e-9)
    np.testing.assert_allclose(sim.physicalVX, ref.u[0, 0], atol=1e-9)
    sim3 = RKColorGradient3D(os.path.join(REF_INI, "cgp3d"), verbose=False)
    zz = np.indices((40, 10, 12))[0]
    sim3.initialRedRegion = zz < 20
    sim3.runRKColorGradient3D()
    ref3 = cgp_dense.CGPDense(cgp_dense.d3q19(), sim3.isDomain, beta=1.0, AkR=7e-3, AkB=7e-3, solid_phi=0.7)
    ref3.set_densities(np.where(zz < 20, 1.0, 0.0), np.where(zz < 20, 0.0, 1.0))
    ref3.step(sim3.timeSteps)
    np.testing.assert_allclose(sim3.fluidsRhoR, ref3.rhoR, atol=1e-9)
    np.testing.assert_allclose(sim3.physicalVZ, ref3.u[2], atol=1e-9)


def test_reference_3d_ini_runs_the_perturbation_operator(hostlib):
    """IniFiles/RKtwophasesetup3D.ini as shipped: no [SurfaceTension] section, AkR / AkB, SolidRhoR / SolidRhoB, velocity inlet,
    pressure outlet, SRT.  The class selects the perturbation operator with those numbers; the run equals the oracle's."""
    from openlbmpm_b200.RKColorGradientD3Q19 import RKColorGradient3D
    from oracle import cgp_dense
    sim = RKColorGradient3D(os.path.join(REF_INI, "cgp3d_ref"), verbose=False)
    assert sim.surfaceTensionType == "'Perturbation'" and sim.relaxationType == "'SRT'"
    assert sim.solidPhi == pytest.approx(0.75) and sim.boundaryTypeInlet == "'Neumann'" and sim.boundaryTypeOutlet == "'Dirichlet'"
    shape = (36, 10, 12)
    zz = np.indices(shape)[0]
    sim.initialRedRegion = zz < 26
    sim.runRKColorGradient3D()
    ref = cgp_dense.CGPDense(cgp_dense.d3q19(), sim.isDomain, beta=1.0, AkR=7e-3, AkB=7e-3, tauR=1.0, tauB=0.9, solid_phi=0.75,
                             relax="SRT", inlet="Neumann", outlet="Dirichlet", v_inlet=-1.0e-3, dBL=1.0, dRL=1.0e-8)
    red = (zz < 26) & sim.isDomain
    ref.set_densities(np.where(red, 1.0, 0.0) * sim.isDomain, np.where(red, 0.0, 1.0) * sim.isDomain)
    ref.step(sim.timeSteps)
    assert np.isfinite(sim.fluidsRhoR).all()
    np.testing.assert_allclose(sim.fluidsRhoR, ref.rhoR, atol=2e-7)        # trace-colour noise, see cases.case_cgp_open
    np.testing.assert_allclose(sim.fluidsRhoB, ref.rhoB, atol=2e-7)
    np.testing.assert_allclose(sim.physicalVZ, ref.u[2], atol=2e-7)
    assert np.abs(sim.physicalVZ[-1][sim.isDomain[-1]] + 1.0e-3).max() < 1e-12        # the inlet plane carries the prescribed velocity


def test_transport_class_five_velocity_branch(hostlib):
    """NumberSchemes = 5 on an open channel: three reacting tracers, Inamuro inlet row, free-flow outlet row; equals the oracle"""
    from openlbmpm_b200.Transport2DRK import Transport2DRK
    from oracle import cg_dense, tr_dense
    sim = Transport2DRK(os.path.join(REF_INI, "tr2d_q5"), verbose=False)
    assert sim.numSchemes == 5 and sim.reaction == "'yes'" and sim.typeTracerOutletBoundary == "'Freeflow'"
    yy, xx = np.indices((sim.yDomain, sim.xDomain))
    sim.initialRedRegion = yy >= sim.yDomain - 12
    sim.runTransport2DMPMCRK()
    flow = cg_dense.CGDense(cg_dense.d2q9(), sim.isDomain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98, tauR=1.0,
                            tauB=1.0, tautype=2, relax="MRT", inlet="Neumann", outlet="Dirichlet", v_inlet=-1.0e-3, dBL=1.0, dRL=5e-8)
    red = sim.initialRedRegion & sim.isDomain
    flow.set_densities(np.where(red, 1.0, 0.0) * sim.isDomain, np.where(red, 0.0, 1.0) * sim.isDomain)
    tr = tr_dense.TracerDenseQ5(flow, dxx=(0.05, 0.1, 0.07), dyy=(0.08, 0.1, 0.07), dxy=(0.01, 0.0, 0.0), dyx=(0.02, 0.0, 0.0),
                                beta=(0.6, 0.2, 0.0), reaction_rate=0.05, diff_j=(0.3, 0.3333, 0.4), inlet_conc=(0.7, 0.2, 0.0),
                                freeflow_outlet=True)
    reg = yy <= sim.yDomain - sim.numBufferingLayers
    tr.set_concentrations(np.stack([np.where(reg, c, 0.0) for c in (1.0, 0.5, 0.1)]))
    tr.step(sim.timeSteps + 1)
    np.testing.assert_allclose(sim.tracerConc, tr.conc[:, 0], atol=5e-9)
    assert np.abs(sim.tracerConc[:, -1][:, sim.isDomain[-1]] - np.array([[0.7], [0.2], [0.0]])).max() < 1e-13
    assert sim.tracerConc[2].sum() > 0.1 * sim.isDomain.sum() * 1.0001        # the product C has been formed


def test_transport_class_and_main(hostlib):
    """`Transport2DRK(ini).runTransport2DMPMCRK()` (main.py:66-68): flow + 2 tracers, MRT tracers; equals the oracle"""
    import main
    from openlbmpm_b200.Transport2DRK import Transport2DRK
    from oracle import cg_dense, tr_dense
    from openlbmpm_b200 import results
    ini = os.path.join(REF_INI, "tr2d")
    sim = Transport2DRK(ini, verbose=False)
    yy, xx = np.indices((sim.yDomain, sim.xDomain))
    sim.initialRedRegion = np.sqrt((yy - 20.3) ** 2 + (xx - 9.6) ** 2) <= 6.
    sim.runTransport2DMPMCRK()
    assert sim.tracerConc.shape == (2, sim.yDomain, sim.xDomain) and np.isfinite(sim.tracerConc).all()
    flow = cg_dense.CGDense(cg_dense.d2q9(), sim.isDomain, sigma=0.1, theta_deg=60.0, wetting=2, beta=0.7, delta=0.98, tauR=1.0,
                            tauB=1.0, tautype=2, relax="MRT")
    red = sim.initialRedRegion & sim.isDomain
    flow.set_densities(np.where(red, 1.0, 0.0) * sim.isDomain, np.where(red, 0.0, 1.0) * sim.isDomain)
    tr = tr_dense.TracerDense(flow, relax="MRT", tau=(0.8, 1.1), dxx=(0.05, 0.1), dyy=(0.08, 0.1), dxy=(0.01, 0.0), dyx=(0.02, 0.0),
                              beta=(0.6, 0.2))
    reg = yy <= sim.yDomain - sim.numBufferingLayers
    tr.set_concentrations(np.stack([np.where(reg, 1.0, 0.0), np.where(reg, 0.5, 0.0)]))
    tr.step(sim.timeSteps + 1)          # the class's last download shows iteration timeSteps, tracer phase included
    np.testing.assert_allclose(sim.tracerConc, tr.conc[:, 0], atol=1e-9)
    rec = results.read_arrays(str(hostlib / "results"), "ConcentrationResults.h5", ["/TransportMacro/TracerConcType1in0"])
    assert rec["/TransportMacro/TracerConcType1in0"].shape == (sim.yDomain, sim.xDomain)
    assert main.main(["2D", "transport", "CG", "--ini", ini]) == 0
    assert main.main(["2D", "transport", "SC", "--ini", ini]) == 2
