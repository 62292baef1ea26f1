"""Slab decomposition on the CPU tier.  The host test hook replaces NCCL by an in-process ring (host_stubs.cu): the
"ranks" are threads of this process, one handle each.  P slabs must reproduce the single-slab run BIT FOR BIT (pull
streaming makes halo arithmetic identical to interior arithmetic), for every model and kernel path -- this checks which
ghost planes the step loops exchange and when, and who owns the open-boundary rows; NCCL itself and the tiled kernels are
checked on the GPUs by tests/mgpu_check.py."""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
from openlbmpm_b200 import _lib

pytestmark = pytest.mark.timeout(300)


@pytest.fixture(scope="module")
def lib():
    return hostcheck_build.build()


def run(lib, lattice, dom, rho, steps, world, reinit=False, **kw):
    """-> stacked [densities..., velocity components...] of the whole lattice after sum(steps) steps"""
    n_flow = dom.shape[0]
    assert n_flow % world == 0
    t = n_flow // world
    engines = [_lib.Engine(lattice, (t,) + dom.shape[1:], lib_path=lib, **kw) for _ in range(world)]
    if world > 1:
        uid = engines[0].nccl_unique_id()
        for r, e in enumerate(engines):
            e.comm_init(r, world, uid)
    out, errors = [None] * world, []

    def work(r):
        try:
            sl = slice(r * t, (r + 1) * t)
            e = engines[r]
            e.set_geometry(dom[sl])
            e.init_equilibrium(*[np.where(dom[sl], a[sl], 0.0) for a in rho])
            if reinit:        # a first run, then the same state again: the factored buffers and the neighbours' mappings are reused
                e.step(3)
                e.init_equilibrium(*[np.where(dom[sl], a[sl], 0.0) for a in rho])
            parts = []
            for n in steps:
                e.step(n)
                d, u = e.download_macros()
                parts.append(np.stack(d + u))
            pdf = e.download_pdfs()
            out[r] = (parts, np.stack(pdf), e.total_mass())
            e.close()                 # collective on slabs with the one-sided exchange (a last handshake): from the rank's own thread
        except Exception as ex:       # a rank that dies would leave the others in the barrier: report and bail out
            errors.append(ex)
            os._exit(3) if world > 1 else None

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    for e in engines:
        e.close()
    macros = [np.concatenate([out[r][0][k] for r in range(world)], axis=1) for k in range(len(steps))]
    return macros, np.concatenate([out[r][1] for r in range(world)], axis=1), sum(out[r][2] for r in range(world))


def geometry(shape, solid, open_ends):
    dom = np.ones(shape, bool)
    if solid:
        idx = np.indices(shape)
        c = [s / 2 for s in shape]
        dom = sum((idx[a] - c[a] + (0.5 if a == 0 else 0.0)) ** 2 for a in range(len(shape))) > 6.0     # straddles the seam at n/2
        if not open_ends:
            dom &= sum((idx[a] - (1 if a == 0 else 3)) ** 2 for a in range(len(shape))) > 3.0           # and the periodic seam
    return dom


def compare(lib, lattice, shape, steps, solid=True, worlds=(2, 3), ncomp=2, **kw):
    open_ends = kw.get("inlet", 0) != 0 or kw.get("outlet", 0) != 0
    dom = geometry(shape, solid, open_ends)
    rng = np.random.default_rng(3)
    base = 0.5 + 0.3 * (rng.random(shape) - 0.5)
    rho = [base, 1.0 - base] + [0.3 + 0.1 * rng.random(shape) for _ in range(ncomp - 2)]
    ref = run(lib, lattice, dom, rho[:ncomp], steps, 1, **kw)
    for P in worlds:
        if shape[0] % P or shape[0] // P < 4:
            continue
        got = run(lib, lattice, dom, rho[:ncomp], steps, P, **kw)
        for k in range(len(steps)):
            assert np.array_equal(got[0][k], ref[0][k]), "P = %d differs after chunk %d (max %.3e)" % (
                P, k, np.abs(got[0][k] - ref[0][k]).max())
        assert np.array_equal(got[1], ref[1]), "populations differ for P = %d" % P
        np.testing.assert_allclose(got[2], ref[2], rtol=1e-13)


OPEN = dict(inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_CONVECTIVE, inlet_velocity=-2.0e-3)
OPEN_P = dict(inlet=_lib.INLET_PRESSURE, outlet=_lib.OUTLET_PRESSURE, rhoRH=1.004, rhoBH=5e-8, rhoRL=5e-8, rhoBL=1.0)


@pytest.mark.parametrize("lattice,shape", [(19, (24, 6, 8)), (9, (24, 10))])
@pytest.mark.parametrize("name,kw", [
    ("fast path", dict(contact_angle_deg=70.0)),
    ("reference-ordered kernels", dict(flags=1, contact_angle_deg=50.0)),
    ("SRT, unequal viscosities", dict(relax=_lib.RELAX_SRT, tauR=1.0, tauB=0.8, tau_type=1)),
    ("open channel, velocity inlet + convective outlet, fast path", dict(OPEN, contact_angle_deg=60.0)),
    ("open channel, reference-ordered kernels", dict(OPEN, flags=1)),
    ("open channel, pressure inlet + pressure outlet", dict(OPEN_P, contact_angle_deg=110.0)),
    ("perturbation operator", dict(surface_tension_type=_lib.ST_PERTURBATION, AkR=8e-3, AkB=1e-2, solid_phi=0.4, tauB=0.85)),
])
def test_colour_gradient_slabs_bit_equal(lattice, shape, name, kw, lib):
    compare(lib, lattice, shape, [1, 2, 6], **kw)


@pytest.mark.parametrize("lattice,shape", [(19, (24, 6, 8)), (9, (24, 10)), (19, (24, 8, 32))])
@pytest.mark.parametrize("name,kw", [
    ("fast path", dict(contact_angle_deg=70.0)),
    ("open channel, velocity inlet + convective outlet", dict(OPEN, contact_angle_deg=60.0)),
])
def test_one_sided_exchange_slabs_bit_equal(lattice, shape, name, kw, lib):
    """LBM_FLAG_PEER_EXCHANGE: the two per-step exchanges of the fast path as stores into the neighbours' ghost planes + a
    release / acquire flag pair instead of the send / recv rendezvous (thread ranks: the neighbours' memory is a pointer, the
    flags are std::atomic); P = 2, 3 slabs bit-equal to one slab, several lbm_step calls with downloads in between; (24, 8, 32)
    runs the tiled kernels on every slab"""
    compare(lib, lattice, shape, [1, 2, 6, 5], **dict(kw, flags=_lib.FLAG_PEER_EXCHANGE))


@pytest.mark.parametrize("lattice,shape", [(19, (24, 6, 8)), (9, (24, 10)), (19, (24, 8, 32))])
@pytest.mark.parametrize("name,kw", [
    ("fast path", dict(contact_angle_deg=70.0)),
    ("open channel, velocity inlet + convective outlet", dict(OPEN, contact_angle_deg=60.0)),
])
def test_send_recv_exchange_slabs_bit_equal(lattice, shape, name, kw, lib):
    """LBM_FLAG_NCCL_EXCHANGE: the one-sided exchange is the default on slabs of equal extents; this flag keeps the send / recv
    ring for the fast path's exchanges too (what GPUs without peer access fall back to)"""
    compare(lib, lattice, shape, [1, 2, 6, 5], **dict(kw, flags=_lib.FLAG_NCCL_EXCHANGE))


@pytest.mark.parametrize("name,kw", [
    ("closed box", dict(contact_angle_deg=70.0)),
    ("open channel, velocity inlet + convective outlet", dict(OPEN, contact_angle_deg=60.0)),
    ("send / recv ring", dict(contact_angle_deg=70.0, flags=_lib.FLAG_NCCL_EXCHANGE)),
])
def test_d2q9_tile_kernels_on_slabs_bit_equal(name, kw, lib):
    """2-D lattice whose slabs admit the 32 x 8 tile kernels (ghost rows instead of the index wrap of a single slab)"""
    compare(lib, 9, (48, 32), [1, 2, 6], worlds=(2, 3), **kw)


def test_reinitialised_slabs_bit_equal(lib):
    """lbm_init_equilibrium on slabs that are already on the fast path (one-sided exchange: buffers stay mapped)"""
    shape = (24, 8, 32)
    dom = geometry(shape, True, False)
    rng = np.random.default_rng(3)
    base = 0.5 + 0.3 * (rng.random(shape) - 0.5)
    ref = run(lib, 19, dom, [base, 1.0 - base], [2, 3], 1, contact_angle_deg=70.0)
    for P in (2, 3):
        got = run(lib, 19, dom, [base, 1.0 - base], [2, 3], P, reinit=True, contact_angle_deg=70.0)
        assert all(np.array_equal(a, b) for a, b in zip(got[0], ref[0])) and np.array_equal(got[1], ref[1]), P


def test_all_fluid_box_slabs_bit_equal(lib):
    compare(lib, 19, (24, 6, 8), [3, 5], solid=False, worlds=(2, 3, 6))


SC_OPEN = dict(inlet=_lib.INLET_VELOCITY, sc_inlet_velocity=[0.0, -5.0e-4], sc_rho_out=[1.0, 0.02])


@pytest.mark.parametrize("lattice,shape", [(9, (24, 12)), (19, (24, 6, 8))])
@pytest.mark.parametrize("name,kw", [
    ("original Shan-Chen", dict(model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, sc_G=[0, 0.9, 0, 0, 0.9, 0])),
    ("explicit forcing SRT", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, sc_G=[0, 0.15, 0, 0, 0.15, 0])),
    ("explicit forcing MRT", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, sc_G=[0, 0.15, 0, 0, 0.15, 0])),
])
def test_shan_chen_slabs_bit_equal(lattice, shape, name, kw, lib):
    compare(lib, lattice, shape, [1, 2, 6], n_components=2, sc_tau=[1.0, 0.9], sc_Gsolid=[-0.1, 0.1], **kw)


def test_shan_chen_d3q19_open_channel_slabs_bit_equal(lib):
    for kw in (dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, outlet=_lib.OUTLET_PRESSURE),
               dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, outlet=_lib.OUTLET_CONVECTIVE),
               dict(model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, outlet=_lib.OUTLET_CONVECTIVE, sc_G=[0, 0.9, 0, 0, 0.9, 0])):
        par = dict(SC_OPEN, n_components=2, sc_tau=[1.0, 0.9], sc_G=[0, 0.15, 0, 0, 0.15, 0], sc_Gsolid=[-0.1, 0.1])
        par.update(kw)
        compare(lib, 19, (24, 6, 8), [1, 2, 6], **par)


@pytest.mark.parametrize("name,kw", [
    ("isotropy 8", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, sc_isotropy=8)),
    ("isotropy 10", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, sc_isotropy=10)),
    ("original Shan-Chen, velocity inlet + convective outlet", dict(SC_OPEN, model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, outlet=_lib.OUTLET_CONVECTIVE,
                                                                     sc_G=[0, 0.9, 0, 0, 0.9, 0])),
    ("explicit forcing, velocity inlet + pressure outlet", dict(SC_OPEN, model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, outlet=_lib.OUTLET_PRESSURE)),
    ("explicit forcing, velocity inlet + convective outlet", dict(SC_OPEN, model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, outlet=_lib.OUTLET_CONVECTIVE)),
    ("explicit forcing isotropy 8, deep boundary rows", dict(SC_OPEN, model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, outlet=_lib.OUTLET_PRESSURE, sc_isotropy=8)),
    ("three components", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, n_components=3, sc_tau=[1.0, 0.9, 1.1],
                              sc_G=[0, 0.1, 0.12, 0, 0.1, 0, 0.08, 0, 0.12, 0.08, 0, 0], sc_Gsolid=[-0.1, 0.1, 0.0])),
])
def test_shan_chen_d2q9_variants_slabs_bit_equal(name, kw, lib):
    par = dict(n_components=2, sc_tau=[1.0, 0.9], sc_G=[0, 0.15, 0, 0, 0.15, 0], sc_Gsolid=[-0.1, 0.1])
    par.update(kw)
    compare(lib, 9, (24, 12), [1, 2, 6], ncomp=par["n_components"], **par)


# ---- the host classes on slabs: `torchrun --nproc-per-node P main.py 3D flow CG` (slab.from_environment); here the ranks
# are threads (slab.ThreadSlabs) over the in-process ring ----
def run_class(cls, ini_dir, world, monkeypatch, tmp_path, edit=None):
    from openlbmpm_b200 import slab
    monkeypatch.setattr(_lib, "LIB_PATH", hostcheck_build.build())
    monkeypatch.setenv("LBM_RESULTS_DIR", str(tmp_path / ("results%d" % world)))
    shared = slab.ThreadSlabs.Shared(world)
    sims, errors = [None] * world, []

    def work(r):
        try:
            sim = cls(ini_dir, verbose=False)
            if world > 1:
                sim.slabs = slab.ThreadSlabs(shared, r)
            if edit:
                edit(sim)
            getattr(sim, "runRKColorGradient2D", None) and sim.runRKColorGradient2D()
            getattr(sim, "runTypeSCmodel", None) and sim.runTypeSCmodel()
            sims[r] = sim
        except BaseException as ex:
            errors.append(ex)
            if world > 1:
                os._exit(3)
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    return sims


@pytest.mark.parametrize("which", ["cg3d", "cg2d", "cgp3d"])
def test_host_classes_on_slabs(which, monkeypatch, tmp_path):
    from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
    from openlbmpm_b200.RKColorGradientD3Q19 import RKColorGradient3D
    ini = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ini", which)
    cls = RKColorGradientLBM if which == "cg2d" else RKColorGradient3D

    def edit(sim):
        shape = sim._shape()
        idx = np.indices(shape)
        sim.initialRedRegion = sum((idx[a] - shape[a] / 2 + 0.5) ** 2 for a in range(len(shape))) < (min(shape) / 3.0) ** 2
    one = run_class(cls, ini, 1, monkeypatch, tmp_path, edit)[0]
    two = run_class(cls, ini, 2, monkeypatch, tmp_path, edit)
    for sim in two:                                   # every rank ends with the whole lattice
        assert np.array_equal(sim.fluidsRhoR, one.fluidsRhoR) and np.array_equal(sim.fluidsRhoB, one.fluidsRhoB)
        assert np.array_equal(sim.physicalVX, one.physicalVX) and np.array_equal(sim.fluidPDFB, one.fluidPDFB)
    import glob
    assert len(glob.glob(str(tmp_path / "results2" / "*"))) == len(glob.glob(str(tmp_path / "results1" / "*"))) > 0   # rank 0 wrote


@pytest.mark.parametrize("which", ["sc", "efs", "efs3d"])
def test_shan_chen_classes_on_slabs(which, monkeypatch, tmp_path):
    from openlbmpm_b200.ShanChenD2Q9 import ShanChenD2Q9
    from openlbmpm_b200.ShanChenD3Q19 import ShanChenD3Q19
    ini = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ini", which)
    cls = ShanChenD3Q19 if which == "efs3d" else ShanChenD2Q9
    if which == "efs3d":
        monkeypatch.setattr(ShanChenD3Q19, "runTypeSCmodel", ShanChenD3Q19.runEFS4LBM3DGPU, raising=False)
    one = run_class(cls, ini, 1, monkeypatch, tmp_path)[0]
    two = run_class(cls, ini, 2, monkeypatch, tmp_path)
    for sim in two:
        assert np.array_equal(sim.fluidsDensity, one.fluidsDensity) and np.array_equal(sim.physicalVY, one.physicalVY)
        assert np.array_equal(sim.fluidPDF, one.fluidPDF)


def test_tracers_on_slabs_bit_equal(lib):
    """flow + tracers on P = 2, 3 slabs vs one slab: concentrations bit-equal (tracer ghost planes travel with the flow's)"""
    q5 = dict(n_schemes=5, reaction=False, inlet_type=_lib.TR_INLET_DIRICHLET, inlet_conc=(0.7, 0.1), outlet_type=_lib.TR_OUTLET_FREEFLOW)
    for lattice, shape, relax, trkw in ((9, (24, 12), _lib.RELAX_MRT, {}), (19, (24, 6, 8), _lib.RELAX_SRT, {}), (9, (24, 12), _lib.RELAX_MRT, q5)):
        dom = geometry(shape, True, False)
        rng = np.random.default_rng(9)
        rhoR = 0.5 + 0.3 * (rng.random(shape) - 0.5)
        conc = 0.2 + rng.random((2,) + shape)
        ref = None
        for world in (1, 2, 3):
            t = shape[0] // world
            engines = [_lib.Engine(lattice, (t,) + shape[1:], lib_path=lib, contact_angle_deg=70.0) for _ in range(world)]
            if world > 1:
                uid = engines[0].nccl_unique_id()
                for r, e in enumerate(engines):
                    e.comm_init(r, world, uid)
            out = [None] * world

            def work(r):
                try:
                    sl = slice(r * t, (r + 1) * t)
                    e = engines[r]
                    e.tracer_setup(n_tracers=2, relax=relax, tau=(0.8, 1.1), dxx=(0.05, 0.1), dyy=(0.08, 0.1), beta=(0.6, 0.3), **trkw)
                    e.set_geometry(dom[sl])
                    e.init_equilibrium(np.where(dom[sl], rhoR[sl], 0.0), np.where(dom[sl], 1.0 - rhoR[sl], 0.0))
                    e.tracer_init(*[np.where(dom[sl], c[sl], 0.0) for c in conc])
                    parts = []
                    for n in (1, 2, 5):
                        e.step(n)
                        parts.append(np.stack(e.tracer_download() + e.download_macros()[0]))
                    out[r] = parts
                    e.close()
                except BaseException:
                    import traceback; traceback.print_exc()
                    os._exit(3)
            threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            for e in engines:
                e.close()
            got = [np.concatenate([out[r][k] for r in range(world)], axis=1) for k in range(3)]
            if ref is None:
                ref = got
            else:
                for k in range(3):
                    assert np.array_equal(got[k], ref[k]), "P = %d, chunk %d: max %.3e" % (world, k, np.abs(got[k] - ref[k]).max())
