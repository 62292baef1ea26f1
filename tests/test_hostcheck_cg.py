"""CPU tier: the library's operator code (openlbmpm_b200/csrc), built for the host by tests/hostcheck
(test hook, g++ -DLBM_HOSTCHECK), driven through the same C ABI and compared with the reference's golden
vectors and the oracle.  The GPU tier (test_gpu_*.py) runs the same cases on liblbmpm.so."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostcheck"))
import build as hostcheck_build
import cases


@pytest.fixture(scope="module")
def lib():
    return hostcheck_build.build()


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_indexing_bit_exact(path, lib):
    cases.check_indexing_vs_gold(path, lib)


@pytest.mark.parametrize("path", cases.GOLD_CG2D, ids=[cases.gold_id(p) for p in cases.GOLD_CG2D])
def test_trajectory_vs_reference(path, lib):
    cases.check_trajectory_vs_gold(path, lib)


def test_chunked_steps_equal_single_steps(lib):
    cases.check_trajectory_vs_gold(cases.GOLD_CG2D[0], lib, chunk=7)


@pytest.mark.parametrize("relax", ["MRT", "SRT"])
def test_d3q19_periodic_vs_oracle(lib, relax):
    m, mo = cases.case_d3q19_periodic(lib, relax=relax)
    assert np.allclose(m, mo, rtol=0, atol=1e-9)


def test_d3q19_sphere_wetting_vs_oracle(lib):
    cases.case_d3q19_sphere(lib)


def test_d2q9_obstacles_vs_oracle(lib):
    cases.case_d2q9_random(lib, steps=8)


# the same cases through the general (reference-ordered) kernels instead of the factored fast path
@pytest.mark.parametrize("path", [p for p in cases.GOLD_CG2D if "channel" not in p],
                         ids=[cases.gold_id(p) for p in cases.GOLD_CG2D if "channel" not in p])
def test_trajectory_vs_reference_general_kernels(path, lib):
    cases.check_trajectory_vs_gold(path, lib, flags=1)


def test_d3q19_general_kernels_vs_oracle(lib):
    cases.case_d3q19_periodic(lib, flags=1)
    cases.case_d3q19_sphere(lib, flags=1)


def test_fast_and_general_paths_agree_on_mass(lib):
    m, mo = cases.case_d3q19_periodic(lib, n=(6, 8, 32), steps=5)
    assert np.allclose(m, mo, rtol=0, atol=1e-9)


def test_d3q19_solid_across_the_periodic_faces(lib):
    """a sphere centred on the box corner: wetting solids and their colour values live in the ghost planes"""
    cases.case_d3q19_sphere(lib, centre=[0.3, 0.2, 0.4])
    cases.case_d3q19_sphere(lib, centre=[0.3, 0.2, 0.4], flags=1)


@pytest.mark.parametrize("inlet,outlet", [("Neumann", "Convective"), ("Neumann", "Dirichlet"), ("Dirichlet", "Convective")])
def test_d3q19_open_boundaries_vs_oracle(lib, inlet, outlet):
    cases.case_d3q19_open_boundaries(lib, inlet, outlet)


# ---- open boundaries on the factored fast path: the treated planes are patched around the two passes; several steps
# per call keep the lattice in factored form between the snapshots (a download after every step would not) ----
CHANNELS = [p for p in cases.GOLD_CG2D if "channel" in p]


@pytest.mark.parametrize("path", CHANNELS, ids=[cases.gold_id(p) for p in CHANNELS])
@pytest.mark.parametrize("chunk", [2, 19])
def test_channel_trajectory_vs_reference_chunked(path, chunk, lib):
    cases.check_trajectory_vs_gold(path, lib, chunk=chunk)


def test_d3q19_pressure_inlet_pressure_outlet_vs_oracle(lib):
    cases.case_d3q19_open_boundaries(lib, "Dirichlet", "Dirichlet", relax="SRT", steps=12)


def test_open_boundary_fast_path_equals_general_kernels(lib):
    """same channel, factored fast path vs reference-ordered kernels: the trajectories agree to rounding"""
    import numpy as np
    out = []
    for flags in (0, 1):
        g, p = cases.load_gold(CHANNELS[-1])
        eng = cases.engine_for_gold(g, p, lib, flags=flags)
        eng.step(25)
        rho, u = eng.download_macros()
        out.append(np.stack(rho + u))
        eng.close()
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=cases.ATOL_GOLD)


def test_fast_wetting_form_of_the_tiled_kernels_equals_the_reference_ordered_form(lib):
    """cg_wetting_akai3_fast (rsqrt, sqrt(1 - dot^2), squared distances) vs cg_wetting<3> (acos / sin / cos, the
    reference's order of operations) on random gradients and solid normals, away from the degenerate alignment
    G || n_s where the reference's own choice between the two candidates is rounding noise"""
    import ctypes
    import numpy as np
    h = ctypes.CDLL(lib)
    rng = np.random.default_rng(12)
    n = 200000
    G = rng.uniform(-1, 1, (n, 3)) * rng.choice([1.0, 1e-3, 1e-6], (n, 1))
    ns = rng.uniform(-1, 1, (n, 3)); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    dot = -(G * ns).sum(1) / np.linalg.norm(G, axis=1)
    keep = np.abs(dot) < 1 - 1e-6
    G, ns = np.ascontiguousarray(G[keep]), np.ascontiguousarray(ns[keep])
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    h.hostcheck_wetting3.argtypes = [ctypes.POINTER(ctypes.c_double)] * 2 + [ctypes.c_double] * 2 + [ctypes.c_int, ctypes.c_int64,
                                                                                                  ctypes.POINTER(ctypes.c_double)]
    for theta in (20.0, 60.0, 90.0, 135.0):
        out = [np.empty_like(G), np.empty_like(G)]
        for fast in (0, 1):
            h.hostcheck_wetting3(p(G), p(ns), np.cos(np.radians(theta)), np.sin(np.radians(theta)), fast, len(G), p(out[fast]))
        rel = np.abs(out[0] - out[1]).max(axis=1) / np.linalg.norm(G, axis=1)
        assert rel.max() < 1e-10, (theta, rel.max())
        assert np.abs(np.linalg.norm(out[1], axis=1) / np.linalg.norm(G, axis=1) - 1).max() < 1e-9      # |G| is preserved
    # tiny gradients (|G| <= 1e-8) and exact alignment: no update
    Gs = np.array([[1e-9, 0, 0], [0.0, 0.0, 0.3], [0.0, 0.0, -0.3]]); nz = np.array([[0, 0, 1.0]] * 3)
    o = np.empty_like(Gs)
    h.hostcheck_wetting3(p(Gs), p(np.ascontiguousarray(nz)), 0.5, np.sqrt(0.75), 1, 3, p(o))
    assert np.array_equal(o, Gs)


# ---- perturbation surface-tension operator ----
@pytest.mark.parametrize("path", cases.GOLD_CGP2D, ids=[os.path.basename(p)[6:-4] for p in cases.GOLD_CGP2D])
@pytest.mark.parametrize("chunk,flags", [(1, 0), (7, 0), (7, 1)])
def test_perturbation_trajectory_vs_reference_kernels(path, chunk, flags, lib):
    """flags = 0: the factored fast path (density pass + PullPerturbCollideOp); 1: the reference-ordered kernels"""
    clean = cases.check_cgp_vs_gold(path, lib, chunk=chunk, flags=flags)
    # the centred droplet loses its conditioning when both colours meet at its antipode (cases.cgp_clean_snapshots);
    # the asymmetric cases are compared over all 40 snapshots
    assert clean == (24 if "block_srt" in path else 40) or "cgp2d_droplet.npz" in path


@pytest.mark.parametrize("lattice,n", [(19, (10, 12, 14)), (9, (14, 18))])
@pytest.mark.parametrize("solid", [False, True])
@pytest.mark.parametrize("flags", [0, 1])
def test_perturbation_vs_dense_oracle(lattice, n, solid, flags, lib):
    m, m_ref = cases.case_cgp_dense(lib, lattice, n, solid=solid, flags=flags)
    assert abs(m[0] - m_ref[0]) < 1e-9 and abs(m[1] - m_ref[1]) < 1e-9


def test_perturbation_rejects_what_the_reference_cannot_run(lib):
    from openlbmpm_b200 import _lib
    for bad in (dict(relax=_lib.RELAX_SRT, body_force=[1e-5, 0.0, 0.0]), dict(surface_tension_type=7)):
        with pytest.raises(_lib.LbmError):
            _lib.Engine(9, (8, 8), lib_path=lib, **dict(dict(surface_tension_type=_lib.ST_PERTURBATION), **bad))


def test_edge_cases(lib):
    cases.check_edge_cases(lib)


def test_perturbation_fields_download(lib):
    """lbm_download_fields with the perturbation operator: phi of the output point, G of the last collision, no K / F"""
    import numpy as np
    from oracle import cgp_dense
    dom = np.ones((12, 14), bool); dom[4:6, 5:8] = False
    rng = np.random.default_rng(0)
    r = 0.5 + 0.3 * (rng.random(dom.shape) - 0.5)
    par = dict(beta=0.8, AkR=1e-2, AkB=1e-2, tauR=1.0, tauB=0.9, solid_phi=0.3)
    eng = cases.cgp_engine(9, dom, lib, **par)
    eng.init_equilibrium(r * dom, (1 - r) * dom); eng.step(3)
    sim = cgp_dense.CGPDense(cgp_dense.d2q9(), dom, **par)
    sim.set_densities(r, 1 - r); sim.step(3)
    f = eng.download_fields()
    np.testing.assert_allclose(f["phi"], sim.phi[0], atol=1e-13)
    np.testing.assert_allclose(f["G"][0], sim.G[0, 0], atol=1e-13); np.testing.assert_allclose(f["G"][1], sim.G[1, 0], atol=1e-13)
    assert not f["K"].any() and not f["F"][0].any()
    eng.step(2); sim.step(2)                       # the download did not disturb the run
    np.testing.assert_allclose(eng.download_macros()[0][0], sim.rhoR[0], atol=1e-13)
    eng.close()


@pytest.mark.parametrize("k,scale", [(1, 0.25), (2, 0.0625), (3, 0.0625), (4, 0.046875), (5, 0.03125)])
def test_baseline_configurations_run_small(k, scale, lib):
    """synthetic.baseline_config: the five BASELINE.json configurations (what `bench.py --workload cfgN` times), shrunk"""
    from openlbmpm_b200 import synthetic
    import numpy as np
    eng, nodes, what = synthetic.baseline_config(k, scale=scale, lib_path=lib)
    m0 = eng.total_mass()
    eng.step(12)
    rho, u = eng.download_macros()
    m1 = eng.total_mass()
    assert nodes > 0 and ("cfg %d" % k) in what
    assert all(np.isfinite(a).all() for a in rho + u)
    if k in (1, 4):                       # closed boxes conserve every component
        np.testing.assert_allclose(m1, m0, rtol=1e-12)
    eng.step(3); eng.step(10)             # the second call finds the lattice in its steady form (no entry / exit work)
    # launches per step: the launch-bound 2-D configurations live on few, fat launches
    # (cfg 2 on the D2Q9 tile kernels: density tile | open rows as three parallel launches: materialise, row operators, head |
    # colour of the wetting solids (list) | collision tile with the gradient in shared memory, treated open rows included;
    # cfg 1 and cfg 3 on the two-pass form of the Shan-Chen loops:
    # pull-density | [open rows] | pull-collide, between one reference-ordered iteration at the start of a call (2 | 4 launches),
    # its first collision (1) and the streaming + rows + force at its end (1 | 3))
    assert eng.timing()["launches"] == {1: 2 + 1 + 8 * 2 + 1, 2: 60, 3: 4 + 1 + 8 * 3 + 3, 4: 30, 5: 70}[k]
    eng.close()


# ---- solute tracers riding on the colour-gradient CSF flow ----
@pytest.mark.parametrize("path", cases.GOLD_TR2D, ids=[os.path.basename(p)[5:-4] for p in cases.GOLD_TR2D])
@pytest.mark.parametrize("chunk,flags", [(1, 0), (9, 0), (9, 1)])
def test_tracer_trajectory_vs_reference_kernels(path, chunk, flags, lib):
    """flags = 0: the tracers ride on the factored fast path of the flow; 1: on the reference-ordered kernels"""
    cases.check_tracer_vs_gold(path, lib, chunk=chunk, flags=flags)


def test_tracers_ride_on_the_fast_path(lib):
    """closed box: the flow keeps its factored fast path with tracers attached (3 flow + 2 tracer launches per step instead of the
    reference-ordered sequence), same concentrations as on the reference-ordered kernels; downloads in between re-enter it"""
    import numpy as np
    from openlbmpm_b200 import _lib
    shape = (24, 12)
    dom = np.ones(shape, bool); dom[5:8, 3:7] = False
    rng = np.random.default_rng(9)
    rhoR = 0.5 + 0.3 * (rng.random(shape) - 0.5)
    conc = 0.2 + rng.random((2,) + shape)
    res = {}
    for flags in (0, _lib.FLAG_GENERIC_KERNELS):
        for schemes in (9, 5):
            e = _lib.Engine(9, shape, lib_path=lib, contact_angle_deg=70.0, flags=flags)
            e.tracer_setup(n_tracers=2, relax=_lib.RELAX_MRT, dxx=(0.05, 0.1), dyy=(0.08, 0.1), beta=(0.6, 0.3), n_schemes=schemes)
            e.set_geometry(dom)
            e.init_equilibrium(np.where(dom, rhoR, 0.0), np.where(dom, 1.0 - rhoR, 0.0))
            e.tracer_init(*[np.where(dom, c, 0.0) for c in conc])
            e.step(3); c1 = np.stack(e.tracer_download() + e.download_macros()[0])
            e.step(10); launches = e.timing()["launches"]
            res[(flags, schemes)] = (c1, np.stack(e.tracer_download() + e.download_macros()[0]), launches)
            e.close()
    for schemes in (9, 5):
        fast, gen = res[(0, schemes)], res[(_lib.FLAG_GENERIC_KERNELS, schemes)]
        np.testing.assert_allclose(fast[0], gen[0], rtol=0, atol=1e-13)
        np.testing.assert_allclose(fast[1], gen[1], rtol=0, atol=1e-13)
        assert fast[2] < gen[2] and fast[2] <= 50, (fast[2], gen[2])


@pytest.mark.parametrize("lattice,n,relax", [(9, (14, 18), "SRT"), (9, (14, 18), "MRT"), (19, (8, 10, 12), "SRT")])
@pytest.mark.parametrize("solid", [False, True])
def test_tracers_vs_dense_oracle(lattice, n, relax, solid, lib):
    m0, m1 = cases.case_tracer_dense(lib, lattice, n, relax=relax, solid=solid)
    import numpy as np
    np.testing.assert_allclose(m1, m0, rtol=1e-12)          # the tracer is conserved (half-way bounce back at the solids)


@pytest.mark.parametrize("channel", [False, True])
def test_five_velocity_tracers_vs_dense_oracle(channel, lib):
    """NumberSchemes = 5: MRT, interface term, reaction A + B -> C, Inamuro inlet row, free-flow outlet row; closed box and
    open channel of the flow, wetting solid"""
    cases.case_tracer_q5_dense(lib, channel=channel)


def test_tracer_setup_rejects_what_is_not_built(lib):
    from openlbmpm_b200 import _lib
    import numpy as np
    eng = _lib.Engine(9, (8, 8), lib_path=lib)
    for bad in (dict(n_schemes=7), dict(n_schemes=5, relax=_lib.RELAX_SRT), dict(n_schemes=5, relax=_lib.RELAX_MRT, reaction=True, n_tracers=2),
                dict(n_schemes=9, relax=_lib.RELAX_MRT, inlet_type=_lib.TR_INLET_DIRICHLET), dict(n_schemes=5, relax=_lib.RELAX_MRT, outlet_type=4)):
        with pytest.raises(_lib.LbmError):
            eng.tracer_setup(**bad)
    eng.close()
    eng = _lib.Engine(19, (6, 6, 6), lib_path=lib)
    with pytest.raises(_lib.LbmError):
        eng.tracer_setup(n_schemes=5, relax=_lib.RELAX_MRT)         # the 5-velocity lattice is 2-D
    eng.close()
    eng = _lib.Engine(19, (6, 6, 6), lib_path=lib)
    with pytest.raises(_lib.LbmError):
        eng.tracer_setup(relax=_lib.RELAX_MRT)                      # tracer MRT is D2Q9
    eng.close()
    eng = _lib.Engine(9, (8, 8), lib_path=lib, inlet=_lib.INLET_VELOCITY)
    with pytest.raises(_lib.LbmError):
        eng.tracer_setup()                                          # closed boxes only
    eng.close()
    eng = _lib.Engine(9, (8, 8), lib_path=lib)
    eng.set_geometry(np.ones((8, 8), bool)); eng.init_equilibrium(np.full((8, 8), 0.5), np.full((8, 8), 0.5))
    with pytest.raises(_lib.LbmError):
        eng.tracer_setup()                                          # must precede the flow state
    eng.close()


@pytest.mark.parametrize("lattice,n,inlet,outlet", [(19, (22, 8, 10), "Neumann", "Dirichlet"), (19, (22, 8, 10), "Dirichlet", "Convective"),
                                                    (9, (26, 14), "Neumann", "Convective"), (9, (26, 14), "Dirichlet", "Dirichlet")])
def test_perturbation_open_boundaries_vs_oracle(lattice, n, inlet, outlet, lib):
    """the perturbation operator with the open rows of the CSF loop: what the reference's 3-D ini parameterises"""
    cases.case_cgp_open(lib, lattice, n, inlet=inlet, outlet=outlet)


def test_perturbation_fast_path_launches_and_tiled_density(lib):
    """the perturbation model on the factored fast path: 2 launches per step on a closed 2-D box (3 on the reference-ordered
    kernels), and a D3Q19 box whose extents admit the tiled density pass (recolouring weights w_i / |e_i|) against the oracle"""
    import numpy as np
    from openlbmpm_b200 import _lib
    shape = (20, 24)
    dom = np.ones(shape, bool); dom[5:8, 3:7] = False
    rng = np.random.default_rng(4)
    r = 0.5 + 0.3 * (rng.random(shape) - 0.5)
    got = {}
    for flags in (0, 1):
        eng = cases.cgp_engine(9, dom, lib, 0.7, 8e-3, 1e-2, 1.0, 0.9, 0.4, flags=flags)
        eng.init_equilibrium(r * dom, (1 - r) * dom)
        eng.step(2); eng.step(10)
        got[flags] = (eng.timing()["launches"] / 10, np.stack(eng.download_macros()[0]))
        eng.close()
    assert got[0][0] == 2 and got[1][0] == 3, (got[0][0], got[1][0])
    np.testing.assert_allclose(got[0][1], got[1][1], rtol=0, atol=1e-12)
    cases.case_cgp_dense(lib, 19, (10, 16, 32), solid=True)
    cases.case_cgp_dense(lib, 19, (10, 16, 32), solid=False, relax="SRT", body_force=(0.0, 0.0, 0.0))


def test_reinitialisation_reuses_the_factored_buffers(lib):
    """lbm_init_equilibrium / lbm_upload_state on a lattice that is on the fast path keep its factored buffers (and, on slabs, the
    neighbours' mappings of them): the second run must equal a fresh engine's bit for bit; D2Q9 tile kernels, D3Q19 tiled, untiled"""
    import numpy as np
    from openlbmpm_b200 import _lib
    rng = np.random.default_rng(12)
    for lattice, shape in ((9, (16, 32)), (19, (10, 16, 32)), (19, (9, 6, 10))):
        dom = np.ones(shape, bool); dom[(slice(3, 6),) + (slice(2, 5),) * (len(shape) - 1)] = False
        a = 0.5 + 0.3 * (rng.random(shape) - 0.5); b = 0.5 + 0.3 * (rng.random(shape) - 0.5)
        eng = _lib.Engine(lattice, shape, lib_path=lib, contact_angle_deg=70.0)
        eng.set_geometry(dom)
        eng.init_equilibrium(a * dom, (1 - a) * dom); eng.step(5)
        eng.init_equilibrium(b * dom, (1 - b) * dom); eng.step(6)
        second = np.stack(eng.download_macros()[0])
        pdf = eng.download_pdfs()
        eng.upload_state(pdf); eng.step(3)
        third = np.stack(eng.download_macros()[0])
        eng.close()
        fresh = _lib.Engine(lattice, shape, lib_path=lib, contact_angle_deg=70.0)
        fresh.set_geometry(dom)
        fresh.init_equilibrium(b * dom, (1 - b) * dom); fresh.step(6)
        assert np.array_equal(second, np.stack(fresh.download_macros()[0])), (lattice, shape)
        fresh.close()
        fresh = _lib.Engine(lattice, shape, lib_path=lib, contact_angle_deg=70.0)      # (an upload starts with a zero lagged force)
        fresh.set_geometry(dom)
        fresh.upload_state(pdf); fresh.step(3)
        assert np.array_equal(third, np.stack(fresh.download_macros()[0])), (lattice, shape)
        fresh.close()


def test_open_rows_of_wide_planes_use_parallel_launches(lib):
    """planes of >= 16384 nodes: the open-row chain of the fast path runs as three fully parallel launches (materialise | row
    operators | head) instead of one thread per column -- same arithmetic, checked against the reference-ordered kernels"""
    import numpy as np
    from openlbmpm_b200 import _lib
    shape = (10, 128, 128)
    rng = np.random.default_rng(2)
    dom = np.ones(shape, bool); dom[4:6, 20:40, 30:60] = False
    top = np.indices(shape)[0] >= 7
    noise = rng.random(shape)
    kw = dict(contact_angle_deg=60.0, inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_CONVECTIVE, inlet_velocity=-1e-3)
    out = []
    for flags in (2, 1):          # one-thread-per-node fast path (the emulated tiled kernels would take minutes here) | general
        eng = _lib.Engine(19, shape, lib_path=lib, flags=flags, **kw)
        eng.set_geometry(dom)
        # (a generic colour field: a 5e-8 trace colour puts |G| next to the reference's 1e-8 threshold, and a flat interface over
        # the flat top of the block aligns G with the solid normal -- both are knife-edges of the reference's wetting kernel
        # where two equivalent arithmetic orders part ways at 1e-8, see tests/test_gpu_baseline_sizes.py)
        rR = (np.where(top, 0.8, 0.2) + 0.1 * (noise - 0.5)) * dom
        eng.init_equilibrium(rR, (1.0 - rR) * dom)
        eng.step(2); eng.step(4)
        rho, u = eng.download_macros()
        out.append(np.stack(rho + u))
        if flags == 2:
            assert eng.timing()["launches"] / 4 == 7       # density | materialise, row operators, head | gradient | collision | patched planes (5 with the single-launch chain)
        eng.close()
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=1e-10)
