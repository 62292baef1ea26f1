"""Synthetic inputs of the BASELINE.json configurations: the reference ships no porous geometry
(its drivers read `~/StructureImage/structure.png`), so the benchmark / full-size tests generate one."""
import numpy as np


def sphere_pack(shape, seed=7, porosity=0.6, rmin=6.0, rmax=14.0, buffer_planes=40):
    """cfg 5 generator: void mask `[nz, ny, nx]` (True = void) of a random pack of overlapping spheres, periodic in
    x and y, with `buffer_planes` void planes at the inlet (top of z) and the outlet (bottom of z).  Spheres are
    added in batches until the porosity of the core drops to the target."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    dom = np.ones(shape, bool)
    core = slice(buffer_planes, nz - buffer_planes)
    ncore = max(1, nz - 2 * buffer_planes)
    # expected number of spheres for the target porosity (Boolean model: porosity = exp(-n v / V)), added in batches
    vmean = 4.0 / 3.0 * np.pi * np.mean(np.linspace(rmin, rmax, 64) ** 3)
    batch = max(8, int(0.1 * -np.log(porosity) * ncore * ny * nx / vmean))
    while dom[core].mean() > porosity:
        for _ in range(batch):
            r = rng.uniform(rmin, rmax)
            cx, cy = rng.uniform(0, nx), rng.uniform(0, ny)
            cz = rng.uniform(buffer_planes + r, max(buffer_planes + r + 1e-9, nz - buffer_planes - r))
            z0, z1 = max(0, int(cz - r) - 1), min(nz, int(cz + r) + 2)
            ys = np.arange(int(cy - r) - 1, int(cy + r) + 2); xs = np.arange(int(cx - r) - 1, int(cx + r) + 2)
            zz = np.arange(z0, z1)
            d2 = (zz[:, None, None] - cz) ** 2 + (ys[None, :, None] - cy) ** 2 + (xs[None, None, :] - cx) ** 2
            sub = dom[z0:z1][:, ys % ny][:, :, xs % nx] & (d2 > r * r)
            dom[np.ix_(zz, ys % ny, xs % nx)] = sub
    return dom


def disc_pack(shape, seed=7, porosity=0.6, rmin=6.0, rmax=14.0, buffer_rows=40):
    """cfg 3 generator: union of discs `[ny, nx]` (True = void) until the core reaches the target porosity; void
    buffer rows at the inlet and the outlet"""
    rng = np.random.default_rng(seed)
    ny, nx = shape
    dom = np.ones(shape, bool)
    core = slice(buffer_rows, ny - buffer_rows)
    k = 0
    while True:
        if k % 25 == 0 and dom[core].mean() <= porosity:
            return dom
        k += 1
        r = rng.uniform(rmin, rmax); cx = rng.uniform(0, nx); cy = rng.uniform(buffer_rows + r, ny - buffer_rows - r)
        y0, y1 = max(0, int(cy - r) - 1), min(ny, int(cy + r) + 2)
        x0, x1 = max(0, int(cx - r) - 1), min(nx, int(cx + r) + 2)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dom[y0:y1, x0:x1] &= ((xx - cx) ** 2 + (yy - cy) ** 2) > r * r


def baseline_inputs_2d(k, scale=1.0):
    """void mask and initial region of the 2-D BASELINE configurations (1: droplet of fluid 0; 2: red band at the inlet;
    3: fluid 0 below the inlet band), shared by the engine set-up below and by bench.py's reference arm"""
    def ext(n):
        return max(1, int(round(n * scale)))
    if k == 1:
        n = ext(128)
        yy, xx = np.mgrid[0:n, 0:n]
        return np.ones((n, n), bool), (xx - n / 2) ** 2 + (yy - n / 2) ** 2 <= (20 * n / 128.0) ** 2
    if k == 2:
        n = ext(512)
        dom = np.ones((n, n), bool)
        b = max(2, int(round(10 * scale)))
        dom[b:-b, 0] = False; dom[b:-b, -1] = False
        return dom, np.indices((n, n))[0] >= n - 2 * b
    if k == 3:
        n = ext(1024)
        dom = disc_pack((n, n), buffer_rows=max(6, int(round(40 * scale))), rmin=max(2.0, 6.0 * min(1.0, 4 * scale)),
                        rmax=max(4.0, 14.0 * min(1.0, 4 * scale)))
        return dom, np.indices((n, n))[0] < n - max(4, int(round(10 * scale)))
    raise ValueError("2-D configurations are 1..3")


def baseline_config(k, scale=1.0, lib_path=None, device=0, flags=0):
    """Engine + initial state of BASELINE.json configuration `k` (1..5), lattice extents multiplied
    by `scale` (tests run them small).  -> (engine, number of void nodes, description).  Single slab; the slab-decomposed
    runs of configurations 4 / 5 go through bench.py (`--workload box | porous`)."""
    from . import _lib

    def ext(n, multiple=1):
        return max(multiple, int(round(n * scale / multiple)) * multiple)
    if k == 1:      # D2Q9 original Shan-Chen, 128 x 128 periodic droplet (IniFiles/shanchen2D.ini)
        dom, reg = baseline_inputs_2d(1, scale)
        n = dom.shape[0]
        eng = _lib.Engine(9, (n, n), model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, n_components=2, sc_tau=[1.0, 1.0],
                          sc_G=[0, 3.8, 0, 0, 3.8, 0], sc_Gsolid=[-0.4, 0.4], lib_path=lib_path, device=device, flags=flags)
        eng.set_geometry(dom)
        eng.init_equilibrium(np.where(reg, 1.0, 0.06), np.where(reg, 0.06, 1.0))
        return eng, float(dom.sum()), "cfg 1: D2Q9 original Shan-Chen, %d x %d periodic droplet" % (n, n)
    if k == 2:      # D2Q9 CSF colour-gradient MRT, 512 x 512 capillary intrusion (RKtwophasesetup2D.ini)
        dom, red = baseline_inputs_2d(2, scale)
        n = dom.shape[0]
        eng = _lib.Engine(9, (n, n), relax=_lib.RELAX_MRT, sigma=0.1, contact_angle_deg=60.0, wetting_type=2, beta=0.7,
                          delta=0.98, tauR=1.0, tauB=1.0, tau_type=2, inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE,
                          inlet_velocity=-1.0e-4, rhoBL=1.0, rhoRL=5e-8, lib_path=lib_path, device=device, flags=flags)
        eng.set_geometry(dom)
        eng.init_equilibrium(np.where(red, 1.0, 5e-8) * dom, np.where(red, 5e-8, 1.0) * dom)
        return eng, float(dom.sum()), "cfg 2: D2Q9 colour-gradient CSF MRT, %d x %d capillary intrusion, velocity inlet, pressure outlet" % (n, n)
    if k == 3:      # D2Q9 explicit-forcing Shan-Chen MRT, 1024 x 1024 porous drainage (efs2D.ini)
        dom, reg = baseline_inputs_2d(3, scale)
        n = dom.shape[0]
        eng = _lib.Engine(9, (n, n), model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, n_components=2, sc_tau=[1.0, 1.0],
                          sc_G=[0, 0.2, 0, 0, 0.2, 0], sc_Gsolid=[-0.14, 0.14], inlet=_lib.INLET_VELOCITY,
                          outlet=_lib.OUTLET_PRESSURE, sc_inlet_velocity=[0.0, -5.03e-4], sc_rho_out=[1.0, 0.02],
                          lib_path=lib_path, device=device, flags=flags)
        eng.set_geometry(dom)
        eng.init_equilibrium(np.where(reg, 1.0, 0.02) * dom, np.where(reg, 0.02, 1.0) * dom)
        return eng, float(dom.sum()), "cfg 3: D2Q9 explicit-forcing Shan-Chen MRT, %d x %d disc pack, velocity inlet, pressure outlet" % (n, n)
    if k == 4:      # D3Q19 colour-gradient MRT, 256^3 periodic spinodal decomposition (RKtwophasesetup3D.ini)
        n = ext(256, 32 if scale >= 0.125 else 1)
        shape = (n, n, n)
        eng = _lib.Engine(19, shape, relax=_lib.RELAX_MRT, sigma=0.1, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0, tau_type=2,
                          lib_path=lib_path, device=device, flags=flags)
        eng.set_geometry(np.ones(shape, np.uint8))
        eng.init_spinodal_device(0.01, 20260117)
        return eng, float(n) ** 3, "cfg 4: D3Q19 colour-gradient CSF MRT, %d^3 periodic spinodal box" % n
    if k == 5:      # D3Q19 colour-gradient MRT, 512 x 512 x 1024 porous drainage, velocity inlet + convective outlet
        nxy, nz = ext(512, 32 if scale >= 0.0625 else 1), ext(1024)
        buf = max(4, int(round(40 * min(1.0, 4 * scale))))
        dom = sphere_pack((nz, nxy, nxy), buffer_planes=buf, rmin=max(2.0, 6.0 * min(1.0, 4 * scale)),
                          rmax=max(4.0, 14.0 * min(1.0, 4 * scale)))
        red = (np.arange(nz) >= nz - max(3, int(0.75 * buf)))[:, None, None]
        eng = _lib.Engine(19, dom.shape, relax=_lib.RELAX_MRT, sigma=0.1, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0, tau_type=2,
                          wetting_type=2, contact_angle_deg=60.0, inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_CONVECTIVE,
                          inlet_velocity=-5.0e-4, lib_path=lib_path, device=device, flags=flags)
        eng.set_geometry(dom)
        eng.init_equilibrium(np.where(red, 1.0, 5e-8) * dom, np.where(red, 5e-8, 1.0) * dom)
        return eng, float(dom.sum()), "cfg 5: D3Q19 colour-gradient CSF MRT, %d x %d x %d sphere pack, velocity inlet, convective outlet" % (nxy, nxy, nz)
    raise ValueError("BASELINE.json has configurations 1..5")
