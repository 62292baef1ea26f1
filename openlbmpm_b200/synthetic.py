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
