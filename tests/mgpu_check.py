"""Slab decomposition check, run as  torchrun --nproc-per-node P tests/mgpu_check.py  (one rank per GPU):
the P-slab run (ghost planes over NCCL) must reproduce the single-GPU run of the same box BIT FOR BIT
(pull streaming makes the halo arithmetic identical to the interior arithmetic).  Rank 0 prints 'MGPU OK'."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from openlbmpm_b200 import _lib, slab


def run(shape_global, dom, rhoR, steps, rank, world, tracers=None, **kw):
    nz = shape_global[0] // world
    sl = slice(rank * nz, (rank + 1) * nz)
    eng = _lib.Engine(19 if len(shape_global) == 3 else 9, (nz,) + tuple(shape_global[1:]), device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    if world > 1:
        eng.comm_init(rank, world, slab.share_unique_id(dist, eng, rank, device="cuda"))
    if tracers:
        eng.tracer_setup(**tracers)
    eng.set_geometry(dom[sl])
    eng.init_equilibrium(np.where(dom[sl], rhoR[sl], 0.0), np.where(dom[sl], 1.0 - rhoR[sl], 0.0))
    if tracers:
        eng.tracer_init(*[np.where(dom[sl], 0.3 + 0.5 * rhoR[sl] * (k + 1) / tracers["n_tracers"], 0.0) for k in range(tracers["n_tracers"])])
    eng.step(steps)
    rho, u = eng.download_macros()
    conc = eng.tracer_download() if tracers else []
    eng.close()
    return np.stack(rho + u + conc)


OPEN = dict(inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_CONVECTIVE, inlet_velocity=-2.0e-3)


def run_sc(shape_global, dom, rho0, steps, rank, world, **kw):
    """Shan-Chen models on slabs (2-D lattice cut along y)"""
    ny = shape_global[0] // world
    sl = slice(rank * ny, (rank + 1) * ny)
    eng = _lib.Engine(9, (ny,) + tuple(shape_global[1:]), device=int(os.environ.get("LOCAL_RANK", "0")), n_components=2,
                      sc_tau=[1.0, 0.9], sc_Gsolid=[-0.1, 0.1], **kw)
    if world > 1:
        eng.comm_init(rank, world, slab.share_unique_id(dist, eng, rank, device="cuda"))
    eng.set_geometry(dom[sl])
    eng.init_equilibrium(np.where(dom[sl], rho0[sl], 0.0), np.where(dom[sl], 1.1 - rho0[sl], 0.0))
    eng.step(steps)
    rho, u = eng.download_macros()
    eng.close()
    return np.stack(rho + u)


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(3)
    ok = True
    extra = int(os.environ.get("LBM_TEST_FLAGS", "0"))      # e.g. 32 = packed exchange, 16 = overlap (lbmpm.h)
    for name, shape, solid, kw in (("periodic tiled", (16 * world, 16, 32), False, {}),
                                   ("periodic tiled, 4 x-tiles", (40 * world, 16, 128), False, {}),      # interior tiles + two z-chunks per slab
                                   ("sphere wetting tiled", (16 * world, 16, 32), True, dict(contact_angle_deg=70.0)),
                                   ("D2Q9 tile kernels", (16 * world, 64), True, dict(contact_angle_deg=65.0)),
                                   ("D2Q9 tile kernels, open channel", (16 * world, 64), True, dict(OPEN, contact_angle_deg=60.0)),
                                   ("general kernels", (8 * world, 10, 12), True, dict(flags=1, contact_angle_deg=50.0)),
                                   ("untiled fast path", (8 * world, 10, 12), True, dict(flags=2)),
                                   # cfg 5 layout: outlet planes on rank 0, inlet planes on the last rank, solids in between
                                   ("open channel tiled", (16 * world, 16, 32), True, dict(OPEN, contact_angle_deg=60.0)),
                                   ("open channel general", (16 * world, 16, 32), True, dict(OPEN, flags=1)),
                                   ("perturbation operator SRT, open", (16 * world, 16, 32), True,
                                    dict(OPEN, surface_tension_type=_lib.ST_PERTURBATION, relax=_lib.RELAX_SRT, AkR=7e-3, AkB=7e-3, solid_phi=0.6)),
                                   ("tracers, 3-D SRT", (8 * world, 10, 12), True,
                                    dict(tracers=dict(n_tracers=2, relax=_lib.RELAX_SRT, tau=(0.8, 1.1), beta=(0.6, 0.3)))),
                                   ("tracers, 2-D 9-velocity MRT", (16 * world, 40), True,
                                    dict(tracers=dict(n_tracers=2, relax=_lib.RELAX_MRT, dxx=(0.05, 0.1), dyy=(0.08, 0.1), beta=(0.6, 0.3)))),
                                   ("tracers, 2-D 5-velocity in/out", (16 * world, 40), True,
                                    dict(tracers=dict(n_tracers=3, relax=_lib.RELAX_MRT, dxx=(0.05, 0.1, 0.07), dyy=(0.08, 0.1, 0.07),
                                                      beta=(0.6, 0.3, 0.0), n_schemes=5, reaction=True, reaction_rate=0.04,
                                                      diff_j=(0.3, 1. / 3., 0.4), inlet_type=_lib.TR_INLET_DIRICHLET,
                                                      inlet_conc=(0.7, 0.2, 0.0), outlet_type=_lib.TR_OUTLET_FREEFLOW)))):
        dom = np.ones(shape, bool)
        if solid and len(shape) == 2:
            yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
            dom = ((xx - 20) ** 2 + (yy - shape[0] / 2 + 0.5) ** 2) > 16.0
        elif solid:
            z, y, x = np.mgrid[0:shape[0], 0:shape[1], 0:shape[2]]
            dom = ((x - shape[2] / 2) ** 2 + (y - shape[1] / 2) ** 2 + (z - shape[0] / 2 + 0.5) ** 2) > 9.0
            if "inlet" not in kw:      # a second solid straddling the periodic seam (open channels keep their end planes void)
                dom &= ((x - 3) ** 2 + (y - 3) ** 2 + (z - 1) ** 2) > 4.0
        rhoR = 0.5 + 0.3 * (rng.random(shape) - 0.5)
        kw = dict(kw); kw["flags"] = kw.get("flags", 0) | extra
        mine = run(shape, dom, rhoR, 7, rank, world, **kw)
        gathered = [torch.zeros(mine.shape, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine).cuda())
        if rank == 0:
            full = np.concatenate([t.cpu().numpy() for t in gathered], axis=1)
            single = run(shape, dom, rhoR, 7, 0, 1, **kw)
            same = np.array_equal(full, single)
            print("%-32s P=%d bit-equal to P=1: %s  (max diff %.3e)" % (name, world, same, np.abs(full - single).max()), flush=True)
            ok &= same
        dist.barrier()
    for name, kw in (("Shan-Chen", dict(model=_lib.MODEL_SC, relax=_lib.RELAX_SRT, sc_G=[0, 0.9, 0, 0, 0.9, 0])),
                     ("explicit forcing MRT iso 8", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_MRT, sc_G=[0, 0.15, 0, 0, 0.15, 0], sc_isotropy=8)),
                     ("explicit forcing, open channel", dict(model=_lib.MODEL_EFS, relax=_lib.RELAX_SRT, sc_G=[0, 0.15, 0, 0, 0.15, 0],
                                                            inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE,
                                                            sc_inlet_velocity=[0.0, -5.0e-4], sc_rho_out=[1.0, 0.02]))):
        shape = (16 * world, 40)
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        dom = ((xx - 20) ** 2 + (yy - shape[0] / 2 + 0.5) ** 2) > 16.0
        rho0 = 0.6 + 0.3 * (rng.random(shape) - 0.5)
        mine = run_sc(shape, dom, rho0, 25, rank, world, **kw)          # 25 steps: the last 24 replay a CUDA graph on one slab
        gathered = [torch.zeros(mine.shape, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine).cuda())
        if rank == 0:
            full = np.concatenate([t.cpu().numpy() for t in gathered], axis=1)
            single = run_sc(shape, dom, rho0, 25, 0, 1, **kw)
            same = np.array_equal(full, single)
            print("%-32s P=%d bit-equal to P=1: %s  (max diff %.3e)" % (name, world, same, np.abs(full - single).max()), flush=True)
            ok &= same
        dist.barrier()
    if rank == 0:
        print("MGPU OK" if ok else "MGPU FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
