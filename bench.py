#!/usr/bin/env python
"""bench.py -- MLUPS of the D3Q19 colour-gradient MRT collision + streaming step on a periodic box.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)

One "step" = one lattice time step of the whole box (every void node collides, is recoloured and
streams once).  N = 1: the metric's own configuration, 512^3 periodic spinodal start (BASELINE.json
`metric`); N > 1: the same box slab-partitioned along z ("strong" scaling, ghost planes over NCCL).
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = {19: 2 * 19 * 8 * 2, 9: 2 * 9 * 8 * 2}     # one read + one write of every population
SPIN_AMP, SPIN_SEED = 0.01, 20260117


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def pinned_empty(shape):
    """host buffer in pinned memory (torch is only the allocator here)"""
    import torch
    t = torch.empty(shape, dtype=torch.float64, pin_memory=torch.cuda.is_available())
    return t, t.numpy()


def spinodal_host(shape, rank=0):
    rng = np.random.default_rng(SPIN_SEED + rank)
    return 0.5 + SPIN_AMP * (rng.random(shape) - 0.5)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on the host cores (bounded sample of the same workload)
# ---------------------------------------------------------------------------------------------------
def cpu_oracle(lattice, n, threads=None):
    """the oracle's C restatement on a periodic spinodal box of n^D nodes, densities set, one step run (first touch)"""
    from oracle import cg_c
    threads = threads or os.cpu_count()
    shape = (n, n, n) if lattice == 19 else (n, n)
    rhoR = spinodal_host(shape)
    sim = cg_c.CGC(lattice, np.ones(shape, bool), threads=threads)
    sim.set_densities(rhoR, 1.0 - rhoR)
    sim.step(1)
    return sim, rhoR.size, threads


def cpu_oracle_mlups(lattice, n, steps, threads=None):
    sim, nodes, threads = cpu_oracle(lattice, n, threads)
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    return nodes * steps / dt / 1e6, threads, dt


def reference_numba_cfg(cfg, scale, steps, warmup, target="cpu"):
    """BASELINE configurations 1..3 through the UNMODIFIED reference (its drivers and Numba kernels from baseline/_ref, the
    kernels re-targeted to numba.njit(parallel) on the host cores: oracle/ref_numba.py).  -> (MLUPS, ms per step, description)"""
    from oracle import ref_numba
    from openlbmpm_b200 import synthetic
    dom, reg = synthetic.baseline_inputs_2d(cfg, scale)
    ny, nx = dom.shape
    total = warmup + steps + 1
    if cfg == 2:
        out = ref_numba.run_cg2d(nx, ny, total, target=target, dom=dom, red=reg)
    elif cfg == 1:
        out = ref_numba.run_sc2d(nx, ny, total, target=target, model="ShanChen", dom=dom, region0=reg)
    else:
        out = ref_numba.run_sc2d(nx, ny, total, target=target, model="EFS", dom=dom, region0=reg,
                                 par=dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, inlet="Neumann", outlet="Dirichlet", vy1=-5.03e-4,
                                          bg0=0.02, bg1=0.02))
    dt = np.asarray(out["step_seconds"])[-steps:]          # the first iterations carry the JIT compilation
    ms = float(dt.mean()) * 1e3
    what = ("the reference's own driver (%s) and Numba kernels, unmodified, kernels re-targeted from cuda.jit to numba.njit(parallel) over "
            "all threads of each launch (oracle/ref_numba.py), %d x %d, %d timed iterations after %d" % (
                {1: "ShanChenD2Q9.runOptimizedLBM", 2: "RKColorGradientLBM.runRKColorGradient2DCSF", 3: "ShanChenD2Q9.runOptimizedEFLBM"}[cfg],
                nx, ny, len(dt), total - 1 - len(dt)))
    return out["n_fluid"] / (ms * 1e-3) / 1e6, ms, what, out["n_fluid"]


def run_reference(args):
    """CPU arm.  A "step" here is one time step of the SAMPLE box (--cpu-size, default 256^3: the 512^3 box of the metric
    needs ~120 GB in the reference's kernel-per-phase AoS layout and 10 s per step); W untimed steps, then exactly K timed
    ones, each timed on its own.  MLUPS is size-normalised; the sample that ran is named at the top level of the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in ("cfg1", "cfg2", "cfg3"):
        cfg = int(args.workload[3:])
        cores = os.cpu_count()
        v, ms, what, nodes = reference_numba_cfg(cfg, args.scale, max(1, args.steps), max(1, args.warmup))
        print(json.dumps({"impl": "reference", "metric": "MLUPS (void nodes) BASELINE config %d" % cfg, "value": v, "unit": "MLUPS",
                          "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(1, args.warmup), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "BASELINE configuration %d" % cfg, "sample": what, "void_nodes": nodes},
                          "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "reference", "sample": what},
                          "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    lattice = args.lattice
    n = args.cpu_size if lattice == 19 else args.size
    sim, nodes, cores = cpu_oracle(lattice, n)
    if args.warmup > 1:
        sim.step(args.warmup - 1)            # cpu_oracle() already ran one
    per_step = []
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        sim.step(1)
        per_step.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.sum(per_step)) / len(per_step)
    v = nodes / (ms * 1e-3) / 1e6
    dims = "x".join([str(n)] * (3 if lattice == 19 else 2))
    sample = ("%s periodic spinodal box (same generator, parameters and operator as the GPU arm; the lattice is smaller), "
              "%d timed steps after %d warm-up steps, oracle/cg_c: C + OpenMP restatement of the reference's kernel-per-phase "
              "loop on %d host threads" % (dims, len(per_step), max(1, args.warmup), cores))
    line = {"impl": "reference", "metric": metric_name(args, lattice, args.size, args.workload == "porous", 0),
            "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": len(per_step), "warmup": max(1, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "measured_size": dims, "extrapolated": n != args.size,
            "config": {"workload": workload_name(args), "sample": sample},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def metric_name(args, Q, n, porous, cfg):
    if cfg:
        return "MLUPS (void nodes) BASELINE config %d" % cfg
    if Q == 19 and n == 512 and not porous:
        return "MLUPS D3Q19 CG-MRT 512^3"
    if args.workload == "ini3d":
        return "MLUPS D3Q19 CG perturbation-operator SRT open channel %d^3" % n
    if porous:
        return "MLUPS (void nodes) D3Q19 CG-MRT porous"
    return "MLUPS %s CG-MRT %d" % ("D3Q19" if Q == 19 else "D2Q9", n)


def workload_name(args):
    if args.workload.startswith("cfg"):
        return getattr(args, "cfg_description", "BASELINE configuration %s" % args.workload[3:])
    if args.workload == "ini3d":
        return ("D3Q19 colour gradient with the perturbation operator, SRT, velocity inlet + pressure outlet (the variant the reference's "
                "IniFiles/RKtwophasesetup3D.ini selects), %d^3 all-fluid box" % args.size)
    if args.workload == "porous":
        return ("D3Q19 colour-gradient CSF MRT, %d x %d x %d sphere pack (seed 7, porosity ~0.6, half-way bounce back, contact "
                "angle 60), velocity inlet along -z, convective outlet (BASELINE config 5 geometry)" % (args.size, args.size, args.nz or args.size))
    if args.lattice == 19:
        return "D3Q19 colour-gradient CSF MRT, %d^3 periodic all-fluid box, spinodal start (BASELINE metric)" % args.size
    return "D2Q9 colour-gradient CSF MRT, %d^2 periodic all-fluid box, spinodal start" % args.size


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--lattice", type=int, default=19)
    ap.add_argument("--size", type=int, default=0, help="lattice extent (default: 512; ini3d: 256)")
    ap.add_argument("--general", action="store_true", help="force the general (unfused) kernels")
    ap.add_argument("--cpu-size", type=int, default=256)
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="lbm_config.flags (LBM_FLAG_*), for A/B measurements")
    ap.add_argument("--workload", default="box", choices=["box", "porous", "ini3d", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="box: the metric's periodic spinodal box; porous: sphere pack with velocity inlet + convective outlet (cfg 5 "
                         "geometry, any size, slab-decomposable); cfgN: BASELINE.json configuration N at its own size on one GPU "
                         "(--scale shrinks it)")
    ap.add_argument("--scale", type=float, default=1.0, help="cfgN: factor on the lattice extents")
    ap.add_argument("--nz", type=int, default=0, help="porous: planes along the flow axis (default: --size)")
    args = ap.parse_args()
    if args.workload in ("porous", "ini3d"):
        args.lattice = 19
    if not args.size:
        args.size = 256 if args.workload == "ini3d" else 512
    if args.impl == "reference":
        return run_reference(args)

    from openlbmpm_b200 import _lib
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); the CPU arm is --impl reference")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg = int(args.workload[3:]) if args.workload.startswith("cfg") else 0
    if cfg and world > 1:
        raise SystemExit("--workload cfgN runs on one GPU; the slab-decomposed workloads are box and porous")
    if cfg:
        args.lattice = 9 if cfg <= 3 else 19
    Q, n = args.lattice, args.size
    ini3d = args.workload == "ini3d"
    porous = args.workload == "porous" or cfg > 0 or ini3d        # no host-buffer round trip / CPU arm for these lines
    nz = (args.nz or n) if porous else n
    if nz % world:
        raise SystemExit("size must be divisible by the number of GPUs")
    nloc = nz // world
    shape = (nloc, n, n) if Q == 19 else (nloc, n)
    flags = (_lib.FLAG_GENERIC_KERNELS if args.general else 0) | args.flags
    bc = dict(inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_CONVECTIVE, inlet_velocity=-5.0e-4, contact_angle_deg=60.0) if porous else {}
    par = dict(relax=_lib.RELAX_MRT, sigma=0.1, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0, tau_type=2, wetting_type=2)
    if ini3d:       # what IniFiles/RKtwophasesetup3D.ini parameterises (its :9-25, :28-40, :55), on a box of --size^3 instead of 32 x 32 x 96
        bc = dict(inlet=_lib.INLET_VELOCITY, outlet=_lib.OUTLET_PRESSURE, inlet_velocity=-1.0e-4, rhoBL=1.0, rhoRL=1.0e-8)
        par = dict(relax=_lib.RELAX_SRT, surface_tension_type=_lib.ST_PERTURBATION, beta=1.0, AkR=7.0e-3, AkB=7.0e-3, tauR=1.0, tauB=1.0,
                   delta=0.98, solid_phi=0.7)
    if cfg:
        from openlbmpm_b200 import synthetic
        eng, nodes_total, args.cfg_description = synthetic.baseline_config(cfg, scale=args.scale, device=local, flags=flags)
        shape = eng.shape
    else:
        eng = _lib.Engine(Q, shape, model=_lib.MODEL_CG, device=local, flags=flags, **par, **bc)
    if world > 1:
        from openlbmpm_b200 import slab
        eng.comm_init(rank, world, slab.share_unique_id(dist, eng, rank, device="cuda"))
    if cfg:
        pass
    elif ini3d:
        sl = slice(rank * nloc, (rank + 1) * nloc)
        eng.set_geometry(np.ones(shape, np.uint8))
        red = np.broadcast_to((np.arange(nz)[sl] >= nz - nz // 8)[:, None, None], shape)      # the invading fluid fills the top eighth
        eng.init_equilibrium(np.where(red, 1.0, 1e-8), np.where(red, 1e-8, 1.0))
        nodes_total = float(n) * n * nz
    elif porous:
        from openlbmpm_b200 import synthetic
        sl = slice(rank * nloc, (rank + 1) * nloc)
        dom = synthetic.sphere_pack((nz, n, n))[sl]
        eng.set_geometry(dom)
        red = (np.arange(nz)[sl] >= nz - 30)[:, None, None]           # the invading fluid fills the inlet buffer
        eng.init_equilibrium(np.where(red, 1.0, 5e-8) * dom, np.where(red, 5e-8, 1.0) * dom)
        nodes_total = float(dom.sum())
        if dist is not None:
            t = torch.tensor([nodes_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            nodes_total = float(t.item())
        del dom
    else:
        eng.set_geometry(np.ones(shape, np.uint8))
        eng.init_spinodal_device(SPIN_AMP, SPIN_SEED)
        nodes_total = float(n) ** (3 if Q == 19 else 2)

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: K steps in one asynchronous lbm_step call -----------------
    eng.step(args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # Per-kernel CUDA events ride inside the timed region on the large lattices (two events per multi-millisecond kernel).
    # The launch-bound ones (<= 4 M nodes: every 2-D configuration) replay a captured CUDA graph, which per-launch events
    # would switch off: they are timed as the product runs them, and the kernel shares come from a second, profiled pass.
    graph_replay = world == 1 and float(np.prod(shape)) <= float(1 << 22)
    if not graph_replay:
        eng.profile(True)
    t0 = time.perf_counter()
    eng.step(args.steps)
    eng.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    tm = eng.timing()               # CUDA events on the handle's stream around the lbm_step call
    prof_steps = args.steps
    if graph_replay:
        barrier()
        prof_steps = min(args.steps, 200)
        eng.profile(True)
        eng.step(prof_steps)
        eng.synchronize()
    prof = eng.profile_report()
    eng.profile(False)
    barrier()
    clocks = sampler.summary() if sampler else None
    dev_ms = tm["ms"]
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = nodes_total / (ms_per_step * 1e-3) / 1e6
    # mass of each colour and the checksum of the densities over the WHOLE lattice: the slab sums add up (the checksum
    # mod 2^64) to the single-GPU values iff the decomposition is bit-equal to it
    mass = [float(m) for m in eng.total_mass()] 
    checksum = eng.checksum()
    if dist is not None:
        t = torch.tensor(mass, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        mass = [float(x) for x in t.tolist()]
        # uint64 wrap-around sum via two 32-bit halves in int64 (NCCL has no unsigned 64-bit sum in torch)
        t = torch.tensor([[c & 0xFFFFFFFF, c >> 32] for c in checksum], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        checksum = [int((lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF) for lo, hi in t.tolist()]

    # ---- end to end through the C ABI with host buffers: upload densities -> K steps -> download macros
    e2e = None
    if not args.no_e2e and not porous:
        keep = []
        t_r, rhoR = pinned_empty(shape); keep.append(t_r)
        t_b, rhoB = pinned_empty(shape); keep.append(t_b)
        rhoR[...] = spinodal_host(shape, rank); rhoB[...] = 1.0 - rhoR
        outs = [pinned_empty(shape) for _ in range(2 + eng.D)]
        out_rho = [o[1] for o in outs[:2]]; out_u = [o[1] for o in outs[2:]]
        barrier()
        t0 = time.perf_counter()
        eng.init_equilibrium(rhoR, rhoB)                 # H2D of this batch's input
        eng.step(args.steps)
        eng.download_macros(out_rho, out_u)              # D2H of the result (synchronises)
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": nodes_total * args.steps / e2e_s / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": 2 * rhoR.nbytes * world / args.steps,
               "d2h_bytes_per_step": (2 + eng.D) * rhoR.nbytes * world / args.steps,
               "batch": "init_equilibrium(host rhoR, rhoB) -> lbm_step(%d) -> download_macros(host), wall clock" % args.steps,
               "finite": bool(np.isfinite(out_rho[0]).all())}

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the step (all kernels of one time step; dominant kernel listed) ------------------
    peak, peak_src = measured_peaks()
    tot_prof = sum(ms for _, ms in prof.values()) or 1.0
    kernels = sorted(({"name": k, "launches": c, "ms_per_launch": ms / c, "share": ms / tot_prof if graph_replay else ms / (ms_per_step * args.steps)}
                      for k, (c, ms) in prof.items()), key=lambda r: -r["share"])
    step_kernel_ms = min(ms_per_step, tot_prof / prof_steps)
    # what the step spends outside this rank's own kernels: ghost-plane exchanges (NCCL groups or flag waits) and launch gaps
    comm_ms = max(0.0, ms_per_step - step_kernel_ms)
    alg_bytes = B_ALG[Q] * nodes_total / world           # per rank and step
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9    # the whole step as timed (max over ranks), exchanges included
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                "scope": "one whole time step per GPU as timed (%d launches/step, exchanges included), %d B per lattice update x %d "
                         "nodes per GPU" % (round(sum(c for c, _ in prof.values()) / prof_steps), B_ALG[Q], int(nodes_total / world)),
                "kernels_only_frac": alg_bytes / (step_kernel_ms * 1e-3) / 1e9 / peak,
                "kernels": kernels[:6]}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr) and Q == 19 and not args.general and not porous:
        try:       # dram__bytes_read.sum + dram__bytes_write.sum of the step's kernels (ncu --set full at 512^3), scaled to this slab
            t = json.load(open(tr))
            roofline["traffic"] = t["dram_bytes_per_step"] * (nodes_total / world) / t["nodes"]
            roofline["traffic_source"] = "profiles/traffic.json (ncu --set full, 512^3)"
        except Exception:
            pass

    cpu = None
    if not args.no_cpu and world == 1 and cfg in (1, 2, 3):
        try:       # the reference itself on the host cores, same configuration and size (bounded: a few dozen iterations)
            # cfg 3: the reference's Python set-up of a 1024^2 lattice takes minutes on its own, so its sample is the same
            # generator at half the extents (512^2); MLUPS is size-normalised and the sample is named in the line
            v, ms, what, _ = reference_numba_cfg(cfg, args.scale * (0.5 if cfg == 3 else 1.0), 20 if cfg != 1 else 200, 5)
            cpu = {"value": v, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "reference", "sample": what, "ms_per_step": ms}
        except Exception as e:
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: %r" % (e,)}
    if not args.no_cpu and world == 1 and not porous:
        try:
            v, cores, dt = cpu_oracle_mlups(Q, args.cpu_size, args.cpu_steps)
            cpu = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
                   "sample": "%d^%d periodic spinodal box x %d steps (%.1f s), oracle/cg_c: C + OpenMP restatement of the "
                             "reference's kernel-per-phase loop (the reference has no D3Q19 code; where it has code -- BASELINE configurations "
                             "1-3 -- the CPU leg is the reference itself, oracle/ref_numba.py)" % (args.cpu_size, 3 if Q == 19 else 2, args.cpu_steps, dt)}
        except Exception as e:       # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": "unavailable: %r" % (e,)}

    ws_gb = 2 * 2 * Q * 8 * nodes_total / world / 1e9
    line = {"metric": metric_name(args, Q, n, porous, cfg),
            "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args),
                       "l2": ("working set (%.2f GB) far larger than the 126 MB L2" if ws_gb > 0.5 else
                              "working set (%.3f GB) is of the order of the 126 MB L2: this configuration is cache-resident by its "
                              "own size, consecutive steps are timed without a flush") % ws_gb,
                       "path": "general kernels" if args.general else ("Shan-Chen operators" if cfg in (1, 3) else "fused fast path"), "slab": list(shape),
                       "void_nodes": nodes_total},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": tm["launches"],
            "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
            "comm_ms_per_step": comm_ms, "kernel_ms_per_step": step_kernel_ms,
            "mass": mass, "checksum": ["%016x" % c for c in checksum], "pct_hbm_roofline": 100.0 * achieved / peak}
    print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
