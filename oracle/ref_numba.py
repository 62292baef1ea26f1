"""CPU BASELINE / ORACLE INFRASTRUCTURE (test and bench infrastructure, NOT product code).

Runs the UNMODIFIED reference -- its driver classes (RKCG2D/RKD2Q9.py, ShanChen2D/ShanChenD2Q9.py) and its Numba-CUDA kernel
files (AcceleratedRKGPU2D.py, OptimizedD2Q9GPU.py, ExplicitD2Q9GPU.py, AccelerateGPU2D.py) -- with the kernels executed on one
of three targets:

  "cpu"   a stand-in for `numba.cuda` whose `jit` re-targets every kernel, at import time, from one-thread-per-node to
          `numba.njit(parallel=True)` over `prange(all threads of the launch)`: the kernel BODY is the reference's own source,
          untouched -- only `cuda.threadIdx / blockIdx / blockDim` become loop-derived locals, `cuda.shared/local.array` become
          per-thread arrays and `cuda.syncthreads()` disappears (no kernel of the live paths communicates through shared memory:
          they only cache constant tables there).  Same AoS layout, same kernel-per-phase order (the reference's own host loop,
          RKD2Q9.py:1295-1490 / ShanChenD2Q9.py:1492-1629), all host cores.  This is "the reference's own Numba path timed on the
          host cores" (BASELINE.md section 4.1).
  "cuda"  the real `numba.cuda` (the unmodified kernels on the GPU the process sees) -- BASELINE.md section 4.3's probe.
  "sim"   Numba's CUDA simulator (what tests/golden/gen_goldens_*.py use; interpreter speed).

The reference sources are NOT part of this repository: `vendor()` copies the handful of files from /root/reference into
baseline/_ref/ (git-ignored; it travels to the GPU box like the built libraries) when the reference tree is present, and every
entry point fails with a clear message when neither exists.  Only tests/, bench.py's reference arm and scripts/ import this.

Deviations from the reference's run (the same two the golden generators document, forced by undefined behaviour upstream):
`cuda.device_array_like` returns zeroed memory (the reference reads uninitialised force arrays in its first velocity
evaluation, RKD2Q9.py:1261-1262 vs :1362-1365), and on the cpu target a block-shared relaxation vector that every thread
overwrites with its own node's rates (AcceleratedRKGPU2D.py:1953,1982: a data race on the GPU) is private to the node.
"""
import ast
import builtins
import contextlib
import inspect
import io
import os
import shutil
import sys
import tempfile
import textwrap
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VENDOR_DIR = os.path.join(ROOT, "baseline", "_ref")
REFERENCE = os.environ.get("OPENLBMPM_REFERENCE", "/root/reference")
FILES = ["RKCG2D/RKD2Q9.py", "RKCG2D/AcceleratedRKGPU2D.py",
         "ShanChen2D/ShanChenD2Q9.py", "ShanChen2D/OptimizedD2Q9GPU.py", "ShanChen2D/ExplicitD2Q9GPU.py",
         "ShanChen2D/AccelerateGPU2D.py", "ShanChen2D/SimpleD2Q9.py", "ShanChen2D/SimpleGeometry.py"]


def vendor():
    """copy the reference files this module runs into baseline/_ref/ (no-op when the reference tree is absent)"""
    if not os.path.isdir(REFERENCE):
        return False
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(VENDOR_DIR, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    return True


def reference_root():
    for root in (REFERENCE, VENDOR_DIR):
        if os.path.exists(os.path.join(root, "RKCG2D", "RKD2Q9.py")):
            return root
    raise RuntimeError("the reference sources are neither at %s nor vendored under %s (run __graft_entry__.build() where the "
                       "reference tree is mounted)" % (REFERENCE, VENDOR_DIR))


# ---------------------------------------------------------------------------------------------------
# the cpu target: numba.cuda stand-in
# ---------------------------------------------------------------------------------------------------
_THREAD_NAMES = {("threadIdx", "x"): "_t_x", ("threadIdx", "y"): "_t_y", ("blockIdx", "x"): "_b_x", ("blockIdx", "y"): "_b_y",
                 ("blockDim", "x"): "_bd_x", ("blockDim", "y"): "_bd_y", ("gridDim", "x"): "_g_x", ("gridDim", "y"): "_g_y"}


class _Retarget(ast.NodeTransformer):
    """body of a @cuda.jit kernel -> body of one iteration of the thread loop"""

    def visit_Attribute(self, node):
        v = node.value
        if isinstance(v, ast.Attribute) and isinstance(v.value, ast.Name) and v.value.id == "cuda" and (v.attr, node.attr) in _THREAD_NAMES:
            return ast.copy_location(ast.Name(id=_THREAD_NAMES[(v.attr, node.attr)], ctx=ast.Load()), node)
        return self.generic_visit(node)

    def visit_Call(self, node):
        f = node.func
        if (isinstance(f, ast.Attribute) and f.attr == "array" and isinstance(f.value, ast.Attribute) and f.value.attr in ("shared", "local")
                and isinstance(f.value.value, ast.Name) and f.value.value.id == "cuda"):
            shape = node.args[0] if node.args else next(k.value for k in node.keywords if k.arg == "shape")
            dtype = node.args[1] if len(node.args) > 1 else next(k.value for k in node.keywords if k.arg == "dtype")
            name = dtype.id if isinstance(dtype, ast.Name) else dtype.attr
            new = ast.Call(func=ast.Attribute(value=ast.Name(id="_np_", ctx=ast.Load()), attr="empty", ctx=ast.Load()),
                           args=[shape, ast.Attribute(value=ast.Name(id="_np_", ctx=ast.Load()), attr=name, ctx=ast.Load())], keywords=[])
            return ast.copy_location(new, node)
        return self.generic_visit(node)

    def visit_Expr(self, node):
        c = node.value
        if (isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and c.func.attr == "syncthreads"):
            return ast.copy_location(ast.Pass(), node)
        if isinstance(c, ast.Constant) and isinstance(c.value, str):
            return ast.copy_location(ast.Pass(), node)        # stray docstrings between statements
        return self.generic_visit(node)

    def visit_Return(self, node):
        return ast.copy_location(ast.Continue(), node)


class CpuKernel:
    """what `@cuda.jit(...)` returns on the cpu target: `kernel[grid, block](*args)` runs every thread of the launch"""
    launches = {}          # name -> [count, seconds]  (all kernels; the runner reads it)

    def __init__(self, fn):
        import numba
        self.name = fn.__name__
        src = textwrap.dedent(inspect.getsource(fn))
        tree = ast.parse(src)
        fd = tree.body[0]
        fd.decorator_list = []
        body = [_Retarget().visit(s) for s in fd.body]
        idx = ast.parse(textwrap.dedent("""
            _t_x = _t_ % _bd_x
            _r1_ = _t_ // _bd_x
            _t_y = _r1_ % _bd_y
            _r2_ = _r1_ // _bd_y
            _b_x = _r2_ % _g_x
            _b_y = _r2_ // _g_x
        """)).body
        total = ast.parse("_g_x * _g_y * _bd_x * _bd_y", mode="eval").body
        loop = ast.For(target=ast.Name(id="_t_", ctx=ast.Store()),
                       iter=ast.Call(func=ast.Name(id="_prange_", ctx=ast.Load()), args=[total], keywords=[]),
                       body=idx + body, orelse=[])
        extra = [ast.arg(arg=a) for a in ("_g_x", "_g_y", "_bd_x", "_bd_y")]
        fd.args.args = extra + fd.args.args
        fd.body = [loop]
        fd.name = self.name + "__threads"
        ast.fix_missing_locations(tree)
        glob = fn.__globals__
        glob.setdefault("_prange_", numba.prange)
        glob.setdefault("_np_", np)
        code = compile(tree, "<retargeted %s>" % self.name, "exec")
        ns = {}
        exec(code, glob, ns)
        self.py = ns[fd.name]
        self.jitted = numba.njit(parallel=True, nogil=True)(self.py)

    def __getitem__(self, cfg):
        grid, block = cfg[0], cfg[1]
        g = tuple(grid) if isinstance(grid, (tuple, list)) else (int(grid),)
        b = tuple(block) if isinstance(block, (tuple, list)) else (int(block),)
        gx, gy = int(g[0]), int(g[1]) if len(g) > 1 else 1
        bx, by = int(b[0]), int(b[1]) if len(b) > 1 else 1

        def launch(*args):
            a = [np.asarray(x) if isinstance(x, np.ndarray) else x for x in args]
            t0 = time.perf_counter()
            self.jitted(gx, gy, bx, by, *a)
            rec = CpuKernel.launches.setdefault(self.name, [0, 0.0])
            rec[0] += 1; rec[1] += time.perf_counter() - t0
        return launch


class _DevArray(np.ndarray):
    def copy_to_host(self, ary=None):
        if ary is not None:
            ary[...] = self
            return ary
        return np.array(self, copy=True, subok=False)


def _fake_cuda():
    import numba
    m = types.ModuleType("numba.cuda (cpu target of oracle/ref_numba.py)")

    def jit(*args, **kw):
        def deco(fn):
            if kw.get("device"):
                return numba.njit(nogil=True)(fn)
            return CpuKernel(fn)
        if len(args) == 1 and callable(args[0]) and not isinstance(args[0], str):
            return deco(args[0])
        return deco
    m.jit = jit
    m.to_device = lambda a, *x, **k: np.array(a, copy=True).view(_DevArray)
    m.device_array_like = lambda a, *x, **k: np.zeros_like(np.asarray(a)).view(_DevArray)
    m.device_array = lambda shape, dtype=np.float64, *x, **k: np.zeros(shape, dtype).view(_DevArray)
    m.is_available = lambda: True
    m.gpus = "host cores (%d)" % (os.cpu_count() or 1)
    m.detect = lambda: True
    m.synchronize = lambda: None
    m.select_device = lambda *a: None
    m.close = lambda: None
    return m


def _identity_decorator(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda f: f


_loaded = {}


def load(target="cpu"):
    """import the reference modules for `target`; -> namespace(RKD2Q9, RK, SC, cuda).  One target per process."""
    if _loaded:
        if target not in _loaded:
            raise RuntimeError("the reference modules are already loaded for target %r" % (list(_loaded)[0],))
        return _loaded[target]
    root = reference_root()
    if target == "sim":
        os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    import numba
    import scipy as sp
    if target == "cpu":
        cuda = _fake_cuda()
        real = sys.modules["numba"]
        fake = types.ModuleType("numba")
        fake.__dict__.update({k: getattr(real, k) for k in ("float64", "int64", "float32", "int32", "njit", "prange", "vectorize")})
        fake.cuda = cuda
        fake.jit = fake.autojit = _identity_decorator
        fake.__path__ = []
        sys.modules["numba"] = fake
        sys.modules["numba.cuda"] = cuda
    else:
        from numba import cuda
        numba.autojit = _identity_decorator
        numba.jit = _identity_decorator
    for name in ("arange", "empty", "sqrt", "zeros", "ones", "array"):
        if not hasattr(sp, name):
            setattr(sp, name, getattr(np, name))
    if not hasattr(np, "bool"):
        np.bool = np.bool_
    saved = {}
    for modname in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "tables"):
        saved[modname] = sys.modules.get(modname)
        m = types.ModuleType(modname)
        m.use = lambda *a, **k: None
        sys.modules[modname] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    if "SimpleGeometryRK" not in sys.modules:      # the module RKD2Q9.py imports is not in the reference tree; the runner sets defineGeometry
        m = types.ModuleType("SimpleGeometryRK")
        m.defineGeometry = lambda x, y: (np.ones((y, x), bool), np.zeros((y, x), bool))
        sys.modules["SimpleGeometryRK"] = m
    for sub in ("RKCG2D", "ShanChen2D"):
        sys.path.insert(0, os.path.join(root, sub))
    builtins_input = builtins.input
    builtins.input = lambda *a, **k: ""
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import RKD2Q9
            import AcceleratedRKGPU2D as RK
            import ShanChenD2Q9 as SC
    finally:
        builtins.input = builtins_input
        if target == "cpu":
            sys.modules["numba"] = real
            sys.modules.pop("numba.cuda", None) if not hasattr(real, "cuda") else sys.modules.__setitem__("numba.cuda", real.cuda) if "cuda" in real.__dict__ else sys.modules.pop("numba.cuda", None)
        for sub in ("RKCG2D", "ShanChen2D"):
            sys.path.remove(os.path.join(root, sub))
    if target != "cpu":
        def zeroed(a):
            return cuda.to_device(np.zeros_like(np.asarray(a.copy_to_host() if hasattr(a, "copy_to_host") else a)))
        shim = types.SimpleNamespace(**{k: getattr(cuda, k) for k in ("to_device", "is_available")}, device_array_like=zeroed,
                                     gpus="", detect=lambda: True)
        RKD2Q9.cuda = shim
        SC.cuda = shim
    ns = types.SimpleNamespace(RKD2Q9=RKD2Q9, RK=RK, SC=SC, cuda=cuda, target=target, root=root)
    _loaded[target] = ns
    return ns


# ---------------------------------------------------------------------------------------------------
# runners: BASELINE configurations 2 (colour gradient CSF) and 1 (original Shan-Chen) through the reference's own drivers
# ---------------------------------------------------------------------------------------------------
CG_INI = """
[ImageSetup]
Existance = 'no'
[DomainSize]
xDomain = {nx}
yDomain = {ny}
numBufferingLayers = 2
ratioTopToBottom = 0.5
[SurfaceTension]
SurfaceTensionType = 'CSF'
SurfaceTensionValue = {sigma}
ContactAngle = {theta}
WettingType = {wetting}
[RKParameters]
AlphaR = 0.44444444
AlphaB = 0.44444444
BetaThickness = {beta}
AkR = 1.4e-1
AkB = 1.4e-1
DeltaValue = {delta}
[FluidParameters]
TauR = {tauR}
TauB = {tauB}
InitialRhoR = 1.0
InitialRhoB = 1.0
TauType = {tautype}
[BodyForce]
isBodyForce = 'no'
bodyForceX = 0.0
bodyForceY = 0.0
[SolidBoundarySetup]
SolidColorDiff = 0.5
[BoundaryCondition]
BoundaryTypeInlet = '{inlet}'
NeumannType = 'ZouHe'
VelocityYR = {vyr}
VelocityYB = 0.0
densityBH = 5e-8
densityRH = 1.00536
BoundaryTypeOutlet = '{outlet}'
densityBL = 1.0
densityRL = 5e-8
[GradientType]
Type = 'Isotropic'
[TimeSetup]
TimeSteps = {steps}
TimeInterval = {interval}
[Parallelism]
Parallel = 'yes'
xDimension = {xdim}
ThreadsNum = 32
[RelaxationType]
Type = '{relax}'
[CyclesSetup]
IsCycle = 'no'
LastStep = 100
"""


def run_cg2d(nx, ny, steps, target="cpu", dom=None, red=None, minor=5e-8, interval=None, par=None, warm_steps=0):
    """BASELINE configuration 2 through the reference's own driver `runRKColorGradient2DCSF`.
    -> dict(rhoR, rhoB, ux, uy at the last output, step_seconds (per loop iteration, from the launches of the per-step kernel
    `calRecoloringProcessM`), kernel table)"""
    ns = load(target)
    RKD2Q9 = ns.RKD2Q9
    p = dict(sigma=0.1, theta=60.0, wetting=2, beta=0.7, delta=0.98, tauR=1.0, tauB=1.0, tautype=2, inlet="Neumann",
             outlet="Dirichlet", vyr=-1.0e-4, relax="MRT", nx=nx, ny=ny, steps=steps, interval=interval or steps,
             xdim=128 if nx * ny >= 128 * 32 else 32)
    p.update(par or {})
    if dom is None:            # SimpleGeometry.py:18-19 scaled: side walls between the buffer rows
        dom = np.ones((ny, nx), bool)
        dom[10:-10, 0] = False; dom[10:-10, -1] = False
    if red is None:
        red = np.indices((ny, nx))[0] >= ny - 20
    RKD2Q9.defineGeometry = lambda x, y: (dom.copy(), ~dom)
    stamps = []

    class Ref(RKD2Q9.RKColorGradientLBM):
        def _RKColorGradientLBM__createHDF5File(self):
            pass

        def _RKColorGradientLBM__checkGPUAvailability(self):
            pass

        def plotDensityDistributionOPT(self, iStep):
            pass

        def resultInHDF5(self, iStep):
            self.snapshots.append(dict(rhoR=self.fluidsRhoR.copy(), rhoB=self.fluidsRhoB.copy(), ux=self.physicalVX.copy(),
                                       uy=self.physicalVY.copy()))

        def initializeDomainCondition(self):
            # the reference's allocation (RKD2Q9.py:451-456) and equilibrium fill (:561-585) with the colour layout of the case
            n_y, n_x = self.yDomain, self.xDomain
            w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
            self.fluidsRhoR = np.where(red, self.initialRhoR, minor) * self.isDomain
            self.fluidsRhoB = np.where(red, minor, self.initialRhoB) * self.isDomain
            self.fluidPDFR = self.fluidsRhoR[:, :, None] * w
            self.fluidPDFB = self.fluidsRhoB[:, :, None] * w
            self.physicalVX = np.zeros([n_y, n_x]); self.physicalVY = np.zeros([n_y, n_x])

    # per-iteration time stamps: wrap the kernel the loop launches exactly once per iteration
    rk = ns.RK
    orig = rk.calRecoloringProcessM

    class Stamp:
        def __getitem__(self, cfg):
            inner = orig[cfg]

            def run(*a):
                inner(*a)
                if ns.target == "cuda":
                    ns.cuda.synchronize()
                stamps.append(time.perf_counter())
            return run
    rk.calRecoloringProcessM = Stamp()
    RKD2Q9.RKGPU2D = rk
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "RKtwophasesetup2D.ini"), "w") as fh:
        fh.write(CG_INI.format(**p))
    t_all = time.perf_counter()
    sink = io.StringIO()
    keep_input = builtins.input
    builtins.input = lambda *a, **k: ""          # the drivers pause for a key press (RKD2Q9.py:174)
    try:
        with contextlib.redirect_stdout(sink):
            sim = Ref(tmp)
            sim.snapshots = []
            t_setup = time.perf_counter() - t_all
            sim.runRKColorGradient2DCSF()
    finally:
        builtins.input = keep_input
        rk.calRecoloringProcessM = orig
        shutil.rmtree(tmp, ignore_errors=True)
    dt = np.diff(np.array(stamps))
    return dict(snapshots=sim.snapshots, step_seconds=dt, setup_seconds=t_setup, total_seconds=time.perf_counter() - t_all,
                n_fluid=int(sim.fluidNodes.size), kernels={k: tuple(v) for k, v in CpuKernel.launches.items()}, sim=sim)


SC_BASIC = """
[Scheme]
Type = 'SRT'
[Geometry]
length = 1.0
width = 1.0
nx = {nx}
ny = {ny}
[Time]
TimeLength = 1.0
TimeStep = 1.0
[InitialCondition]
VelocityXLB = 0.0
VelocityYLB = 0.0
[BodyForce]
gValue = 0.0
[FlowDomain]
xDomain = 0,{nx}
yDomain = 0,{ny}
"""
SC_TWOPHASE = """
[PictureSetup]
Exist = 'no'
[SeparationBorder]
xGrid = {nx}
yGrid = {ny}
[FluidsTypes]
NumberOfFluids = 2
[InterType]
InteractionType = '{model}'
[Parallelism]
Parallel = 'yes'
xDimension = {xdim}
ThreadsNum = 32
[RelaxationType]
Type = '{relax}'
[DuplicateDomain]
Option = 'no'
[DICycles]
Option = 'no'
"""
SC_MODEL = """
[FluidProperties]
InitialDensities = {rho0},{rho1}
BackgroundDensities = {bg0},{bg1}
FluidsTau = {tau0},{tau1}
[{section}]
InteractionFluid = {G}
InteractionSolid = {Gs0},{Gs1}
[ForceScheme]
ExplicitScheme = {scheme}
[BoundaryDefinition]
BoundaryTypeInlet = '{inlet}'
BoundaryMethod = 'ZouHe'
BoundaryTypeOutlet = '{outlet}'
[VelocityBoundary]
velocityX = 0.0,0.0
velocityY = {vy0},{vy1}
[PressureBoundary]
PressureInlet = 1.0,0.06
PressureOutlet = 1.0,0.06
[BodyForce]
Option = 'no'
[Time]
numberTimeStep = {steps}
"""


def run_sc2d(nx, ny, steps, target="cpu", model="ShanChen", dom=None, region0=None, par=None, keep_states=False):
    """BASELINE configurations 1 (model 'ShanChen': `runOptimizedLBM`, ShanChenD2Q9.py:1433-1629) and 3 (model 'EFS':
    `runOptimizedEFLBM`, :1631-2087) through the reference's own driver.  The loops run `steps + 1` iterations.
    -> dict(step_seconds (per iteration, from the launches of `calPhysicalVelocity`, the last kernel of an iteration), states)"""
    ns = load(target)
    SC = ns.SC
    p = dict(relax="SRT", rho0=1.0, rho1=1.0, bg0=0.06, bg1=0.06, tau0=1.0, tau1=1.0, G=3.8, Gs0=-0.4, Gs1=0.4, inlet="Periodic",
             outlet="Periodic", vy0=0.0, vy1=-1.0e-3, steps=steps, scheme=4, nx=nx, ny=ny, model=model,
             xdim=128 if nx * ny >= 128 * 32 else 32)
    p.update(par or {})
    p["section"] = "ShanChenParameters" if model == "ShanChen" else "EFSParameters"
    if dom is None:
        dom = np.ones((ny, nx), bool)
    if region0 is None:        # shanchen2D.ini's droplet of fluid 0
        yy, xx = np.mgrid[0:ny, 0:nx]
        region0 = (xx - nx // 2) ** 2 + (yy - ny // 2) ** 2 <= (min(nx, ny) * 20 // 128) ** 2
    SC.defineGeometry = lambda x, y: (dom.copy(), ~dom)

    class Ref(SC.ShanChenD2Q9):
        def _ShanChenD2Q9__createHDF5File(self):
            pass

        def plotDensityDistributionOPT(self, *a, **k):
            pass

        def plotPhysicalVelocity(self, *a, **k):
            pass

        def resultInHDF5(self, *a, **k):
            pass

        def initializeDomainCondition(self):
            # the reference's allocation (ShanChenD2Q9.py:740-745) and rest-equilibrium fill (:759-768) with the case's layout
            n_y, n_x = self.ny, self.nx
            d = np.stack([np.where(region0, self.initialDensities[0], self.backgroundDensities[0]),
                          np.where(region0, self.backgroundDensities[1], self.initialDensities[1])]) * self.isDomain
            self.fluidsDensity = d.astype(np.float64)
            self.fluidPDF = self.fluidsDensity[..., None] * np.asarray(self.weightsCoeff)
            self.physicalVX = np.zeros([n_y, n_x]); self.physicalVY = np.zeros([n_y, n_x])
            self.forceX = np.zeros([self.typesFluids, n_y, n_x]); self.forceY = np.zeros([self.typesFluids, n_y, n_x])

    stamps, states = [], []
    real = SC.calPhysicalVelocity

    class Tap:
        def __getitem__(self, cfg):
            inner = real[cfg]

            def run(*a):
                inner(*a)
                if ns.target == "cuda":
                    ns.cuda.synchronize()
                stamps.append(time.perf_counter())
                if keep_states:      # calPhysicalVelocity(totalNodes, numFluids, xDim, fluidPDF, fluidRho, forceX, forceY, physVX, physVY)
                    states.append(dict(rho=a[4].copy_to_host(), ux=a[7].copy_to_host(), uy=a[8].copy_to_host()))
            return run
    SC.calPhysicalVelocity = Tap()
    tmp = tempfile.mkdtemp()
    for fname, text in (("basicsetup.ini", SC_BASIC), ("twophasesetup.ini", SC_TWOPHASE), ("shanchen2D.ini", SC_MODEL), ("efs2D.ini", SC_MODEL)):
        with open(os.path.join(tmp, fname), "w") as fh:
            fh.write(text.format(**p))
    keep_input = builtins.input
    builtins.input = lambda *a, **k: ""
    t_all = time.perf_counter()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim = Ref(tmp)
            t_setup = time.perf_counter() - t_all
            if model == "ShanChen":
                sim.runOptimizedLBM()
            else:
                sim.runOptimizedEFLBM()
    finally:
        builtins.input = keep_input
        SC.calPhysicalVelocity = real
        shutil.rmtree(tmp, ignore_errors=True)
    if model == "EFS":
        stamps = stamps[1::2]; states = states[1::2]      # two launches per iteration (:1902 and :2016)
    return dict(step_seconds=np.diff(np.array(stamps)), states=states, setup_seconds=t_setup, n_fluid=int(sim.fluidNodes.size),
                total_seconds=time.perf_counter() - t_all, sim=sim, kernels={k: tuple(v) for k, v in CpuKernel.launches.items()})
