"""ctypes binding of liblbmpm.so (include/lbmpm.h) -- the only bridge between the Python host code and
the sm_100a kernels.  There is no CPU fallback: if the shared library is missing or no CUDA device is
present, loading / `lbm_create` fails with an explicit error."""
import ctypes
import os

import numpy as np

ABI_VERSION = 1
OK = 0
MODEL_CG, MODEL_SC, MODEL_EFS = 0, 1, 2
RELAX_SRT, RELAX_MRT = 0, 1
BC_PERIODIC, INLET_VELOCITY, INLET_PRESSURE = 0, 1, 2
OUTLET_CONVECTIVE, OUTLET_PRESSURE = 1, 2
FLAG_GENERIC_KERNELS = 1
FLAG_GHOST_PLANES = 64
FLAG_PERSISTENT = 128
FLAG_PEER_EXCHANGE = 256
FLAG_NCCL_EXCHANGE = 512
ST_CSF, ST_PERTURBATION = 0, 1

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)


class LbmConfig(ctypes.Structure):
    """Mirror of `struct lbm_config` (include/lbmpm.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("lattice", ctypes.c_int32), ("model", ctypes.c_int32),
        ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("nz", ctypes.c_int32),
        ("relax", ctypes.c_int32), ("tau_type", ctypes.c_int32), ("wetting_type", ctypes.c_int32),
        ("inlet", ctypes.c_int32), ("outlet", ctypes.c_int32), ("device", ctypes.c_int32),
        ("flags", ctypes.c_uint32), ("n_components", ctypes.c_int32), ("sc_isotropy", ctypes.c_int32),
        ("surface_tension_type", ctypes.c_int32), ("reserved_i", ctypes.c_int32 * 1),
        ("sigma", ctypes.c_double), ("contact_angle_deg", ctypes.c_double), ("beta", ctypes.c_double),
        ("delta", ctypes.c_double), ("tauR", ctypes.c_double), ("tauB", ctypes.c_double),
        ("inlet_velocity", ctypes.c_double),
        ("rhoBH", ctypes.c_double), ("rhoRH", ctypes.c_double), ("rhoBL", ctypes.c_double), ("rhoRL", ctypes.c_double),
        ("sc_tau", ctypes.c_double * 4), ("sc_G", ctypes.c_double * 16), ("sc_Gsolid", ctypes.c_double * 4),
        ("sc_inlet_velocity", ctypes.c_double * 4), ("sc_rho_in", ctypes.c_double * 4),
        ("sc_rho_out", ctypes.c_double * 4),
        ("AkR", ctypes.c_double), ("AkB", ctypes.c_double), ("solid_phi", ctypes.c_double),
        ("body_force", ctypes.c_double * 3), ("reserved_d", ctypes.c_double * 2),
    ]


class LbmTracerConfig(ctypes.Structure):
    """Mirror of `struct lbm_tracer_config` (include/lbmpm.h)."""
    _fields_ = [("n_tracers", ctypes.c_int32), ("relax", ctypes.c_int32), ("tau", ctypes.c_double * 4),
                ("dxx", ctypes.c_double * 4), ("dyy", ctypes.c_double * 4), ("dxy", ctypes.c_double * 4),
                ("dyx", ctypes.c_double * 4), ("beta", ctypes.c_double * 4), ("criterion", ctypes.c_double),
                ("n_schemes", ctypes.c_int32), ("reaction", ctypes.c_int32), ("inlet_type", ctypes.c_int32),
                ("outlet_type", ctypes.c_int32), ("reaction_rate", ctypes.c_double), ("diff_j", ctypes.c_double * 4),
                ("inlet_conc", ctypes.c_double * 4)]


TR_NONE, TR_INLET_DIRICHLET, TR_OUTLET_FREEFLOW = 0, 1, 1


# name -> (restype, argtypes); every symbol include/lbmpm.h declares
PROTOTYPES = {
    "lbm_abi_version": (ctypes.c_int, []),
    "lbm_create": (ctypes.c_int, [ctypes.POINTER(LbmConfig), ctypes.POINTER(ctypes.c_void_p)]),
    "lbm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "lbm_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "lbm_set_geometry": (ctypes.c_int, [ctypes.c_void_p, c_uint8_p]),
    "lbm_index_sizes": (ctypes.c_int, [ctypes.c_void_p, c_int64_p, c_int64_p, c_int64_p]),
    "lbm_export_indexing": (ctypes.c_int, [ctypes.c_void_p, c_int64_p, c_int64_p, c_int64_p, c_int64_p,
                                           c_int64_p, c_int64_p, c_double_p]),
    "lbm_init_equilibrium": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32]),
    "lbm_upload_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.POINTER(c_double_p),
                                        ctypes.c_int32]),
    "lbm_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "lbm_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "lbm_download_macros": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32,
                                           ctypes.POINTER(c_double_p)]),
    "lbm_download_macros_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32,
                                                 ctypes.POINTER(c_double_p)]),
    "lbm_output_wait": (ctypes.c_int, [ctypes.c_void_p]),
    "lbm_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64]),
    "lbm_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "lbm_download_pdfs": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32]),
    "lbm_download_fields": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.POINTER(c_double_p),
                                           ctypes.POINTER(c_double_p), c_double_p]),
    "lbm_tracer_setup": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(LbmTracerConfig)]),
    "lbm_tracer_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32]),
    "lbm_tracer_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_double_p), ctypes.c_int32]),
    "lbm_total_mass": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int32]),
    "lbm_state_checksum": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32]),
    "lbm_get_timing": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_int64_p, c_int64_p]),
    "lbm_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "lbm_profile_report": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "lbm_init_spinodal_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_uint64]),
    "lbm_nccl_unique_id": (ctypes.c_int, [c_uint8_p]),
    "lbm_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_uint8_p]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblbmpm.so")
_cache = {}


class LbmError(RuntimeError):
    pass


def load(path=None):
    """Load the shared library and attach the prototypes.  `path` is for the test hook only."""
    path = path or LIB_PATH
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise LbmError("%s not found: build it with `python -m openlbmpm_b200.build` "
                       "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lbm_abi_version() != ABI_VERSION:
        raise LbmError("liblbmpm ABI version mismatch")
    _cache[path] = lib
    return lib


def _ptr_array(arrs, ctype=c_double_p):
    """array of pointers (NULL for None) from a list of numpy arrays"""
    out = (ctype * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data_as(ctype) if a is not None else ctype()
    return out


class Engine:
    """One lbm_handle: a lattice (one slab) on one GPU.  Thin, NumPy in / NumPy out."""

    def __init__(self, lattice, shape, model=MODEL_CG, relax=RELAX_MRT, lib_path=None, device=0, **params):
        self.lib = load(lib_path)
        self.Q = int(lattice)
        self.D = 2 if self.Q == 9 else 3
        shape = tuple(int(s) for s in shape)
        if self.D == 2:
            if len(shape) != 2:
                raise LbmError("D2Q9 needs shape (ny, nx)")
            nz, (ny, nx) = 1, shape
        else:
            if len(shape) != 3:
                raise LbmError("D3Q19 needs shape (nz, ny, nx)")
            nz, ny, nx = shape
        self.shape = shape
        cfg = LbmConfig()
        cfg.abi_version = ABI_VERSION
        cfg.lattice, cfg.model, cfg.relax = self.Q, model, relax
        cfg.nx, cfg.ny, cfg.nz = nx, ny, nz
        cfg.device = device
        cfg.tau_type, cfg.wetting_type = 2, 2
        cfg.sigma, cfg.contact_angle_deg, cfg.beta, cfg.delta = 0.1, 90.0, 0.7, 0.98
        cfg.tauR = cfg.tauB = 1.0
        cfg.n_components = 2
        for k, v in params.items():
            cur = getattr(cfg, k)           # AttributeError on unknown names
            if hasattr(cur, "__len__"):
                vals = list(np.asarray(v, dtype=float).ravel())
                for i, x in enumerate(vals):
                    cur[i] = x
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.ncomp = 2 if model == MODEL_CG else int(cfg.n_components)
        self._h = ctypes.c_void_p()
        rc = self.lib.lbm_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != OK:
            msg = self.lib.lbm_last_error(None)
            self._h = ctypes.c_void_p()
            raise LbmError("lbm_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != OK:
            msg = self.lib.lbm_last_error(self._h)
            raise LbmError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.lbm_output_wait(self._h)
            self.lib.lbm_destroy(self._h)
            self._h = ctypes.c_void_p()
        for p in getattr(self, "_pinned", []):
            self.lib.lbm_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dense(self, a, dtype=np.float64, extra=()):
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.shape != self.shape + tuple(extra):
            raise LbmError("array of shape %s expected, got %s" % (self.shape + tuple(extra), a.shape))
        return a

    # -- multi-GPU --------------------------------------------------------------------------
    def nccl_unique_id(self):
        buf = np.zeros(128, np.uint8)
        rc = self.lib.lbm_nccl_unique_id(buf.ctypes.data_as(c_uint8_p))
        if rc != OK:
            raise LbmError("lbm_nccl_unique_id failed (%d)" % rc)
        return buf

    def comm_init(self, rank, nranks, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._check(self.lib.lbm_comm_init(self._h, rank, nranks, uid.ctypes.data_as(c_uint8_p)), "lbm_comm_init")

    # -- geometry ---------------------------------------------------------------------------
    def set_geometry(self, is_domain):
        dom = self._dense(np.asarray(is_domain) != 0, np.uint8)
        self._check(self.lib.lbm_set_geometry(self._h, dom.ctypes.data_as(c_uint8_p)), "lbm_set_geometry")

    def index_sizes(self):
        n = [ctypes.c_int64() for _ in range(3)]
        self._check(self.lib.lbm_index_sizes(self._h, *[ctypes.byref(x) for x in n]), "lbm_index_sizes")
        return tuple(x.value for x in n)

    def export_indexing(self):
        nf, nw, nn = self.index_sizes()
        S = self.Q - 1
        out = dict(fluidNodes=np.empty(nf, np.int64), neighboringNodes=np.empty(nf * S, np.int64),
                   wettingSolidNodes=np.empty(nw, np.int64), neighboringWettingSolidNodes=np.empty(nw * S, np.int64),
                   fluidNodesWithSolidGPU=np.empty(nn, np.int64), fluidNodesWithSolidOriginal=np.empty(nn, np.int64))
        ns = np.empty((self.D, nn), np.float64)
        p = lambda a: a.ctypes.data_as(c_int64_p)
        self._check(self.lib.lbm_export_indexing(
            self._h, p(out["fluidNodes"]), p(out["neighboringNodes"]), p(out["wettingSolidNodes"]),
            p(out["neighboringWettingSolidNodes"]), p(out["fluidNodesWithSolidGPU"]),
            p(out["fluidNodesWithSolidOriginal"]), ns.ctypes.data_as(c_double_p)), "lbm_export_indexing")
        out["nsX"], out["nsY"] = ns[0], ns[1]
        if self.D == 3:
            out["nsZ"] = ns[2]
        return out

    # -- state ------------------------------------------------------------------------------
    def init_equilibrium(self, *rho):
        arrs = [self._dense(r) for r in rho]
        self._check(self.lib.lbm_init_equilibrium(self._h, _ptr_array(arrs), len(arrs)), "lbm_init_equilibrium")

    def init_spinodal_device(self, amplitude=0.01, seed=20260117):
        self._check(self.lib.lbm_init_spinodal_device(self._h, amplitude, seed), "lbm_init_spinodal_device")

    def upload_state(self, pdfs, rhos=None):
        p = [self._dense(a, extra=(self.Q,)) for a in pdfs]
        r = [None if a is None else self._dense(a) for a in (rhos or [None] * len(p))]
        self._check(self.lib.lbm_upload_state(self._h, _ptr_array(p), _ptr_array(r), len(p)), "lbm_upload_state")

    # -- hot path ---------------------------------------------------------------------------
    def step(self, n=1):
        self._check(self.lib.lbm_step(self._h, int(n)), "lbm_step")

    def synchronize(self):
        self._check(self.lib.lbm_synchronize(self._h), "lbm_synchronize")

    # -- results ----------------------------------------------------------------------------
    def download_macros(self, out_rho=None, out_u=None):
        """-> (list of densities, list of velocity components), dense arrays, zero on solid nodes"""
        rho = out_rho or [np.empty(self.shape) for _ in range(self.ncomp)]
        u = out_u or [np.empty(self.shape) for _ in range(self.D)]
        self._check(self.lib.lbm_download_macros(self._h, _ptr_array(rho), len(rho), _ptr_array(u + [None] * (3 - len(u)))),
                    "lbm_download_macros")
        return rho, u

    # -- solute tracers riding on the colour-gradient CSF flow ------------------------------------
    def tracer_setup(self, n_tracers=1, relax=RELAX_SRT, tau=(1.0,), dxx=(0.0,), dyy=(0.0,), dxy=(0.0,), dyx=(0.0,),
                     beta=(0.0,), criterion=0.5, n_schemes=9, reaction=False, reaction_rate=0.0, diff_j=(1. / 3.,),
                     inlet_type=TR_NONE, inlet_conc=(0.0,), outlet_type=TR_NONE):
        """before init_equilibrium / upload_state (the transport loop starts with the flow's streaming)"""
        cfg = LbmTracerConfig()
        cfg.n_tracers, cfg.relax, cfg.criterion = int(n_tracers), int(relax), float(criterion)
        cfg.n_schemes, cfg.reaction, cfg.reaction_rate = int(n_schemes), int(bool(reaction)), float(reaction_rate)
        cfg.inlet_type, cfg.outlet_type = int(inlet_type), int(outlet_type)
        for name, vals in (("tau", tau), ("dxx", dxx), ("dyy", dyy), ("dxy", dxy), ("dyx", dyx), ("beta", beta),
                           ("diff_j", diff_j), ("inlet_conc", inlet_conc)):
            arr = getattr(cfg, name)
            vals = list(np.asarray(vals, float).ravel())
            for i in range(4):
                arr[i] = vals[i] if i < len(vals) else vals[-1]
        self._check(self.lib.lbm_tracer_setup(self._h, ctypes.byref(cfg)), "lbm_tracer_setup")
        self.n_tracers = int(n_tracers)

    def tracer_init(self, *conc):
        arrs = [self._dense(c) for c in conc]
        self._check(self.lib.lbm_tracer_init(self._h, _ptr_array(arrs), len(arrs)), "lbm_tracer_init")

    def tracer_download(self):
        out = [np.empty(self.shape) for _ in range(self.n_tracers)]
        self._check(self.lib.lbm_tracer_download(self._h, _ptr_array(out), len(out)), "lbm_tracer_download")
        return out

    # -- asynchronous output (include/lbmpm.h: lbm_download_macros_async) ----------------------
    def host_alloc(self, shape=None):
        """float64 array of `shape` (default: the lattice) in page-locked host memory, owned by the engine"""
        shape = tuple(shape or self.shape)
        n = int(np.prod(shape))
        p = ctypes.c_void_p()
        rc = self.lib.lbm_host_alloc(ctypes.byref(p), n * 8)
        if rc != OK:
            raise LbmError("lbm_host_alloc of %d bytes failed (%d)" % (n * 8, rc))
        if not hasattr(self, "_pinned"):
            self._pinned = []
        self._pinned.append(p)
        return np.ctypeslib.as_array(ctypes.cast(p, c_double_p), shape=(n,)).reshape(shape)

    def download_macros_async(self, out_rho, out_u):
        """enqueue the copy of the output point's densities and velocity into the given (page-locked) arrays and return;
        `output_wait()` blocks until they have landed"""
        for a in list(out_rho) + list(out_u):
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != self.shape:
                raise LbmError("C-contiguous float64 arrays of shape %s expected" % (self.shape,))
        u = list(out_u) + [None] * (3 - len(out_u))
        self._check(self.lib.lbm_download_macros_async(self._h, _ptr_array(list(out_rho)), len(out_rho), _ptr_array(u)),
                    "lbm_download_macros_async")

    def output_wait(self):
        self._check(self.lib.lbm_output_wait(self._h), "lbm_output_wait")

    def download_pdfs(self):
        pdf = [np.empty(self.shape + (self.Q,)) for _ in range(self.ncomp)]
        self._check(self.lib.lbm_download_pdfs(self._h, _ptr_array(pdf), len(pdf)), "lbm_download_pdfs")
        return pdf

    def download_fields(self):
        phi = np.empty(self.shape); K = np.empty(self.shape)
        G = [np.empty(self.shape) for _ in range(self.D)]
        F = [np.empty(self.shape) for _ in range(self.D)]
        pad = [None] * (3 - self.D)
        self._check(self.lib.lbm_download_fields(self._h, phi.ctypes.data_as(c_double_p), _ptr_array(G + pad),
                                                 _ptr_array(F + pad), K.ctypes.data_as(c_double_p)), "lbm_download_fields")
        return dict(phi=phi, G=G, F=F, K=K)

    def total_mass(self):
        m = np.zeros(self.ncomp)
        self._check(self.lib.lbm_total_mass(self._h, m.ctypes.data_as(c_double_p), self.ncomp), "lbm_total_mass")
        return m

    def checksum(self):
        """per component: wrap-around uint64 sum of the bit patterns of rho at the output point (slab sums add up mod 2^64)"""
        s = (ctypes.c_uint64 * self.ncomp)()
        self._check(self.lib.lbm_state_checksum(self._h, s, self.ncomp), "lbm_state_checksum")
        return [int(x) for x in s]

    def profile(self, on):
        self._check(self.lib.lbm_profile_enable(self._h, 1 if on else 0), "lbm_profile_enable")

    def profile_report(self):
        """-> {kernel name: (launches, total ms)} measured with CUDA events on the launching stream"""
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self.lib.lbm_profile_report(self._h, buf, len(buf)), "lbm_profile_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split("\t")
            out[name] = (int(n), float(ms))
        return out

    def timing(self):
        ms = ctypes.c_double(); nl = ctypes.c_int64(); nn = ctypes.c_int64()
        self._check(self.lib.lbm_get_timing(self._h, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(nn)), "lbm_get_timing")
        return dict(ms=ms.value, launches=nl.value, nodes=nn.value)
