"""Result files with the reference's layout (PyTables groups /FluidMacro, /FluidPDF, /FluidVelocity and its
dataset names, RKD2Q9.py:348-356,938-957; ShanChenD2Q9.py:940-955).  `tables` / `h5py` are optional: without
them the same names are written to an .npz per output step.  Never on the timed path."""
import os

import numpy as np


def initial_dir():
    """where the drainage-imbibition restarts look for the previous run's results (RKD2Q9.py:491, ShanChenD2Q9.py:790)"""
    return os.environ.get("LBM_INITIAL_DIR", os.path.expanduser("~/LBMInitial"))


def read_arrays(directory, filename, paths):
    """{'/Group/Name': ndarray} for the given dataset paths of a result file written by ResultFile (any backend).
    npz backend: the datasets of one output step live in `<base>.stepNNNNNNNN.npz`; all step files are searched."""
    import glob
    base = os.path.join(directory, filename)
    out = {}
    if os.path.exists(base):
        try:
            import tables as tb
            with tb.open_file(base, "r") as f:
                for p in paths:
                    out[p] = f.get_node(p)[...]
            return out
        except ImportError:
            pass
        try:
            import h5py
            with h5py.File(base, "r") as f:
                for p in paths:
                    out[p] = f[p][...]
            return out
        except ImportError:
            pass
    stem = os.path.splitext(base)[0]
    files = sorted(glob.glob(stem + ".step*.npz")) + ([stem + ".npz"] if os.path.exists(stem + ".npz") else [])
    want = {p.strip("/").replace("/", "__"): p for p in paths}
    for fn in files:
        with np.load(fn) as z:
            for key, p in want.items():
                if key in z.files and p not in out:
                    out[p] = z[key]
    missing = [p for p in paths if p not in out]
    if missing:
        raise FileNotFoundError("datasets %s not found in %s (searched %s*)" % (missing, directory, os.path.basename(stem)))
    return out


def results_dir():
    d = os.environ.get("LBM_RESULTS_DIR", os.path.expanduser("~/LBMResults"))
    os.makedirs(d, exist_ok=True)
    return d


class ResultFile:
    def __init__(self, filename, groups=("FluidMacro", "FluidPDF", "FluidVelocity"), directory=None):
        self.base = os.path.join(directory or results_dir(), filename)
        self.groups = groups
        self.backend = None
        try:
            import tables  # noqa: F401
            self.backend = "tables"
        except Exception:
            try:
                import h5py  # noqa: F401
                self.backend = "h5py"
            except Exception:
                self.backend = "npz"
        if self.backend == "tables":
            import tables as tb
            f = tb.open_file(self.base, "w")
            for g in groups:
                f.create_group(f.root, g, g)
            f.close()
        elif self.backend == "h5py":
            import h5py
            with h5py.File(self.base, "w") as f:
                for g in groups:
                    f.create_group(g)

    def write(self, step, arrays, single_file=False):
        """arrays: {'/Group/Name': ndarray}; npz backend: one file per output step, or `<base>.npz` with single_file"""
        if self.backend == "tables":
            import tables as tb
            f = tb.open_file(self.base, "a")
            for path, a in arrays.items():
                grp, name = path.rsplit("/", 1)
                f.create_array(grp, name, a)
            f.close()
        elif self.backend == "h5py":
            import h5py
            with h5py.File(self.base, "a") as f:
                for path, a in arrays.items():
                    f[path] = a
        else:
            stem = os.path.splitext(self.base)[0]
            np.savez_compressed(stem + ".npz" if single_file else "%s.step%08d.npz" % (stem, step),
                                **{p.strip("/").replace("/", "__"): a for p, a in arrays.items()})


class AsyncMacroOutput:
    """Output that does not stall the step loop.  The reference blocks in six
    copy_to_host calls and an HDF5 append every TimeInterval steps (RKD2Q9.py:1382-1393); here a record is snapshot on
    the device (`lbm_download_macros_async`), copied into one of two page-locked host buffer sets on a second stream and
    written by a writer thread while the owner thread keeps queueing steps.  Records hold densities and velocities."""

    def __init__(self, engine, write_record, nsets=2):
        import threading
        self.engine, self.write_record = engine, write_record
        self.sets = [([engine.host_alloc() for _ in range(engine.ncomp)], [engine.host_alloc() for _ in range(engine.D)])
                     for _ in range(nsets)]
        self.writers = [None] * nsets
        self.lock = threading.Lock()
        self.records = 0
        self.errors = []

    def snapshot(self):
        """call between two `engine.step` calls: enqueues record number `self.records`"""
        import threading
        k = self.records % len(self.sets)
        if self.writers[k] is not None:
            self.writers[k].join()            # the buffer set is free once its previous record is on disk
        rho, u = self.sets[k]
        self.engine.download_macros_async(rho, u)
        rec = self.records

        def work():
            try:
                self.engine.output_wait()
                with self.lock:
                    self.write_record(rec, rho, u)
            except Exception as e:            # surfaced by close()
                self.errors.append(e)
        self.writers[k] = threading.Thread(target=work, daemon=True)
        self.writers[k].start()
        self.records += 1

    def close(self):
        for w in self.writers:
            if w is not None:
                w.join()
        if self.errors:
            raise self.errors[0]
