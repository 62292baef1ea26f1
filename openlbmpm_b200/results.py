"""Result files with the reference's layout (PyTables groups /FluidMacro, /FluidPDF, /FluidVelocity and its
dataset names, RKD2Q9.py:348-356,938-957; ShanChenD2Q9.py:940-955).  `tables` / `h5py` are optional: without
them the same names are written to an .npz per output step.  Never on the timed path."""
import os

import numpy as np


def results_dir():
    d = os.environ.get("LBM_RESULTS_DIR", os.path.expanduser("~/LBMResults"))
    os.makedirs(d, exist_ok=True)
    return d


class ResultFile:
    def __init__(self, filename, groups=("FluidMacro", "FluidPDF", "FluidVelocity")):
        self.base = os.path.join(results_dir(), filename)
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

    def write(self, step, arrays):
        """arrays: {'/Group/Name': ndarray}"""
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
            np.savez_compressed("%s.step%08d.npz" % (os.path.splitext(self.base)[0], step),
                                **{p.strip("/").replace("/", "__"): a for p, a in arrays.items()})
