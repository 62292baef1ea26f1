"""D3Q19 Shan-Chen models (`main.py 3D flow SC`): two-pass form (csrc/sc_fast.cuh) against the reference-ordered operators on one
GPU -- MLUPS of `steps` iterations in one lbm_step call (CUDA events of lbm_get_timing), and the largest difference between the two."""
import json
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/scripts", 1)[0])
from openlbmpm_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
shape = (n, n, n)
rng = np.random.default_rng(1)
r0 = 0.6 + 0.3 * (rng.random(shape) - 0.5)
dom = np.ones(shape, bool)
zz, yy, xx = np.indices(shape)
dom &= (zz - n / 2) ** 2 + (yy - n / 2) ** 2 + (xx - n / 2) ** 2 > (n / 5) ** 2        # a solid sphere
for model, relax, G in ((_lib.MODEL_SC, _lib.RELAX_SRT, 0.9), (_lib.MODEL_EFS, _lib.RELAX_MRT, 0.15)):
    out = {}
    for name, flags in (("two_pass", 0), ("reference_ordered", _lib.FLAG_GENERIC_KERNELS)):
        e = _lib.Engine(19, shape, model=model, relax=relax, n_components=2, sc_tau=[1.0, 0.9], sc_G=[0, G, 0, 0, G, 0],
                        sc_Gsolid=[-0.1, 0.1], flags=flags)
        e.set_geometry(dom)
        e.init_equilibrium(r0 * dom, (1.1 - r0) * dom)
        e.step(5); e.synchronize()
        e.step(steps); e.synchronize()
        t = e.timing()
        rho, u = e.download_macros()
        out[name] = (float(dom.sum()) * steps / (t["ms"] * 1e-3) / 1e6, np.stack(rho + u))
        e.close()
    print(json.dumps({"lattice": "D3Q19", "size": n, "model": "EFS-MRT" if model == _lib.MODEL_EFS else "SC", "steps": steps,
                      "mlups_two_pass": out["two_pass"][0], "mlups_reference_ordered": out["reference_ordered"][0],
                      "max_abs_diff": float(np.abs(out["two_pass"][1] - out["reference_ordered"][1]).max())}))
