#!/usr/bin/env python
"""The UNMODIFIED reference kernels and drivers on the GPU this process sees (real numba.cuda; oracle/ref_numba.py target "cuda"):
BASELINE.md section 4.3's "reference Numba-CUDA" row for configurations 1-3.  Prints one JSON line per configuration.
    python scripts/ref_numba_cuda.py [cfg ...]          (GPU box; the reference files are vendored under baseline/_ref)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_numba
from openlbmpm_b200 import synthetic


def main():
    for cfg in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
        scale = 0.5 if cfg == 3 else 1.0        # the reference's Python set-up of 1024^2 takes minutes: same generator at 512^2
        dom, reg = synthetic.baseline_inputs_2d(cfg, scale)
        ny, nx = dom.shape
        steps = {1: 1000, 2: 400, 3: 400}[cfg]
        if cfg == 2:
            out = ref_numba.run_cg2d(nx, ny, steps, target="cuda", dom=dom, red=reg)
        elif cfg == 1:
            out = ref_numba.run_sc2d(nx, ny, steps, target="cuda", model="ShanChen", dom=dom, region0=reg)
        else:
            out = ref_numba.run_sc2d(nx, ny, steps, target="cuda", model="EFS", dom=dom, region0=reg,
                                     par=dict(relax="MRT", G=0.2, Gs0=-0.14, Gs1=0.14, inlet="Neumann", outlet="Dirichlet", vy1=-5.03e-4,
                                              bg0=0.02, bg1=0.02))
        dt = np.asarray(out["step_seconds"])[steps // 4:]
        print(json.dumps({"impl": "reference Numba-CUDA (unmodified kernels and driver loop, device synchronised once per iteration)",
                          "config": "BASELINE configuration %d, %d x %d" % (cfg, nx, ny), "void_nodes": out["n_fluid"],
                          "ms_per_step": float(dt.mean()) * 1e3, "value": out["n_fluid"] / float(dt.mean()) / 1e6, "unit": "MLUPS",
                          "timed_iterations": int(dt.size)}), flush=True)


if __name__ == "__main__":
    main()
