#!/usr/bin/env python
"""Wall-clock cost of writing records every `interval` steps: blocking download (the reference's pattern,
RKD2Q9.py:1382-1393) vs lbm_download_macros_async + a writer thread.
    python scripts/bench_output.py [--size 256] [--steps 120] [--interval 10]
Prints one JSON line.  D3Q19 colour-gradient MRT periodic box; a "record" = rhoR, rhoB, ux, uy, uz copied to the host
(and, with --write, compressed into an .npz by the writer)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openlbmpm_b200 import _lib
from openlbmpm_b200.results import AsyncMacroOutput


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--interval", type=int, default=10)
    ap.add_argument("--write", action="store_true")
    a = ap.parse_args()
    n = (a.size,) * 3
    eng = _lib.Engine(19, n, sigma=0.1, beta=0.7)
    eng.set_geometry(np.ones(n, np.uint8))
    eng.init_spinodal_device()
    eng.step(10); eng.synchronize()
    tmp = tempfile.mkdtemp()

    def sink(rec, rho, u):
        if a.write:
            np.savez(os.path.join(tmp, "rec%04d.npz" % rec), rhoR=rho[0], rhoB=rho[1], ux=u[0], uy=u[1], uz=u[2])

    def run(mode):
        out = AsyncMacroOutput(eng, sink) if mode == "async" else None
        bufs = ([np.empty(n), np.empty(n)], [np.empty(n) for _ in range(3)])
        t0 = time.perf_counter()
        for s in range(0, a.steps, a.interval):
            if mode == "async":
                out.snapshot()
            elif mode == "blocking":
                eng.download_macros(*bufs)
                sink(s // a.interval, *bufs)
            eng.step(a.interval)
        eng.synchronize()
        if out is not None:
            out.close()
        return time.perf_counter() - t0

    res = {}
    for mode in ("none", "blocking", "async", "none", "blocking", "async"):
        res.setdefault(mode, []).append(run(mode))
    best = {k: min(v) for k, v in res.items()}
    nodes = float(a.size) ** 3
    print(json.dumps({"workload": "D3Q19 CG-MRT %d^3 periodic box, %d steps, one record every %d steps (%.0f MB each)%s" % (
                          a.size, a.steps, a.interval, 5 * nodes * 8 / 1e6, ", written as .npz" if a.write else ""),
                      "seconds": best, "mlups": {k: nodes * a.steps / v / 1e6 for k, v in best.items()},
                      "output_overhead_pct": {k: 100.0 * (best[k] / best["none"] - 1.0) for k in ("blocking", "async")}}))
    eng.close()


if __name__ == "__main__":
    main()
