"""Pins oracle/ref_numba.py's cpu target -- the UNMODIFIED reference drivers and Numba kernels with the kernels re-targeted
from cuda.jit to numba.njit(parallel) -- against the committed golden vectors (which the same reference code produced under
Numba's CUDA simulator, tests/golden/gen_goldens_*.py).  It is bench.py's reference arm for BASELINE configurations 1-3, so the
thing timed there is shown here to compute what the reference computes.  Skipped where the reference sources are neither
mounted (/root/reference) nor vendored (baseline/_ref, done by __graft_entry__.build())."""
import os

import numpy as np
import pytest

from oracle import ref_numba

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.timeout(600)


def _available():
    try:
        ref_numba.reference_root()
        import numba  # noqa: F401
        return True
    except Exception:
        return False


needs_reference = pytest.mark.skipif(not _available(), reason="reference sources / numba not available")


def _params(g):
    return dict(zip([str(k) for k in g["params_keys"]], [str(v) for v in g["params_vals"]]))


@needs_reference
def test_colour_gradient_cpu_target_reproduces_the_reference_vectors():
    for case in ("channel_neumann_dirichlet_mrt", "block_w2_mrt"):
        g = np.load(os.path.join(GOLD, "cg2d_%s.npz" % case), allow_pickle=True)
        p = _params(g)
        par = {k: float(p[k]) for k in ("sigma", "theta", "beta", "delta", "tauR", "tauB", "vyr")}
        par.update(wetting=int(p["wetting"]), tautype=int(p["tautype"]), inlet=p["inlet"], outlet=p["outlet"], relax=p["relax"])
        out = ref_numba.run_cg2d(int(g["nx"]), int(g["ny"]), int(p["steps"]), target="cpu", dom=g["is_domain"], red=g["red_mask"],
                                 minor=float(g["minor"]), interval=1, par=par)
        assert len(out["snapshots"]) == g["rhoR"].shape[0]
        for k in ("rhoR", "rhoB", "ux", "uy"):
            got = np.stack([s[k] for s in out["snapshots"]])
            np.testing.assert_allclose(got, g[k], rtol=0, atol=1e-13, err_msg="%s %s" % (case, k))
        assert np.array_equal(out["sim"].neighboringNodes, g["neighboringNodes"])


@needs_reference
def test_shan_chen_cpu_target_reproduces_the_reference_vectors():
    for case in ("sc_droplet", "efs_channel_neumann_dirichlet_mrt"):
        g = np.load(os.path.join(GOLD, "sc2d_%s.npz" % case), allow_pickle=True)
        p = _params(g)
        par = {k: float(p[k]) for k in ("rho0", "rho1", "bg0", "bg1", "tau0", "tau1", "G", "Gs0", "Gs1", "vy0", "vy1")}
        par.update(relax=p["relax"], inlet=p["inlet"], outlet=p["outlet"], scheme=int(p.get("scheme", 4)))
        nx, ny = int(g["nx"]), int(g["ny"])
        out = ref_numba.run_sc2d(nx, ny, int(p["steps"]), target="cpu", model=str(g["model"]), dom=g["is_domain"], region0=g["region0"],
                                 par=par, keep_states=True)
        fn = out["sim"].fluidNodes
        rho = np.zeros((len(out["states"]), 2, ny * nx))
        for i, s in enumerate(out["states"]):
            rho[i][:, fn] = s["rho"]
        np.testing.assert_allclose(rho.reshape(-1, 2, ny, nx), g["rho"], rtol=0, atol=1e-13, err_msg=case)
