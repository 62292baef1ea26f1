#!/usr/bin/env python
"""Runnable equivalent of the reference's main.py (main.py:1-81): same prompts, same class names and run
methods, but the three `input()` answers can also come from argv or the environment, e.g.
    python main.py 2D flow CG            (or LBM_DIMENSION=2D LBM_MODEL=flow LBM_METHOD=CG python main.py)
    python main.py 3D flow CG --ini IniFiles
    torchrun --nproc-per-node 8 main.py 3D flow CG --ini IniFiles      (slab decomposition along the flow axis, one GPU per rank)
The .ini directory defaults to ./IniFiles (the reference hard-codes the same relative path)."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the reference imports its classes as top-level modules; make the same names importable
sys.path.insert(0, os.path.join(ROOT, "openlbmpm_b200"))

from openlbmpm_b200.RKD2Q9 import RKColorGradientLBM
from openlbmpm_b200.RKColorGradientD3Q19 import RKColorGradient3D
from openlbmpm_b200.ShanChenD2Q9 import ShanChenD2Q9
from openlbmpm_b200.ShanChenD3Q19 import ShanChenD3Q19
from openlbmpm_b200.Transport2D import Transport2D, Transport2DRK  # noqa: F401


def ask(pos, env, prompt, args):
    if len(args) > pos:
        return args[pos]
    if os.environ.get(env):
        return os.environ[env]
    return input(prompt)


def main(argv):
    ini = os.path.join(os.getcwd(), "IniFiles")
    if "--ini" in argv:
        i = argv.index("--ini")
        ini = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    os.makedirs(os.path.expanduser("~/LBMResults"), exist_ok=True)
    dim = ask(0, "LBM_DIMENSION", "Please choose 2D/3D model (enter 2D or 3D):", argv)
    kind = ask(1, "LBM_MODEL", "Please choose the type of the simulation (enter flow or transport): ", argv)
    method = ask(2, "LBM_METHOD", "Please choose ShanChen(SC) or Color Gradient(CG) methods for flow: ", argv)
    if kind == "transport" and dim == "2D" and method == "CG":
        Transport2DRK(ini).runTransport2DMPMCRK()               # main.py:66-68
        return 0
    if kind != "flow":
        print("Solute transport on the Shan-Chen flow is outside this build's scope. Stop here.")
        return 2
    if dim == "2D" and method == "SC":
        ShanChenD2Q9(ini).runTypeSCmodel()
    elif dim == "2D" and method == "CG":
        RKColorGradientLBM(ini).runModifiedRKColorGradient2D()
    elif dim == "3D" and method == "CG":
        RKColorGradient3D(ini).runRKColorGradient3D()
    elif dim == "3D" and method == "SC":
        ShanChenD3Q19(ini).runEFS4LBM3DGPU()
    else:
        print("The chosen type does not exist in current version. Stop here.")
        return 2
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # launched as `torchrun --nproc-per-node P main.py ...`: one slab per GPU
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
