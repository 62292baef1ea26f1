"""Builds tests/hostcheck/libhostcheck.so: the library's operator code (openlbmpm_b200/csrc/*.cuh, *.cu)
compiled for the HOST with g++ -DLBM_HOSTCHECK.  TEST HOOK ONLY: it lets the CPU-only test tier check
the node arithmetic and the step orchestration against the oracle without a GPU.  The package never
loads it (openlbmpm_b200/_lib.py only knows liblbmpm.so) and the tiled CUDA kernels are not in it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "openlbmpm_b200", "csrc")
OUT = os.path.join(HERE, "libhostcheck.so")
SOURCES = ["lbm_api.cu", "sc_api.cu", "tr_api.cu", "cg_fast.cu", "host_stubs.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(d) for d in deps):
        return OUT
    # one compiler process per translation unit, side by side (the five units take ~3 min one after the other)
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-DLBM_HOSTCHECK", "-ffp-contract=off", "-fPIC", "-pthread"]
    objs = [os.path.join(objdir, os.path.basename(s_) + ".o") for s_ in srcs]
    procs = [subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", s_, "-o", o]) for s_, o in zip(srcs, objs)]
    if any(p.wait() != 0 for p in procs):
        raise subprocess.CalledProcessError(1, "g++ (host test hook)")
    subprocess.check_call(["g++", "-shared", "-pthread"] + objs + ["-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
