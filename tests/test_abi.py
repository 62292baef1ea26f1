"""The C-ABI library loads and exports every symbol include/lbmpm.h declares (no compute calls: there is
no GPU in the CPU tier), and the ctypes mirror of lbm_config has the C layout."""
import ctypes
import os
import re
import subprocess

import pytest

from openlbmpm_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lbmpm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.PROTOTYPES, "no ctypes prototype for " + s
    assert lib.lbm_abi_version() == _lib.ABI_VERSION


def test_config_layout_matches_c(tmp_path):
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lbmpm.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                 'sizeof(lbm_config), offsetof(lbm_config, sigma), offsetof(lbm_config, sc_tau), offsetof(lbm_config, reserved_d));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out == [ctypes.sizeof(_lib.LbmConfig), _lib.LbmConfig.sigma.offset, _lib.LbmConfig.sc_tau.offset,
                   _lib.LbmConfig.reserved_d.offset]


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.LbmError, match="no CUDA device|CUDA"):
        _lib.Engine(9, (8, 8))


def test_bad_config_rejected(lib):
    cfg = _lib.LbmConfig()
    cfg.abi_version = 99
    h = ctypes.c_void_p()
    assert lib.lbm_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"abi_version" in lib.lbm_last_error(None)


def test_c_program_against_the_abi(tmp_path):
    """examples/abi_minimal.c: the header is plain C (gcc -std=c99) and a C caller can drive the whole path; linked here
    against the host test hook (same entry points), on a GPU box against liblbmpm.so"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostcheck"))
    import build as hostcheck_build
    so = hostcheck_build.build()
    exe = tmp_path / "abi_minimal"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_minimal.c"), so, "-Wl,-rpath," + os.path.dirname(so), "-lm", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "abi_minimal ok" in out.stdout, out.stdout + out.stderr
