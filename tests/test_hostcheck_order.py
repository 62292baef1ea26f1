"""The items of one GPU launch run in no particular order; the host test hook's forward loop would hide an operator that reads
what another item of the same launch writes.  LBM_HOST_ORDER (backend.h::launch) runs the items backwards / in a strided
permutation: the oracle and golden comparisons of the CPU tier must hold unchanged (the variable is read once per process, so
the suites run in a subprocess)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("order,files", [("scatter", ["test_hostcheck_cg.py", "test_hostcheck_sc.py"]),
                                         ("reverse", ["test_hostcheck_slabs.py", "test_hostcheck_properties.py"])])
def test_no_operator_depends_on_the_item_order_of_its_launch(order, files):
    env = dict(os.environ, LBM_HOST_ORDER=order)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider"] + [os.path.join(HERE, f) for f in files],
                       env=env, cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
