"""N > 1 host logic on the CPU: two gloo processes partition a lattice, exchange ghost planes with the same
pairing as comm.cu (ring_exchange_reference) and must reproduce the periodic wrap of the global array; the
unique-id broadcast and the slab gather are exercised too."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from openlbmpm_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from openlbmpm_b200 import slab
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    rng = np.random.default_rng(7)
    glob = rng.random((12, 5, 6))
    mine = slab.take_slab(glob, rank, world)
    NG = 3
    padded = torch.zeros((mine.shape[0] + 2 * NG,) + mine.shape[1:], dtype=torch.float64)
    padded[NG:-NG] = torch.from_numpy(mine)
    for gp in (1, 3):
        slab.ring_exchange_reference(dist, padded, NG, gp, rank, world)
        lo, hi = slab.slab_bounds(glob.shape[0], rank, world)
        want_lo = np.take(glob, range(lo - gp, lo), axis=0, mode="wrap")
        want_hi = np.take(glob, range(hi, hi + gp), axis=0, mode="wrap")
        assert np.array_equal(padded[NG - gp:NG].numpy(), want_lo), "low ghost"
        assert np.array_equal(padded[-NG:padded.shape[0] - NG + gp].numpy(), want_hi), "high ghost"
    class FakeEngine:
        def nccl_unique_id(self):
            return np.arange(128, dtype=np.uint8)
    uid = slab.share_unique_id(dist, FakeEngine(), rank)
    assert np.array_equal(uid, np.arange(128, dtype=np.uint8))
    full = slab.gather_slabs(dist, mine, world)
    assert np.array_equal(full, glob)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_slab_bounds():
    assert slab.slab_bounds(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        slab.slab_bounds(10, 0, 4)
    with pytest.raises(ValueError):
        slab.slab_bounds(8, 0, 4)          # 2 planes per slab


def test_two_rank_ring_exchange_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
