"""Host side of the slab decomposition (one process per GPU, `torch.distributed` for the plumbing only).

The lattice is cut along the flow axis (z in 3-D, y in 2-D) into `world` slabs of equal thickness; rank r owns
planes [r n/P, (r+1) n/P).  Ghost planes travel inside liblbmpm.so (NCCL send/recv, comm.cu); this module only
partitions host arrays, hands the NCCL unique id around and gathers results.  `ring_exchange_reference` restates
the send/recv pairing of comm.cu::ring_exchange with torch.distributed point-to-point calls so that the CPU test
tier (gloo, world_size 2) can check the protocol without a GPU."""
import numpy as np


def slab_bounds(n, rank, world):
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    if n % world:
        raise ValueError("the flow axis (%d planes) must be divisible by the number of slabs (%d)" % (n, world))
    t = n // world
    if t < 3:
        raise ValueError("a slab needs at least 3 planes")
    return rank * t, (rank + 1) * t


def take_slab(a, rank, world, axis=0):
    lo, hi = slab_bounds(a.shape[axis], rank, world)
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(lo, hi)
    return np.ascontiguousarray(a[tuple(idx)])


def share_unique_id(dist, engine, rank, device=None):
    """rank 0 creates the NCCL unique id (lbm_nccl_unique_id), everybody receives it"""
    import torch
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(engine.nccl_unique_id().copy())
    if device is not None:
        uid = uid.to(device)
    dist.broadcast(uid, 0)
    return uid.cpu().numpy()


def gather_slabs(dist, local, world, device=None):
    """all ranks receive the slabs of every rank concatenated along axis 0"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.concatenate([p.cpu().numpy() for p in parts], axis=0)


def ring_exchange_reference(dist, padded, ng, gp, rank, world):
    """Fill the ghost planes of `padded` ([n_local + 2 ng, ...] torch tensor) exactly as comm.cu does:
    my top gp planes -> low ghost of rank+1, my bottom gp planes -> high ghost of rank-1 (periodic ring)."""
    up, down = (rank + 1) % world, (rank - 1 + world) % world
    n = padded.shape[0] - 2 * ng
    top = padded[ng + n - gp:ng + n].contiguous()
    bottom = padded[ng:ng + gp].contiguous()
    lo = padded[ng - gp:ng].clone()
    hi = padded[ng + n:ng + n + gp].clone()
    ops = [dist.P2POp(dist.isend, top, up), dist.P2POp(dist.irecv, lo, down),
           dist.P2POp(dist.isend, bottom, down), dist.P2POp(dist.irecv, hi, up)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    padded[ng - gp:ng] = lo
    padded[ng + n:ng + n + gp] = hi
    return padded


# ---------------------------------------------------------------------------------------------------
# what a host class needs from "the other ranks": the NCCL id, gathers of result slabs, small sums
# ---------------------------------------------------------------------------------------------------
class TorchSlabs:
    """one process per GPU, launched by `torchrun` (RANK / LOCAL_RANK / WORLD_SIZE in the environment)"""

    def __init__(self, dist, rank, world, local_rank):
        self.dist, self.rank, self.world, self.local_rank = dist, rank, world, local_rank
        self.device_index = local_rank

    def unique_id(self, engine):
        return share_unique_id(self.dist, engine, self.rank, device="cuda")

    def gather(self, local):
        return gather_slabs(self.dist, local, self.world, device="cuda")

    def sum(self, values):
        import torch
        t = torch.as_tensor(np.asarray(values, dtype=np.float64), device="cuda")
        self.dist.all_reduce(t)
        return t.cpu().numpy()


def from_environment():
    """-> TorchSlabs when the process was launched as one rank of several (`torchrun`), else None"""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return TorchSlabs(dist, int(os.environ.get("RANK", "0")), world, local)


class ThreadSlabs:
    """the ranks are threads of one process (CPU test tier, with the host test hook's in-process ring)"""

    class Shared:
        def __init__(self, world):
            import threading
            self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world, self.device_index = shared, rank, shared.world, 0

    def _exchange(self, value):
        s = self.shared
        s.slots[self.rank] = value
        s.barrier.wait()
        got = list(s.slots)
        s.barrier.wait()
        return got

    def unique_id(self, engine):
        return self._exchange(engine.nccl_unique_id() if self.rank == 0 else None)[0]

    def gather(self, local):
        return np.concatenate(self._exchange(np.ascontiguousarray(local)), axis=0)

    def sum(self, values):
        return np.sum(self._exchange(np.asarray(values, dtype=np.float64)), axis=0)
