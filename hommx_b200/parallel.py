"""Sharding of macro cells over GPUs and the one exchange step of the hot path.

The reference partitions macro cells with the DOLFINx mesh partition, every MPI rank loops over
its owned cells (/root/reference/src/hommx/hmm.py:307) and the only communication on the hot
path is PETSc's assembly of rows shared between ranks (``self._A.assemble()``, hmm.py:442).
Here: one process per GPU (``torch.distributed``, NCCL over NVLink), contiguous blocks of macro
cells per rank (``assembly.shard_range``), every rank scatters into its own CSR value array, and
the value slots touched by more than one rank are summed with a single all-reduce.
"""
from __future__ import annotations


class HaloExchange:
    """Sum of the shared CSR value slots across ranks.

    ``pack(vals, slots, n, buf)`` and ``unpack(vals, slots, n, buf)`` are the device kernels of
    libhmx (``CellSolver.halo_pack_dev`` / ``halo_unpack_dev``); they are parameters so that the
    host logic can be exercised with the gloo backend on CPU tensors in the tests.
    """

    def __init__(self, shared_slots, buf, pack, unpack, group=None):
        self.slots, self.buf, self.pack, self.unpack, self.group = shared_slots, buf, pack, unpack, group
        self.n = int(shared_slots.numel())

    def sum(self, vals):
        import torch.distributed as dist

        if self.n == 0:
            return
        self.pack(vals, self.slots, self.n, self.buf)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        self.unpack(vals, self.slots, self.n, self.buf)

    @property
    def bytes_per_exchange(self):
        return self.n * 8


def sample_cells(lo, hi, max_samples=256):
    """Indices of at most ``max_samples`` macro cells spread evenly over the rank's block ``[lo, hi)`` (a contiguous
    first block would not be representative of an x-dependent coefficient)."""
    import numpy as np

    n = hi - lo
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    m = min(n, max_samples)
    return lo + (np.arange(m, dtype=np.int64) * n) // m


def agree_on_mean(local_sum, local_count, device=None, group=None):
    """Mean of a per-rank statistic over ALL ranks (one all-reduce of two numbers), so that a decision derived from
    it -- which cell kernel to run -- is the same on every rank: the sharded assembly then equals the single-GPU one
    bit for bit wherever the summation order allows.  Without an initialised process group: the local mean."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sum / max(local_count, 1)
    t = torch.tensor([float(local_sum), float(local_count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    s, c = t.tolist()
    return s / max(c, 1.0)
