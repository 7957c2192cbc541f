"""world_size-2 (and 3) test of the sharded assembly logic on CPU with the gloo backend: cell
ranges, per-rank gather maps, shared slots and the halo exchange reproduce the single-rank CSR
values.  (The per-cell matrices are synthetic here; the kernels are covered by the gpu tests.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hommx_b200 import assembly, mesh, parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, bs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = mesh.create_box((0, 0, 0), (1, 0.4, 0.1), (4, 2, 3))
    nb = 4 * bs
    pat = assembly.build_pattern(m.cells, m.num_nodes, bs)
    S = np.random.default_rng(7).normal(size=(m.num_cells, nb * nb))  # same on every rank
    lo, hi = assembly.shard_range(m.num_cells, rank, world)
    gm = assembly.build_gather(pat.slot_map[lo:hi], pat.nnz)
    Sl = S[lo:hi].ravel()
    vals = torch.tensor([Sl[gm.src[gm.ptr[s] : gm.ptr[s + 1]]].sum() for s in range(pat.nnz)])
    sh = torch.as_tensor(assembly.shared_slots(pat.slot_map, m.num_cells, world, pat.nnz))
    halo = parallel.HaloExchange(
        sh, torch.zeros(len(sh), dtype=torch.float64),
        pack=lambda v, s, n, b: b.copy_(v[s]), unpack=lambda v, s, n, b: v.index_copy_(0, s, b),
    )  # fmt: skip
    halo.sum(vals)
    full = np.zeros(pat.nnz)
    np.add.at(full, pat.slot_map.ravel(), S.ravel())
    touched = np.zeros(pat.nnz, dtype=bool)
    touched[pat.slot_map[lo:hi].ravel()] = True
    is_shared = np.zeros(pat.nnz, dtype=bool)
    is_shared[sh.numpy()] = True
    got = vals.numpy()
    # slots this rank touches, and every shared slot, hold the global sum; the rest stay zero
    ok = np.allclose(got[touched | is_shared], full[touched | is_shared], rtol=1e-13, atol=1e-13)
    ok = ok and np.all(got[~touched & ~is_shared] == 0.0)
    # complete values everywhere (what the scipy macro solve of the stand-in host uses)
    v = vals.clone()
    if rank != 0:
        v[sh] = 0.0
    dist.all_reduce(v)
    ok = ok and np.allclose(v.numpy(), full, rtol=1e-13, atol=1e-13)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bs", [(2, 1), (2, 3), (3, 1)])
def test_sharded_assembly_with_halo_exchange(world, bs):
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), bs, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def _decision_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 sees an easy coefficient region (10 iterations per solve), rank 1 a hard one (200), rank 2 has no cells:
    # decided alone they would pick different cell kernels; the collective mean is the same number everywhere
    n_solves = [256 * 6, 128 * 6, 0][rank]
    its = [10, 200, 0][rank] * n_solves
    out[rank] = parallel.agree_on_mean(its, n_solves)
    dist.destroy_process_group()


def test_cell_solver_decision_is_collective():
    """cell_solver="auto" (hommx_b200/hmm.py): the PCG-vs-direct decision is taken from the iteration mean over the
    samples of ALL ranks, so that every rank runs the same kernel."""
    world = 3
    out = mp.Manager().dict()
    mp.spawn(_decision_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    want = (10 * 256 * 6 + 200 * 128 * 6) / (256 * 6 + 128 * 6)
    assert all(abs(out[r] - want) < 1e-12 for r in range(world)), dict(out)
    assert parallel.agree_on_mean(30.0, 3) == 10.0  # no process group: the local mean


def test_sample_cells_spreads_over_the_block():
    idx = parallel.sample_cells(1000, 1000 + 5000, 256)
    assert len(idx) == 256 and idx[0] == 1000 and idx[-1] >= 1000 + 5000 - 5000 // 256 - 1 and np.all(np.diff(idx) > 0)
    assert list(parallel.sample_cells(7, 10)) == [7, 8, 9] and len(parallel.sample_cells(5, 5)) == 0
