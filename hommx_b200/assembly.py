"""Macro CSR pattern and the deterministic slot map (host side, setup only).

The reference inserts every local matrix with ``MatSetValues(..., ADD_VALUES)`` into an
un-preallocated AIJ matrix (/root/reference/src/hommx/hmm.py:144-149, 325-330).  Here the
sparsity is computed once from the macro dofmap, every local entry ``(cell, i, j)`` gets its CSR
value slot, and the slot map is inverted into gather lists so that the GPU sums the sources of
a slot in a fixed order (no atomics -> bitwise reproducible).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def unroll_dofs(cells, bs):
    """Blocked -> unrolled dof indices, node-major / component-minor (hmm.py:31-40)."""
    cells = np.asarray(cells, dtype=np.int64)
    if bs == 1:
        return cells
    return (cells[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(len(cells), -1)


@dataclass
class MacroPattern:
    """CSR pattern of the macro stiffness matrix and the slot of every local entry."""

    n_dofs: int
    indptr: np.ndarray  # (n_dofs+1,) int64
    indices: np.ndarray  # (nnz,) int32 column indices, sorted within each row
    slot_map: np.ndarray  # (n_cells, nb*nb) int64: CSR slot of S_loc[c].ravel()[k]

    @property
    def nnz(self):
        return len(self.indices)


def build_pattern(cells, n_nodes, bs):
    dofs = unroll_dofs(cells, bs)  # (nc, nb)
    nc, nb = dofs.shape
    n_dofs = int(n_nodes) * bs
    rows = np.repeat(dofs, nb, axis=1).ravel()  # entry k = i*nb + j -> row dof_i
    cols = np.tile(dofs, (1, nb)).ravel()
    key = rows * n_dofs + cols
    uniq, inv = np.unique(key, return_inverse=True)
    urows = uniq // n_dofs
    indptr = np.zeros(n_dofs + 1, dtype=np.int64)
    np.add.at(indptr, urows + 1, 1)
    np.cumsum(indptr, out=indptr)
    return MacroPattern(n_dofs, indptr, (uniq % n_dofs).astype(np.int32), inv.reshape(nc, nb * nb).astype(np.int64))


@dataclass
class GatherMap:
    """Inverse of a slot map restricted to a set of cells: sources of every CSR slot."""

    ptr: np.ndarray  # (nnz+1,) int64
    src: np.ndarray  # (n_local_cells*nb*nb,) int32 indices into the local S_loc array


def build_gather(slot_map_local, nnz):
    flat = np.asarray(slot_map_local, dtype=np.int64).ravel()
    if len(flat) > np.iinfo(np.int32).max:
        raise ValueError("too many local matrix entries for int32 gather indices; shard the macro cells")
    order = np.argsort(flat, kind="stable")  # fixed order: by slot, then by (cell, i, j)
    ptr = np.zeros(nnz + 1, dtype=np.int64)
    np.add.at(ptr, flat + 1, 1)
    np.cumsum(ptr, out=ptr)
    return GatherMap(ptr, order.astype(np.int32))


def shard_range(n_cells, rank, world):
    """Contiguous block of macro cells owned by ``rank`` (the reference loops over the cells its
    MPI rank owns, hmm.py:307)."""
    base, rem = divmod(int(n_cells), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shared_slots(slot_map, n_cells, world, nnz):
    """CSR slots that receive contributions from more than one rank: the payload of the one
    exchange step (MatAssembly of shared rows, hmm.py:442)."""
    touched = np.zeros(nnz, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_range(n_cells, r, world)
        mark = np.zeros(nnz, dtype=bool)
        mark[slot_map[lo:hi].ravel()] = True
        touched += mark
    return np.nonzero(touched > 1)[0].astype(np.int64)


@dataclass
class LocalShard:
    """Everything one rank keeps on its GPU, in LOCAL numbering: only the macro nodes its cells reference and only
    the CSR value slots its cells touch (+ one dummy slot), so that the device state and the host<->device traffic
    of a rank do not grow with the number of ranks (the reference keeps owned rows + ghosts per MPI rank the same
    way).  ``slots`` maps local slot -> global CSR slot (sorted)."""

    lo: int
    hi: int
    cells: np.ndarray  # (n_local, d+1) int32, local node ids
    nodes: np.ndarray  # (n_local_nodes,) int64 global node id of every local node
    slots: np.ndarray  # (nnz_local,) int64 global CSR slot of every local slot
    gather: GatherMap  # over local slots
    shared: np.ndarray  # (n_shared_global,) int64: local slot of every globally shared slot, nnz_local (dummy) if untouched
    owned: np.ndarray  # (nnz_local,) bool: this rank contributes the slot to a globally complete value array

    @property
    def nnz(self):
        return len(self.slots)


def build_local_shard(cells, slot_map, nnz, rank, world):
    """Local view of rank ``rank`` of ``world`` (contiguous cell blocks, ``shard_range``)."""
    cells = np.asarray(cells)
    n_cells = len(cells)
    lo, hi = shard_range(n_cells, rank, world)
    nodes, inv = np.unique(cells[lo:hi], return_inverse=True)
    cells_l = inv.reshape(hi - lo, cells.shape[1]).astype(np.int32)
    slots, sinv = np.unique(slot_map[lo:hi], return_inverse=True)
    gm = build_gather(sinv.reshape(hi - lo, -1), len(slots))
    owned = np.ones(len(slots), dtype=bool)
    shared_local = np.zeros(0, dtype=np.int64)
    if world > 1:
        # lowest rank touching every slot (host, set-up only)
        first = np.full(nnz, world, dtype=np.int32)
        count = np.zeros(nnz, dtype=np.int32)
        for r in range(world - 1, -1, -1):
            rlo, rhi = shard_range(n_cells, r, world)
            t = np.unique(slot_map[rlo:rhi])
            first[t] = r
            count[t] += 1
        sh = np.nonzero(count > 1)[0].astype(np.int64)
        pos = np.searchsorted(slots, sh)
        hit = (pos < len(slots)) & (slots[np.minimum(pos, max(len(slots) - 1, 0))] == sh) if len(slots) else np.zeros(len(sh), bool)
        shared_local = np.where(hit, pos, len(slots)).astype(np.int64)
        owned = first[slots] == rank
    return LocalShard(lo, hi, cells_l, nodes.astype(np.int64), slots.astype(np.int64), gm, shared_local, owned)
