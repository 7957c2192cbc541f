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
