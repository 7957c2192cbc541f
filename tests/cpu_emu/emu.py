"""Build and run the CPU emulation of a cell kernel (TEST INFRASTRUCTURE ONLY).

Compiles hommx_b200/csrc/hmx_cell_entry.cu -- the very source nvcc compiles for the GPU --
with g++ and -DHMX_EMULATE (tests/cpu_emu/cuda_shim.h), so kernel logic can be checked against
the oracle without a GPU.  The product never imports this module.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from hommx_b200 import native

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


class CellParams(C.Structure):
    _fields_ = [
        ("n_pts", C.c_int64), ("x_pts", C.c_void_p), ("cell_nodes", C.c_void_p), ("node_xyz", C.c_void_p),
        ("A_hom", C.c_void_p), ("S_loc", C.c_void_p), ("iters", C.c_void_p), ("resid", C.c_void_p),
        ("qp", C.c_void_p), ("qw", C.c_void_p), ("scratch", C.c_void_p), ("chi", C.c_void_p), ("work", C.c_void_p),
        ("nq", C.c_int32), ("max_it", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double), ("mesh", C.c_void_p),
    ]  # fmt: skip


class MicroMesh(C.Structure):
    """struct MicroMesh of csrc/hmx_cell_common.cuh (here with host pointers)."""

    _fields_ = [
        ("n_elem", C.c_int32), ("n_nodes", C.c_int32), ("nnzb", C.c_int32), ("nq", C.c_int32),
        ("elem_nodes", C.c_void_p), ("elem_grad", C.c_void_p), ("elem_vol", C.c_void_p), ("elem_yq", C.c_void_p),
        ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("blk_ptr", C.c_void_p), ("blk_src", C.c_void_p),
        ("node_ptr", C.c_void_p), ("node_src", C.c_void_p), ("diag", C.c_void_p),
    ]  # fmt: skip


def build(prog, n, threads=None, variant=None, collapse=False):
    threads, min_blocks, variant, coll = native.resolve(prog, n, threads, None, variant, collapse)
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(HERE, "cuda_shim.h"), "rb") as f:
        shim = f.read()
    with open(os.path.join(HERE, "emu_runtime.h"), "rb") as f:
        shim += f.read()
    key = native.kernel_key(prog, n, threads, min_blocks, variant, coll) + "_" + hashlib.sha1(shim).hexdigest()[:8]
    so = os.path.join(BUILD, key + ".so")
    if not os.path.exists(so):
        coeff = os.path.join(BUILD, key + ".coeff.h")
        with open(coeff, "w") as f:
            f.write(prog.source)
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DHMX_EMULATE", "-I", HERE, "-I", native.CSRC,
               *native.kernel_defines(prog, n, threads, coeff, min_blocks, variant, coll), *native._extra_flags(), "-x", "c++", os.path.join(native.CSRC, "hmx_cell_entry.cu"),
               "-o", so + ".tmp", "-lpthread"]  # fmt: skip
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stderr[-6000:])
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    lib.hmx_emu_launch.argtypes = [C.POINTER(CellParams), C.c_int, C.c_int]
    lib.hmx_emu_info.argtypes = [C.POINTER(C.c_int)]
    lib.hmx_emu_cluster.restype = C.c_int
    return lib


class EmuSolver:
    """Same call surface as hommx_b200.native.CellSolver's host entry points."""

    def __init__(self, prog, n, qp, qw, rtol=1e-8, atol=1e-10, max_it=10000, threads=None, grid=4, variant=None, collapse=False,
                 micro_tables=None):
        self.tables = t = micro_tables
        self.mesh = None
        if t is not None:  # general periodic micro mesh: the element-list kernel
            n, variant, collapse, qw = 0, native.ELEMENT_LIST, False, t.qw
            qp = np.zeros((1, len(qw), prog.dim))
            p = lambda a: a.ctypes.data  # noqa: E731
            self.mesh = MicroMesh(t.n_elem, t.n_nodes, t.nnzb, t.nq, p(t.elem_nodes), p(t.elem_grad), p(t.elem_vol), p(t.elem_yq),
                                  p(t.row_ptr), p(t.col), p(t.blk_ptr), p(t.blk_src), p(t.node_ptr), p(t.node_src), p(t.diag))  # fmt: skip
        self.prog, self.n = prog, n
        self.lib = build(prog, n, threads, variant, collapse)
        self.collapse = collapse
        info = (C.c_int * 8)()
        self.lib.hmx_emu_info(info)
        self.info = list(info)
        self.qp = np.ascontiguousarray(qp, dtype=np.float64)
        self.qw = np.ascontiguousarray(qw, dtype=np.float64)
        self.rtol, self.atol, self.max_it, self.grid = rtol, atol, max_it, grid
        self.m = prog.n_rhs
        self.nb = (prog.dim + 1) * (1 if prog.kind == 0 else prog.dim)

    def _launch(self, P, n):
        cl = self.lib.hmx_emu_cluster()  # CTAs per thread-block cluster (1: ordinary launch)
        grid = cl * max(1, min(self.grid // cl if cl > 1 else self.grid, n))
        per_cta = self.info[6]
        if self.tables is not None:
            bs = 1 if self.prog.kind == 0 else self.prog.dim
            per_cta = self.tables.scratch_doubles(self.prog.natoms, bs, self.m)
            P.mesh = C.addressof(self.mesh)
        scratch = np.zeros(max(1, per_cta * grid))
        P.scratch = scratch.ctypes.data
        P.qp, P.qw, P.nq = self.qp.ctypes.data, self.qw.ctypes.data, len(self.qw)
        P.max_it, P.rtol, P.atol = self.max_it, self.rtol, self.atol
        self.lib.hmx_emu_launch(C.byref(P), grid, min(grid, os.cpu_count() or 1))

    def cell_tensors(self, x_pts, return_stats=False):
        x = np.ascontiguousarray(np.asarray(x_pts, dtype=np.float64).reshape(-1, 3))
        n = len(x)
        A = np.zeros((n, self.m, self.m))
        it = np.zeros(n, dtype=np.int32)
        res = np.zeros(n)
        P = CellParams()
        P.n_pts, P.x_pts, P.A_hom, P.iters, P.resid = n, x.ctypes.data, A.ctypes.data, it.ctypes.data, res.ctypes.data
        self._launch(P, n)
        return (A, it, res) if return_stats else A

    def correctors(self, x_pts):
        x = np.ascontiguousarray(np.asarray(x_pts, dtype=np.float64).reshape(-1, 3))
        n = len(x)
        d = self.prog.dim
        coll = native.collapse_mask(self.prog, self.collapse)
        shape = [1 if (coll >> a) & 1 else self.n for a in range(d)]
        if self.tables is not None:
            shape = [self.tables.n_nodes]
        bs = 1 if self.prog.kind == 0 else d
        chi = np.zeros((n, self.m, bs) + tuple(reversed(shape)))
        A = np.zeros((n, self.m, self.m))
        P = CellParams()
        P.n_pts, P.x_pts, P.A_hom, P.chi = n, x.ctypes.data, A.ctypes.data, chi.ctypes.data
        self._launch(P, n)
        return chi

    def local_matrices(self, cell_nodes, node_xyz):
        cells = np.ascontiguousarray(cell_nodes, dtype=np.int32)
        xyz = np.ascontiguousarray(node_xyz, dtype=np.float64)
        n = len(cells)
        S = np.zeros((n, self.nb, self.nb))
        A = np.zeros((n, self.m, self.m))
        P = CellParams()
        P.n_pts, P.cell_nodes, P.node_xyz, P.S_loc, P.A_hom = n, cells.ctypes.data, xyz.ctypes.data, S.ctypes.data, A.ctypes.data
        self._launch(P, n)
        return S, A
