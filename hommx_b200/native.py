"""ctypes binding of libhmx.so (include/hmx.h) and the build of the coefficient-specialised
cell kernels.

The reference JIT-compiles ``1 + n_b + n_b^2`` FFCx forms per assembly
(/root/reference/src/hommx/hmm.py:259-274, 306).  Here one CUDA kernel is compiled per
(coefficient program, micro mesh size) with ``nvcc`` for sm_100a into a cubin that is cached
in-tree (``hommx_b200/_kcache``) and handed to ``hmx_create`` as an image.  There is no CPU
fallback: if the library or a CUDA device is missing every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import threading

import numpy as np

from .codegen import ELASTICITY, POISSON, CoefficientProgram

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
KCACHE = os.path.join(_HERE, "_kcache")
LIB_PATH = os.path.join(_HERE, "libhmx.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
_HEADERS = ("hmx_platform.cuh", "hmx_cell_common.cuh", "hmx_cell_poisson.cuh", "hmx_cell_coarse.cuh", "hmx_cell_elasticity.cuh",
            "hmx_cell_dense.cuh", "hmx_cell_cluster.cuh", "hmx_cell_generic.cuh", "hmx_cell_entry.cu")
MATRIX_FREE, DENSE, CLUSTER, ELEMENT_LIST = 0, 3, 4, 5  # (1, 2: the slower assembled-operator experiments, experiments/)
DENSE_MAX_DOF = 192  # register tile of the dense Cholesky kernel: 12 x 12 blocks of 16 x 16 threads
SMEM_LIMIT = 227 * 1024
ABI_VERSION = 8  # HMX_ABI_VERSION of include/hmx.h these bindings were written for


class HmxError(RuntimeError):
    pass


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise HmxError("nvcc not found: the cell kernels cannot be built (there is no CPU fallback)")
    return exe


# ----------------------------------------------------------------------------
# libhmx.so
# ----------------------------------------------------------------------------
def build_library(force=False, verbose=False):
    """Compile hommx_b200/csrc/hmx_lib.cu into hommx_b200/libhmx.so (in-tree, so that it
    travels to the GPU box)."""
    src = os.path.join(CSRC, "hmx_lib.cu")
    deps = [src, os.path.join(INCLUDE, "hmx.h")] + [os.path.join(CSRC, h) for h in _HEADERS[:2]]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = [_nvcc(), *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl",
           "-o", LIB_PATH, src]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise HmxError(f"building libhmx.so failed:\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB_PATH


class hmx_micro_mesh(C.Structure):
    _fields_ = [
        ("n_elem", C.c_int32), ("n_nodes", C.c_int32), ("nnzb", C.c_int32),
        ("elem_nodes", C.c_void_p), ("elem_grad", C.c_void_p), ("elem_vol", C.c_void_p), ("elem_yq", C.c_void_p),
        ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("blk_ptr", C.c_void_p), ("blk_src", C.c_void_p),
        ("node_ptr", C.c_void_p), ("node_src", C.c_void_p), ("diag", C.c_void_p),
    ]  # fmt: skip

    @classmethod
    def from_tables(cls, t):
        """``t``: hommx_b200.micro.ElementListTables (its arrays must outlive the call that uses the struct)."""
        p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        return cls(t.n_elem, t.n_nodes, t.nnzb, p(t.elem_nodes), p(t.elem_grad), p(t.elem_vol), p(t.elem_yq), p(t.row_ptr),
                   p(t.col), p(t.blk_ptr), p(t.blk_src), p(t.node_ptr), p(t.node_src), p(t.diag))  # fmt: skip


class hmx_desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("kind", C.c_int32), ("n_micro", C.c_int32), ("nq", C.c_int32),
        ("qp", C.POINTER(C.c_double)), ("qw", C.POINTER(C.c_double)),
        ("kernel_image", C.c_void_p), ("kernel_image_size", C.c_size_t),
        ("rtol", C.c_double), ("atol", C.c_double), ("max_it", C.c_int32), ("device", C.c_int32),
        ("micro_mesh", C.POINTER(hmx_micro_mesh)),
    ]  # fmt: skip


_lib = None
_lib_lock = threading.Lock()

# name -> (restype, argtypes); every symbol include/hmx.h declares
SYMBOLS = {
    "hmx_abi_version": (C.c_int32, []),
    "hmx_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(hmx_desc)]),
    "hmx_destroy": (None, [C.c_void_p]),
    "hmx_last_error": (C.c_char_p, [C.c_void_p]),
    "hmx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hmx_set_tolerances": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int32]),
    "hmx_set_grid": (C.c_int, [C.c_void_p, C.c_int32]),
    "hmx_kernel_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "hmx_cell_tensors": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_cell_tensors_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_cell_correctors_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_assemble_macro": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_assemble_macro_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_gather_csr_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_cluster_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "hmx_rhs_iterations": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "hmx_halo_pack_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hmx_halo_unpack_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hmx_halo_sum_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hmx_macro_elements_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmx_macro_load_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "hmx_macro_lift_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "hmx_macro_pcg_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "hmx_measure_peaks": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "hmx_sync": (C.c_int, [C.c_void_p]),
}  # fmt: skip


def load_library():
    """dlopen hommx_b200/libhmx.so.  With nvcc on the host the library is (re)built first whenever it is missing or
    older than its sources; without nvcc the shipped binary is used.  Either way its ABI version must be the one
    these bindings were written for: a stale library is refused instead of being called with the wrong arguments."""
    global _lib
    with _lib_lock:
        if _lib is None:
            have_nvcc = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
            if have_nvcc or not os.path.exists(LIB_PATH):
                build_library()  # no-op when the library is newer than its sources
            lib = C.CDLL(LIB_PATH)
            try:
                lib.hmx_abi_version.restype = C.c_int32
                got = int(lib.hmx_abi_version())
            except AttributeError:
                got = -1
            if got != ABI_VERSION:
                raise HmxError(f"{LIB_PATH} has ABI version {got}, these bindings need {ABI_VERSION}: rebuild it "
                               "(python -c 'import __graft_entry__ as g; g.build()')")
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


# ----------------------------------------------------------------------------
# cell kernels
# ----------------------------------------------------------------------------
def dense_dofs(prog, n, coll=0):
    d = prog.dim
    return d * n ** (d - bin(coll & ((1 << d) - 1)).count("1"))


def dense_fits(prog, n, coll=0):
    """The direct (dense Cholesky) elasticity kernel: all unknowns of the cell in a 192 x 192 register tile."""
    return prog.kind != POISSON and prog.dim + 1 <= dense_dofs(prog, n, coll) <= DENSE_MAX_DOF


def default_variant(prog, n, collapse=False):
    """Elasticity: the matrix-free PCG kernel; ``HMX_ELASTICITY_VARIANT=dense`` forces the direct kernel where the
    cell fits it (by default the drop-in classes decide from measured PCG iterations: hmm.py, cell_solver="auto")."""
    if prog.kind == POISSON:
        return MATRIX_FREE
    coll = collapse_mask(prog, collapse)
    if os.environ.get("HMX_ELASTICITY_VARIANT") == "dense" and dense_fits(prog, n, coll):
        return DENSE
    # cells that exceed one SM (3-D, n >= 10: the matrix-free kernel would keep its vectors in L2): the assembled
    # stencil resident in the distributed shared memory of a thread-block cluster, where a portable cluster (<= 8
    # CTAs) holds it (measured on B200, 10^3 fibre cell: 7.9k against 5.2k-5.8k cell solves/s; profiles/r02_cluster.md)
    # ... and 8^3 cells whose coefficient varies along all three micro axes: no coarse space fits either kernel there
    # (block Jacobi, ~370 iterations for a stiff ball), and the cheaper operator wins (34.9k against 29.3k cell solves/s)
    if prog.dim == 3 and coll == 0 and os.environ.get("HMX_ELASTICITY_VARIANT") != "matrix-free":
        big = vectors_in_l2(prog, n, 0)
        no_coarse = n >= 8 and n % 2 == 0 and coarse_dofs(prog, n, MATRIX_FREE, 0) == 0 and precond_mode(prog, MATRIX_FREE) == 1
        if (big or no_coarse) and 2 <= cluster_size(prog, n) <= 8:
            return CLUSTER
    return MATRIX_FREE


def collapse_mask(prog, collapse=True):
    """Bit mask of the micro axes the coefficient does not depend on (they can be collapsed exactly)."""
    if not collapse:
        return 0
    return ~prog.ydep & ((1 << prog.dim) - 1)


def cluster_tpn():
    """Threads per node of the cluster kernel (2: each thread owns half of the right-hand sides)."""
    return int(os.environ.get("HMX_TPN", "2"))


def cluster_coarse_dofs(prog, n, cl, threads=512):
    """Unknowns of the cluster kernel's coarse space (mirrors ``ClusterLayout::TWO`` in csrc/hmx_cell_cluster.cuh): the
    level-1 space summed along micro axis 0 when the coefficient does not depend on it; 0 = block Jacobi only."""
    if os.environ.get("HMX_PRECOND", "twolevel") == "jacobi" or n % 2 or n < 4 or threads // 6 < 32:
        return 0
    h = n // 2
    if 3 * h**3 <= 96 or (prog.ydep & 1) or 3 * h * h > 96:
        return 0
    lanes = cluster_tpn() * n  # line sums by warp shuffles: an x-line is an aligned power-of-two segment of a warp
    if (lanes > 32 or lanes & (lanes - 1)) and "-DHMX_CLUSTER_TWO_ANY=1" not in _extra_flags():
        return 0  # (10^3: measured slower than block Jacobi with the generic line sums, csrc/hmx_cell_cluster.cuh)
    return 3 * h * h


def cluster_threads(prog, n, cl):
    own = (n // cl) * n * n * cluster_tpn()
    nt = 32 * (-(-own // 32))
    ncd = cluster_coarse_dofs(prog, n, cl)
    return max(nt, 64, 32 * (-(-(6 * ncd + 6) // 32)) if ncd else 0)


def cluster_smem_bytes(prog, n, cl):
    """Shared memory per CTA of the cluster kernel (mirrors ``ClusterLayout``)."""
    nt = cluster_threads(prog, n, cl)
    pz, npl = n // cl, n * n
    nown, npb = pz * npl, (pz + 2) * npl
    ncd = cluster_coarse_dofs(prog, n, cl, nt) or 2
    two = cluster_coarse_dofs(prog, n, cl, nt) > 0
    nrec = (max(6 * ncd + 6, 36) + 1) // 2 * 2
    nblk = (ncd + 31) // 32
    cbuf = 0
    if two:
        pad = max((ncd + 15) // 16 * 16, -(-ncd // (nt // 16)) * (nt // 16))
        cbuf = max(3 * pad + 2, 6 * ncd) + 2
    work = max(pz * n * 18, 6 * ncd + 6 * nblk + 2, cbuf, (nt // 36) * 36, 110)
    ntri = ncd * (ncd + 1) // 2 if two else 0
    doubles = 12 + (nt // 32) * 8 + cl * 8 + cl * nrec + 8 + 6 * ncd + work + 1 + ntri + 1 + npb * 18 + 63 * nown + 36 * npl
    if two and 8 * doubles > SMEM_LIMIT:  # the inverse coarse matrix moves to the global scratch (ClusterLayout::EIG)
        doubles -= ntri
    return 8 * doubles


def cluster_size(prog, n):
    """CTAs per cluster of the cluster kernel: the smallest split of the cell into z-slabs whose share of the stencil
    fits in 227 KB (8^3: 2, 10^3: 5); ``HMX_CLUSTER_SIZE`` overrides.  0: the cell cannot be split (n has no divisor
    <= 16 that fits)."""
    forced = os.environ.get("HMX_CLUSTER_SIZE")
    if forced:
        return int(forced)
    for cl in range(2, 17):
        if n % cl == 0 and cluster_threads(prog, n, cl) <= 1024 and cluster_smem_bytes(prog, n, cl) <= SMEM_LIMIT:
            return cl
    return 0


def default_threads(dim, kind, n, variant=MATRIX_FREE, coll=0, prog=None):
    """Threads per CTA for the cell kernel of an n^dim micro mesh (minus its collapsed axes).  ``prog`` (optional)
    lets the Poisson choice see how much shared memory the atoms of the coefficient take."""
    N = n ** (dim - bin(coll).count("1"))
    if kind != POISSON and variant == DENSE:
        return 256  # 16 x 16 owners of the register tile
    if kind == POISSON:
        # four nodes per thread and several CTAs per SM beat two nodes per thread and one CTA
        # (measured on B200, scripts/probe_occ.py: C3 12.2M -> 19.0M, C2 7.0M -> 11.6M points/s)
        nt = -(-N // 4)
        nt = max(64, min(1024, 32 * (-(-nt // 32))))
        if prog is not None and not vectors_in_l2(prog, n, coll):
            # many atoms on a fully y-dependent coefficient: shared memory, not registers, limits the CTAs per SM, and
            # the atom means (one element per thread) dominate -- 16 warps per SM at least (measured, 4 atoms at 8^3,
            # one CTA per SM: 128 threads 133k, 256 threads 236k, 512 threads 338k cell solves/s)
            ndep = bin(prog.ydep & ((1 << dim) - 1)).count("1")
            atoms = max(1, prog.natoms) * (2 if dim == 2 else 6) * n**ndep
            smem = 8 * ((2**dim - 1) * N + max(dim * N, atoms) + 512)
            ctas = max(1, SMEM_LIMIT // smem)
            if atoms > dim * N and ctas * nt < 512 and ctas <= 2:
                return min(512, 32 * (-(-512 // (32 * ctas))))
        if coll == 0 and dim >= 2:
            # ... and a count with which the kernel can give every thread a piece of a grid line (PoissonLayout
            # TILED: n divisible by the nodes per thread; the line pieces rounded up to whole warps).  At most 512
            # threads: more nodes per thread at >= 128 registers beat more threads that spill (measured, 48^2: 384
            # threads x 6 nodes 2.49M, 576 x 4 1.58M-1.78M cell solves/s; 64^2: 512 x 8 0.73M-0.75M, 1024 x 4 0.56M-0.58M;
            # 5 nodes per thread at 10^3: 0.86x of 256 threads without line pieces)
            for npt in (4, 3, 6, 8):
                if n % npt == 0:
                    cand = 32 * (-(-(N // n) * (n // npt) // 32))
                    if 64 <= cand <= 512 and -(-N // cand) == npt:
                        return cand
        return nt
    nrhs = dim * (dim + 1) // 2
    ncol = 2 ** (dim - bin(coll).count("1"))
    per = -(-N // ncol)  # cubes of one colour (even n)
    if per <= 16 and nrhs % 2 == 0:
        return 16 * nrhs  # two right-hand sides share a warp (ElasticityLayout::SUBW)
    tpr = max(32, min(64, 32 * (-(-per // 32))))
    if dim == 3 and coll == 0 and n % 2 == 0 and (n // 2) % (tpr // 32) != 0:
        tpr = 32  # the 2x2x1 block sweep needs whole last-axis planes per warp (measured, n = 10: 4.7k -> 5.8k cells/s)
    return tpr * nrhs


def _src_hash():
    hsh = hashlib.sha1()
    for name in _HEADERS:
        p = os.path.join(CSRC, name)
        if os.path.exists(p):
            with open(p, "rb") as f:
                hsh.update(f.read())
    hsh.update(" ".join(_extra_flags()).encode())
    return hsh.hexdigest()[:12]


def _extra_flags():
    """Extra nvcc flags for kernel experiments (``HMX_EXTRA_NVCC="-DHMX_NO_BLOCK_SWEEP"``); part of the cache key."""
    return os.environ.get("HMX_EXTRA_NVCC", "").split()


def default_min_blocks(dim, kind, n, threads, coll=0, vglob=0):
    """__launch_bounds__ minimum of resident CTAs per SM (caps registers per thread).  The Poisson
    kernels run few PCG iterations with ~25 barriers per point: two or more CTAs per SM hide them."""
    if kind != POISSON and dim == 3 and coll == 0 and threads >= 32 * 6:
        # full 3-D elasticity cells: the sweep wants 168 registers per thread (n = 6 at 192 threads, measured:
        # 3 CTAs/SM at 96 registers 119k, 2 at 168 registers 183k, 1 at 254 registers 167k cell solves/s);
        # with the vectors in L2 a second CTA only adds L2 contention (n = 10: 5.8k vs 4.5k)
        return 1 if vglob else max(1, 65536 // (threads * 168))
    if kind == POISSON and dim == 3 and coll == 0 and threads <= 128:
        # full 3-D Poisson cells with 4+ nodes per thread: the line-tiled K p keeps NPT + 1 neighbour values of three
        # right-hand sides in flight; measured at 8^3 / 128 threads: 3 CTAs at 168 registers 24.3M, 4 at 128 registers
        # (spilling) 22.7M cell solves/s; 64 threads: 6 CTAs 18.4M, 8 CTAs 16.6M
        return max(1, min(8, 65536 // (threads * 168)))
    # (elasticity: the small / axis-collapsed kernels; measured on the collapsed C4 kernel: 215k -> 276k cell
    # solves/s going from 2 to 4-5 CTAs per SM)
    return max(1, min(8, 65536 // (threads * 112)))


def vectors_in_l2(prog, n, coll=0):
    """Matrix-free elasticity: 1 when p and y = K p of all right-hand sides do not fit in shared memory next
    to the preconditioner and the atoms (3-D, n >= 10): they then live in the L2 scratch."""
    if os.environ.get("HMX_FORCE_VGLOB") == "1":  # tests: exercise the fallback on small cells
        return 1
    d = prog.dim
    if prog.kind == POISSON:  # Poisson: the atoms move out when they do not fit next to the half stencil
        N = n ** (d - bin(coll).count("1"))
        ndep = bin(prog.ydep & ((1 << d) - 1)).count("1")
        atoms = max(1, prog.natoms) * (2 if d == 2 else 6) * n**ndep
        need = 8 * ((2**d - 1) * N + max(d * N, atoms) + 4096)
        return 1 if need > SMEM_LIMIT else 0
    ext = [1 if (coll >> a) & 1 else n for a in range(d)]
    slots = 2**d * int(np.prod([(e + 1) // 2 for e in ext]))
    nrhs = d * (d + 1) // 2
    ndep = bin(prog.ydep & ((1 << d) - 1)).count("1")
    atoms = max(1, prog.natoms) * (2 if d == 2 else 6) * (2**ndep) * ((n + 1) // 2) ** ndep
    need = 8 * (2 * nrhs * d * slots + nrhs * slots + atoms + 2048)
    return 1 if need > SMEM_LIMIT else 0


def precond_mode(prog, variant=MATRIX_FREE):
    """1: the matrix-free elasticity kernel is built with the additive two-level preconditioner (block Jacobi + an
    exactly inverted Galerkin coarse matrix, csrc/hmx_cell_coarse.cuh) wherever its coarse space fits next to the
    vectors (the kernel decides: ``CoarseSpace::ON``); 0: block Jacobi only (``HMX_PRECOND=jacobi``)."""
    if prog.kind == POISSON or variant not in (MATRIX_FREE, CLUSTER):
        return 0
    return 0 if os.environ.get("HMX_PRECOND", "twolevel") == "jacobi" else 1


def coarse_dofs(prog, n, variant=MATRIX_FREE, coll=0):
    """Unknowns of the coarse space the two-level PCG kernel builds for this coefficient and micro mesh (mirrors
    ``CoarseSpace`` in csrc/hmx_cell_coarse.cuh, up to its shared-memory fit test); 0 = block Jacobi only."""
    d = prog.dim
    if not precond_mode(prog, variant) or coll or n % 2 or n < 4 or vectors_in_l2(prog, n, coll):
        return 0
    h = n // 2
    nodes = h**d
    if d * nodes > 96:  # the level-1 space does not fit: sum it up along an axis the coefficient does not depend on
        if prog.ydep == (1 << d) - 1:
            return 0
        nodes //= h
    return d * nodes if d * nodes <= 96 else 0


def kernel_key(prog: CoefficientProgram, n, threads, min_blocks=1, variant=MATRIX_FREE, coll=0):
    kind = "p" if prog.kind == POISSON else "e"
    vg = vectors_in_l2(prog, n, coll) if variant == MATRIX_FREE else 0
    cl = f"k{cluster_size(prog, n)}x{cluster_tpn()}" if variant == CLUSTER else ""
    return f"{kind}{prog.dim}_n{n}_t{threads}b{min_blocks}v{variant}{cl}c{coll}g{vg}p{precond_mode(prog, variant)}_{prog.key}_{_src_hash()}"


def kernel_defines(prog, n, threads, coeff_path, min_blocks=1, variant=MATRIX_FREE, coll=0):
    vg = vectors_in_l2(prog, n, coll) if variant == MATRIX_FREE else 0
    cl = [f"-DHMX_CLUSTER={cluster_size(prog, n)}", f"-DHMX_TPN={cluster_tpn()}"] if variant == CLUSTER else []
    return [f'-DHMX_COEFF_FILE="{coeff_path}"', f"-DHMX_KIND={prog.kind}", f"-DHMX_NM={n}", f"-DHMX_NT={threads}",
            f"-DHMX_MINB={min_blocks}", f"-DHMX_VARIANT={variant}", f"-DHMX_COLL={coll}", f"-DHMX_VGLOB={vg}",
            f"-DHMX_PRECOND={precond_mode(prog, variant)}", *cl]  # fmt: skip


def resolve(prog, n, threads=None, min_blocks=None, variant=None, collapse=False):
    variant = default_variant(prog, n, collapse) if variant is None else variant
    coll = collapse_mask(prog, collapse) if variant in (MATRIX_FREE, DENSE) else 0
    if variant == DENSE and not dense_fits(prog, n, coll):
        raise HmxError(f"the dense variant holds at most {DENSE_MAX_DOF} unknowns per cell")
    if variant == ELEMENT_LIST:  # general periodic micro mesh: the mesh is run-time data (n = 0)
        return threads or 256, min_blocks or 2, variant, 0
    if variant == CLUSTER:
        if prog.kind == POISSON or prog.dim != 3:
            raise HmxError("the cluster variant is the 3-D elasticity kernel")
        cl = cluster_size(prog, n)
        if cl < 2 or n % cl:
            raise HmxError(f"an {n}^3 cell cannot be split into z-slabs over a thread-block cluster")
        return threads or cluster_threads(prog, n, cl), 1, variant, 0
    threads = threads or default_threads(prog.dim, prog.kind, n, variant, coll, prog)
    vg = vectors_in_l2(prog, n, coll) if variant == MATRIX_FREE else 0
    min_blocks = min_blocks or (1 if variant == DENSE else default_min_blocks(prog.dim, prog.kind, n, threads, coll, vg))
    return threads, min_blocks, variant, coll


def compile_kernel(prog: CoefficientProgram, n, threads=None, force=False, keep_log=True, min_blocks=None, variant=None,
                   collapse=False):
    """nvcc -cubin of the cell kernel for this coefficient program; returns the cubin path."""
    threads, min_blocks, variant, coll = resolve(prog, n, threads, min_blocks, variant, collapse)
    os.makedirs(KCACHE, exist_ok=True)
    key = kernel_key(prog, n, threads, min_blocks, variant, coll)
    cubin = os.path.join(KCACHE, key + ".cubin")
    if os.path.exists(cubin) and not force:
        return cubin
    # every rank of a torchrun launch may compile the same key at once on a cold cache: each process writes its own
    # temporary files (coefficient header, cubin, log) and renames them into place, so nobody reads a half-written one
    uniq = f".tmp{os.getpid()}_{threading.get_ident()}"
    coeff = os.path.join(KCACHE, key + ".coeff.cuh")
    coeff_tmp = os.path.join(KCACHE, key + uniq + ".coeff.cuh")
    with open(coeff_tmp, "w") as f:
        f.write(prog.source)
    tmp = cubin + uniq
    cmd = [_nvcc(), *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-cubin", "-Xptxas", "-v", "-I", CSRC,
           *kernel_defines(prog, n, threads, coeff_tmp, min_blocks, variant, coll), *_extra_flags(), "-o", tmp, os.path.join(CSRC, "hmx_cell_entry.cu")]  # fmt: skip
    try:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise HmxError(f"nvcc failed for cell kernel {key}:\n{r.stderr[-4000:]}")
        if keep_log:
            log_tmp = os.path.join(KCACHE, key + uniq + ".ptxas.log")
            with open(log_tmp, "w") as f:
                f.write(" ".join(cmd).replace(coeff_tmp, coeff) + "\n" + r.stderr)
            os.replace(log_tmp, os.path.join(KCACHE, key + ".ptxas.log"))
        os.replace(coeff_tmp, coeff)
        os.replace(tmp, cubin)
    finally:
        for leftover in (coeff_tmp, tmp):
            if os.path.exists(leftover):
                os.remove(leftover)
    return cubin


def compile_load_kernel(lprog, force=False):
    """nvcc -cubin of csrc/hmx_load_entry.cu for the traced right-hand side ``lprog`` (codegen.LoadProgram); returns the
    cubin image (bytes).  Cached in-tree like the cell kernels; NVRTC on hosts without nvcc."""
    os.makedirs(KCACHE, exist_ok=True)
    with open(os.path.join(CSRC, "hmx_load_entry.cu"), "rb") as f:
        entry = f.read()
    key = f"load{lprog.dim}_bs{lprog.bs}_{lprog.key}_{hashlib.sha1(entry).hexdigest()[:12]}"
    cubin = os.path.join(KCACHE, key + ".cubin")
    if not os.path.exists(cubin) or force:
        have_nvcc = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
        uniq = f".tmp{os.getpid()}_{threading.get_ident()}"
        if have_nvcc and os.environ.get("HMX_COMPILER", "") != "nvrtc":
            hdr = os.path.join(KCACHE, key + uniq + ".load.cuh")
            with open(hdr, "w") as f:
                f.write(lprog.source)
            cmd = [_nvcc(), *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-cubin", f'-DHMX_LOAD_FILE="{hdr}"', "-o", cubin + uniq,
                   os.path.join(CSRC, "hmx_load_entry.cu")]  # fmt: skip
            try:
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise HmxError(f"nvcc failed for the load kernel:\n{r.stderr[-3000:]}")
                os.replace(hdr, os.path.join(KCACHE, key + ".load.cuh"))
                os.replace(cubin + uniq, cubin)
            finally:
                for leftover in (hdr, cubin + uniq):
                    if os.path.exists(leftover):
                        os.remove(leftover)
        else:
            image = _nvrtc_compile(entry.decode(), "hmx_load_entry.cu", [lprog.source.encode()], [b"hmx_load_program.cuh"],
                                   ['-DHMX_LOAD_FILE="hmx_load_program.cuh"'])
            with open(cubin + uniq, "wb") as f:
                f.write(image)
            os.replace(cubin + uniq, cubin)
    with open(cubin, "rb") as f:
        return f.read()


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class CellSolver:
    """One ``hmx_t`` handle: a (problem class, coefficient, micro mesh, quadrature) combination.

    ``qp`` has shape (T, nq, dim) (cube-local quadrature points per element type, in units of
    h) and ``qw`` (nq,) normalised weights -- built by ``hommx_b200.micro.quadrature_table``.
    """

    def __init__(self, prog: CoefficientProgram, n_micro, qp, qw, rtol=1e-8, atol=1e-10, max_it=10000, device=0, threads=None,
                 min_blocks=None, variant=None, collapse=False, micro_tables=None):
        """``micro_tables`` (hommx_b200.micro.ElementListTables): a general periodic micro mesh -- the element-list
        kernel (variant ELEMENT_LIST) is used, ``n_micro`` and ``qp`` are ignored."""
        self.prog = prog
        self.device = int(device)
        self.micro_tables = micro_tables
        if micro_tables is not None:
            n_micro, variant, collapse = 0, ELEMENT_LIST, False
            qw = micro_tables.qw
            qp = np.zeros((2 if prog.dim == 2 else 6, len(qw), prog.dim))
        elif variant == ELEMENT_LIST:
            raise ValueError("the element-list kernel needs micro_tables")
        self.dim, self.kind, self.n = prog.dim, prog.kind, int(n_micro)
        self.m = prog.n_rhs
        self.nb = (self.dim + 1) * (1 if self.kind == POISSON else self.dim)
        self._h = C.c_void_p()
        self.lib = load_library()
        image = kernel_image(prog, self.n, threads, min_blocks, variant, collapse)
        self.variant = resolve(prog, self.n, threads, min_blocks, variant, collapse)[2]
        self.collapse_mask = collapse_mask(prog, collapse) if self.variant in (MATRIX_FREE, DENSE) else 0
        self._image = C.create_string_buffer(image, len(image))
        qp = np.ascontiguousarray(qp, dtype=np.float64)
        qw = np.ascontiguousarray(qw, dtype=np.float64)
        T = 2 if self.dim == 2 else 6
        if qp.shape != (T, len(qw), self.dim):
            raise ValueError(f"quadrature points must have shape {(T, len(qw), self.dim)}, got {qp.shape}")
        mm = hmx_micro_mesh.from_tables(micro_tables) if micro_tables is not None else None
        d = hmx_desc(self.dim, self.kind, self.n, len(qw), qp.ctypes.data_as(C.POINTER(C.c_double)),
                     qw.ctypes.data_as(C.POINTER(C.c_double)), C.cast(self._image, C.c_void_p), len(image),
                     rtol, atol, max_it, device, C.pointer(mm) if mm is not None else None)  # fmt: skip
        rc = self.lib.hmx_create(C.byref(self._h), C.byref(d))
        if rc != 0:
            msg = self.lib.hmx_last_error(None).decode()
            self._h = C.c_void_p()
            if rc == -1:
                raise ValueError(msg)
            raise HmxError(f"hmx_create failed ({rc}): {msg}")
        info = (C.c_int32 * 8)()
        self._check(self.lib.hmx_kernel_info(self._h, info))
        self.info = dict(zip(("smem_bytes", "threads", "n_rhs", "m", "n_b", "ctas_per_sm", "sms", "scratch_doubles"), info))
        cinfo = (C.c_int32 * 2)()
        self._check(self.lib.hmx_cluster_info(self._h, cinfo))
        self.info["cluster"], self.info["resident_clusters"] = int(cinfo[0]), int(cinfo[1])

    # -- plumbing ----------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.hmx_last_error(self._h).decode()
            if rc == -1:
                raise ValueError(msg)
            raise HmxError(f"libhmx error {rc}: {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.hmx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._check(self.lib.hmx_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_tolerances(self, rtol, atol, max_it=0):
        self._check(self.lib.hmx_set_tolerances(self._h, rtol, atol, max_it))

    def set_grid(self, n_ctas):
        self._check(self.lib.hmx_set_grid(self._h, int(n_ctas)))

    def sync(self):
        self._check(self.lib.hmx_sync(self._h))

    # -- host-buffer entry points ------------------------------------------------------
    def cell_tensors(self, x_pts, return_stats=False):
        x = np.ascontiguousarray(np.asarray(x_pts, dtype=np.float64).reshape(-1, 3))
        n = len(x)
        A = np.empty((n, self.m, self.m))
        it = np.empty(n, dtype=np.int32)
        res = np.empty(n)
        self._check(self.lib.hmx_cell_tensors(self._h, n, _ptr(x), _ptr(A), _ptr(it), _ptr(res)))
        return (A, it, res) if return_stats else A

    def assemble_macro(self, cell_nodes, node_xyz, gather_ptr, gather_src, want_local=False, return_stats=False):
        cells = np.ascontiguousarray(cell_nodes, dtype=np.int32).reshape(-1, self.dim + 1)
        xyz = np.ascontiguousarray(node_xyz, dtype=np.float64).reshape(-1, 3)
        gp = np.ascontiguousarray(gather_ptr, dtype=np.int64)
        gs = np.ascontiguousarray(gather_src, dtype=np.int32)
        nnz = len(gp) - 1
        nc = len(cells)
        vals = np.empty(nnz)
        S = np.empty((nc, self.nb, self.nb)) if want_local else None
        it = np.empty(nc, dtype=np.int32)
        res = np.empty(nc)
        self._check(self.lib.hmx_assemble_macro(self._h, nc, _ptr(cells), len(xyz), _ptr(xyz), nnz, _ptr(gp), _ptr(gs),
                                                _ptr(vals), _ptr(S), _ptr(it), _ptr(res)))  # fmt: skip
        out = (vals,)
        if want_local:
            out += (S,)
        if return_stats:
            out += (it, res)
        return out if len(out) > 1 else vals

    # -- device-pointer entry points (torch tensors or raw addresses) -----------------------
    @staticmethod
    def _dp(t):
        if t is None:
            return None
        return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))

    def cell_tensors_dev(self, n_pts, x_pts, A_hom, iters=None, resid=None):
        dp = self._dp
        self._check(self.lib.hmx_cell_tensors_dev(self._h, int(n_pts), dp(x_pts), dp(A_hom), dp(iters), dp(resid)))

    def cell_correctors(self, x_pts):
        """(A_hom (n, m, m), chi (n, n_rhs, bs, *grid)) with the grid in (z, y, x) order; collapsed axes are
        broadcast back to the full n^d periodic grid."""
        import torch

        x = np.ascontiguousarray(np.asarray(x_pts, dtype=np.float64).reshape(-1, 3))
        n, d = len(x), self.dim
        bs = 1 if self.kind == POISSON else d
        shape = [1 if (self.collapse_mask >> a) & 1 else self.n for a in range(d)]
        if self.micro_tables is not None:  # general mesh: values on the periodic nodes, (n, n_rhs, bs, n_nodes)
            shape = [self.micro_tables.n_nodes]
        dev = torch.device("cuda", self.device)  # the handle's device, whatever the caller's current device is
        with torch.cuda.device(dev):
            # the zero-fill and the uploads are torch work: the kernel must run on the same stream (or after them)
            self.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            xd = torch.as_tensor(x, device=dev)
            A = torch.empty((n, self.m, self.m), dtype=torch.float64, device=dev)
            chi = torch.zeros((n, self.m, bs) + tuple(reversed(shape)), dtype=torch.float64, device=dev)
            self._check(self.lib.hmx_cell_correctors_dev(self._h, n, self._dp(xd), self._dp(A), self._dp(chi)))
            self.sync()
            if self.micro_tables is not None:
                return A.cpu().numpy(), chi.cpu().numpy()
            full = (n, self.m, bs) + (self.n,) * d
            return A.cpu().numpy(), np.broadcast_to(chi.cpu().numpy(), full).copy()

    def assemble_macro_dev(self, n_cells, cell_nodes, n_nodes, node_xyz, nnz, gather_ptr, gather_src, csr_vals, S_loc=None,
                           iters=None, resid=None):
        dp = self._dp
        self._check(self.lib.hmx_assemble_macro_dev(self._h, int(n_cells), dp(cell_nodes), int(n_nodes), dp(node_xyz), int(nnz),
                                                    dp(gather_ptr), dp(gather_src), dp(csr_vals), dp(S_loc), dp(iters),
                                                    dp(resid)))  # fmt: skip

    def gather_csr_dev(self, nnz, gather_ptr, gather_src, S_loc, csr_vals):
        dp = self._dp
        self._check(self.lib.hmx_gather_csr_dev(self._h, int(nnz), dp(gather_ptr), dp(gather_src), dp(S_loc), dp(csr_vals)))

    def rhs_iterations(self, reset=True):
        """Total PCG iterations (summed over points and right-hand sides) since the last reset."""
        v = C.c_int64()
        self._check(self.lib.hmx_rhs_iterations(self._h, C.byref(v), 1 if reset else 0))
        return v.value

    def halo_sum_dev(self, nccl_comm, csr_vals, slots, n, buf):
        """pack + ncclAllReduce + unpack on the handle's stream; ``nccl_comm`` is a raw ncclComm_t (an int address)."""
        dp = self._dp
        self._check(self.lib.hmx_halo_sum_dev(self._h, C.c_void_p(int(nccl_comm)), dp(csr_vals), dp(slots), int(n), dp(buf)))

    def macro_elements_dev(self, n_cells, cell_nodes, node_xyz, weights, A_pts, S_loc):
        """S_loc of every macro cell from tensors at ``len(weights)`` macro quadrature points per cell."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        dp = self._dp
        self._check(self.lib.hmx_macro_elements_dev(self._h, int(n_cells), dp(cell_nodes), dp(node_xyz), len(w), _ptr(w),
                                                    dp(A_pts), dp(S_loc)))  # fmt: skip

    def macro_load_dev(self, image, n_cells, cell_nodes, node_xyz, qp, qw, Fe):
        """Element load vectors Fe [n_cells][(dim+1)*bs] of the right-hand side compiled into ``image``."""
        qp = np.ascontiguousarray(qp, dtype=np.float64)
        qw = np.ascontiguousarray(qw, dtype=np.float64)
        buf = C.create_string_buffer(image, len(image))
        dp = self._dp
        self._check(self.lib.hmx_macro_load_dev(self._h, C.cast(buf, C.c_void_p), len(image), int(n_cells), dp(cell_nodes), dp(node_xyz),
                                                len(qw), _ptr(qp), _ptr(qw), dp(Fe)))  # fmt: skip

    def macro_lift_dev(self, n_dofs, indptr, indices, csr_vals, bc_mask, bc_values, b):
        dp = self._dp
        self._check(self.lib.hmx_macro_lift_dev(self._h, int(n_dofs), dp(indptr), dp(indices), dp(csr_vals), dp(bc_mask),
                                                dp(bc_values), dp(b)))  # fmt: skip

    def macro_pcg_dev(self, n_dofs, indptr, indices, csr_vals, b, x, rtol=1e-12, atol=0.0, max_it=20000):
        dp = self._dp
        it, res = C.c_int32(), C.c_double()
        self._check(self.lib.hmx_macro_pcg_dev(self._h, int(n_dofs), dp(indptr), dp(indices), dp(csr_vals), dp(b), dp(x),
                                               rtol, atol, max_it, C.byref(it), C.byref(res)))  # fmt: skip
        return it.value, res.value

    def halo_pack_dev(self, csr_vals, slots, n, buf):
        dp = self._dp
        self._check(self.lib.hmx_halo_pack_dev(self._h, dp(csr_vals), dp(slots), int(n), dp(buf)))

    def halo_unpack_dev(self, csr_vals, slots, n, buf):
        dp = self._dp
        self._check(self.lib.hmx_halo_unpack_dev(self._h, dp(csr_vals), dp(slots), int(n), dp(buf)))


def measure_peaks(device=0):
    """(FP64 DFMA TFLOP/s, device copy GB/s) measured on `device` by libhmx's microbenchmarks."""
    lib = load_library()
    f, c = C.c_double(), C.c_double()
    rc = lib.hmx_measure_peaks(device, C.byref(f), C.byref(c))
    if rc != 0:
        raise HmxError(lib.hmx_last_error(None).decode())
    return f.value, c.value


# ----------------------------------------------------------------------------
# NVRTC: the same kernel translation unit compiled in-process (hosts without nvcc)
# ----------------------------------------------------------------------------
def _nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    try:  # the CUDA runtime wheels torch depends on ship one as well
        import nvidia.cuda_nvrtc as pkg

        lib_dir = os.path.join(os.path.dirname(pkg.__file__), "lib")
        for f in sorted(os.listdir(lib_dir)):
            if f.startswith("libnvrtc.so"):
                return C.CDLL(os.path.join(lib_dir, f))
    except Exception:
        pass
    raise HmxError("neither nvcc nor libnvrtc found: the cell kernels cannot be built (there is no CPU fallback)")


def _nvrtc_compile(src, name, headers, header_names, defines):
    """cubin image of one translation unit compiled in-process with NVRTC for sm_100a."""
    rt = _nvrtc()
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-lineinfo"] + [d.encode() for d in defines]
    prg = C.c_void_p()
    rc = rt.nvrtcCreateProgram(C.byref(prg), src.encode(), name.encode(), len(headers),
                               (C.c_char_p * len(headers))(*headers), (C.c_char_p * len(header_names))(*header_names))  # fmt: skip
    if rc != 0:
        raise HmxError(f"nvrtcCreateProgram failed ({rc})")
    try:
        rc = rt.nvrtcCompileProgram(prg, len(opts), (C.c_char_p * len(opts))(*opts))
        if rc != 0:
            size = C.c_size_t()
            rt.nvrtcGetProgramLogSize(prg, C.byref(size))
            log = C.create_string_buffer(size.value)
            rt.nvrtcGetProgramLog(prg, log)
            raise HmxError(f"NVRTC failed for {name}:\n{log.value.decode()[-4000:]}")
        size = C.c_size_t()
        if rt.nvrtcGetCUBINSize(prg, C.byref(size)) != 0 or size.value == 0:
            raise HmxError("NVRTC produced no cubin")
        image = C.create_string_buffer(size.value)
        rt.nvrtcGetCUBIN(prg, image)
        return image.raw
    finally:
        rt.nvrtcDestroyProgram(C.byref(prg))


def compile_kernel_nvrtc(prog: CoefficientProgram, n, threads=None, min_blocks=None, variant=None, collapse=False):
    """cubin image (bytes) of the cell kernel, compiled with NVRTC for sm_100a; same source and macros as
    ``compile_kernel``."""
    threads, min_blocks, variant, coll = resolve(prog, n, threads, min_blocks, variant, collapse)
    with open(os.path.join(CSRC, "hmx_cell_entry.cu")) as f:
        src = f.read()
    headers, names = [], []
    for h in _HEADERS[:-1]:
        with open(os.path.join(CSRC, h)) as f:
            headers.append(f.read().encode())
        names.append(h.encode())
    headers.append(prog.source.encode())
    names.append(b"hmx_coeff_program.cuh")
    defs = kernel_defines(prog, n, threads, "hmx_coeff_program.cuh", min_blocks, variant, coll)
    return _nvrtc_compile(src, "hmx_cell_entry.cu", headers, names, defs)


def kernel_image(prog: CoefficientProgram, n, threads=None, min_blocks=None, variant=None, collapse=False):
    """cubin image of the cell kernel: from the in-tree cache, else built with nvcc, else (no nvcc on the host, or
    HMX_COMPILER=nvrtc) with NVRTC; the result is cached either way."""
    t, mb, v, coll = resolve(prog, n, threads, min_blocks, variant, collapse)
    cubin = os.path.join(KCACHE, kernel_key(prog, n, t, mb, v, coll) + ".cubin")
    want = os.environ.get("HMX_COMPILER", "")
    if not os.path.exists(cubin) or want == "nvrtc":
        have_nvcc = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
        if want == "nvrtc" or not have_nvcc:
            image = compile_kernel_nvrtc(prog, n, threads, min_blocks, variant, collapse)
            os.makedirs(KCACHE, exist_ok=True)
            if want != "nvrtc":  # an explicit nvrtc request is for testing: do not overwrite the nvcc-built cache
                with open(cubin + f".tmp{os.getpid()}", "wb") as f:
                    f.write(image)
                os.replace(cubin + f".tmp{os.getpid()}", cubin)
            return image
        compile_kernel(prog, n, threads, min_blocks=min_blocks, variant=variant, collapse=collapse)
    with open(cubin, "rb") as f:
        return f.read()
