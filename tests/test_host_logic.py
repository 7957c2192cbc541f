"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol of
include/hmx.h, the macro CSR pattern / slot map / gather lists, cell sharding, the FEM stand-ins
and the constructor checks of the drop-in classes."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import coefficients as Cf
from hommx_b200 import LinearElasticityStratifiedHMM, PoissonHMM, PoissonStratifiedHMM, assembly, fem, mesh, micro, native
from hommx_b200 import ufl as pufl
from oracle import hmm_oracle as ho
from oracle import meshes as omesh
from oracle import npufl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hmx.h")).read()
    declared = set(re.findall(r"\b(hmx_[a-z_]+)\s*\(", header))
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    lib = native.load_library()
    for name in declared:
        assert hasattr(lib, name)
    # argument errors are reported without a GPU, and without a device creation fails loudly
    assert lib.hmx_create(None, None) == -1
    assert b"null" in lib.hmx_last_error(None)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    m = mesh.create_unit_square(2, 2)
    s = PoissonHMM(m, Cf.smooth_sin(pufl), lambda x: 1.0, mesh.create_unit_square(4, 4), 0.1)
    with pytest.raises(native.HmxError):
        s.solve()


@pytest.mark.parametrize("bs,dim", [(1, 2), (1, 3), (2, 2), (3, 3)])
def test_pattern_and_gather_reproduce_coo_assembly(bs, dim):
    m = mesh.create_rectangle((0, 0), (2, 1), (5, 3)) if dim == 2 else mesh.create_box((0, 0, 0), (1, 0.4, 0.1), (3, 2, 2))
    pat = assembly.build_pattern(m.cells, m.num_nodes, bs)
    nb = (dim + 1) * bs
    S = np.random.default_rng(0).normal(size=(m.num_cells, nb, nb))
    dofs = assembly.unroll_dofs(m.cells, bs)
    ref = sp.coo_matrix(
        (S.ravel(), (np.repeat(dofs, nb, axis=1).ravel(), np.tile(dofs, (1, nb)).ravel())), shape=(pat.n_dofs,) * 2
    ).tocsr()
    ref.sort_indices()
    assert np.array_equal(ref.indptr, pat.indptr) and np.array_equal(ref.indices, pat.indices)
    g = assembly.build_gather(pat.slot_map, pat.nnz)
    vals = np.array([S.ravel()[g.src[g.ptr[s] : g.ptr[s + 1]]].sum() for s in range(pat.nnz)])
    assert np.allclose(vals, ref.data, rtol=1e-13, atol=1e-13)
    # sources of a slot are in increasing (cell, i, j) order: the sum order is fixed
    for s in range(0, pat.nnz, 7):
        seg = g.src[g.ptr[s] : g.ptr[s + 1]]
        assert np.all(np.diff(seg) > 0)


def test_sharding_covers_every_cell_once_and_shared_slots_are_exact():
    m = mesh.create_unit_cube(3, 3, 2)
    pat = assembly.build_pattern(m.cells, m.num_nodes, 1)
    for world in (1, 2, 3, 4, 8):
        ranges = [assembly.shard_range(m.num_cells, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == m.num_cells
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        sh = assembly.shared_slots(pat.slot_map, m.num_cells, world, pat.nnz)
        S = np.random.default_rng(1).normal(size=(m.num_cells, 16))
        full = np.zeros(pat.nnz)
        np.add.at(full, pat.slot_map.ravel(), S.ravel())
        parts = []
        for lo, hi in ranges:
            v = np.zeros(pat.nnz)
            np.add.at(v, pat.slot_map[lo:hi].ravel(), S[lo:hi].ravel())
            parts.append(v)
        owners = sum((np.abs(p) > 0).astype(int) for p in parts)
        assert set(np.nonzero(owners > 1)[0]) <= set(sh)
        halo = sum(p[sh] for p in parts)
        for p in parts:
            q = p.copy()
            q[sh] = halo
            touched = np.abs(p) > 0
            assert np.allclose(q[touched], full[touched])


def test_load_vector_matches_oracle():
    for dim in (2, 3):
        m = mesh.create_unit_square(4, 3) if dim == 2 else mesh.create_unit_cube(2, 3, 2)
        om = omesh.create_rectangle([0, 0], [1, 1], [4, 3]) if dim == 2 else omesh.create_box([0, 0, 0], [1, 1, 1], [2, 3, 2])
        f = lambda u: (lambda x: 1.0 + u.sin(u.pi * x[0]) * x[1])  # noqa: E731
        b = fem.assemble_load(fem.FunctionSpace(m, 1), f(pufl))
        bo = ho.assemble_rhs(om, f(npufl), 1, degree=5)  # UFL estimate: sin -> 1+2, times x1 -> 4, times the test function -> 5
        assert np.allclose(b, bo, rtol=1e-12, atol=1e-14)


def test_constructor_checks_and_default_boundary_condition():
    m2, m3 = mesh.create_unit_square(3, 3), mesh.create_unit_cube(2, 2, 2)
    A = Cf.smooth_sin(pufl)
    with pytest.raises(ValueError, match="same dimensionality"):  # hmm.py:114-115
        PoissonHMM(m2, A, lambda x: 1.0, m3, 0.1)
    with pytest.raises(ValueError):  # unstructured / non-unit micro meshes are rejected, no fallback
        PoissonHMM(m2, A, lambda x: 1.0, mesh.create_rectangle((0, 0), (2, 1), (4, 4)), 0.1)
    with pytest.raises(ValueError, match="2x2"):  # the examples' 2x1 Jacobian
        PoissonStratifiedHMM(m2, A, lambda x: 1.0, mesh.create_unit_square(4, 4), 0.1, lambda x: pufl.as_matrix([[1.0], [0.0]]))
    s = PoissonHMM(m2, A, lambda x: 1.0, mesh.create_unit_square(4, 4), 0.1)
    assert sorted(s._bcs[0].dofs) == sorted(ho.boundary_nodes(omesh.create_unit_square(3, 3)))
    assert s.function_space.num_dofs == 16
    s2 = PoissonStratifiedHMM(m2, A, lambda x: 1.0, mesh.create_unit_square(4, 4), 0.1, Cf.dtheta_wavy(pufl))
    assert s2._bcs == []  # hmm.py:670-757: no default condition


def test_kernel_choice_for_elasticity_cells():
    """Host side of K5 and of the block sweep: which kernel variant / launch shape a cell gets."""
    import cases as K

    c4 = K.program(K.BY_NAME["e3_fibre_rot_n8_c4"])
    assert native.collapse_mask(c4, True) == 1  # the fibre runs along y0
    assert not native.dense_fits(c4, 8, 0) and native.dense_fits(c4, 8, 1)  # 1,536 vs 192 unknowns
    assert native.resolve(c4, 8, variant=native.DENSE, collapse=True) == (256, 1, native.DENSE, 1)
    with pytest.raises(native.HmxError, match="192"):
        native.resolve(c4, 8, variant=native.DENSE)  # loud, no fallback
    assert native.resolve(c4, 8)[:3] == (384, 1, native.MATRIX_FREE)  # PCG stays the default variant
    # full 3-D cells: threads chosen so that every warp owns whole planes of the last axis (block sweep)
    assert native.default_threads(3, 1, 8) == 384 and native.default_threads(3, 1, 10) == 192 and native.default_threads(3, 1, 12) == 384
    assert native.resolve(c4, 6)[:2] == (192, 2) and native.resolve(c4, 10, variant=native.MATRIX_FREE)[:2] == (192, 1)
    # cells that exceed one SM go to the cluster-resident stencil where a portable cluster holds them (10^3: 5 CTAs)
    assert native.resolve(c4, 10)[:3] == (416, 1, native.CLUSTER) and native.resolve(c4, 12)[2] == native.MATRIX_FREE
    poisson = K.program(K.BY_NAME["p2_smooth_n8"])
    assert not native.dense_fits(poisson, 8)
    m3 = mesh.create_unit_cube(2, 2, 2)
    with pytest.raises(ValueError, match="cell_solver"):
        LinearElasticityStratifiedHMM(m3, Cf.hooke_fibre_3d(pufl), lambda x: pufl.as_vector([0.0, 0.0, 1.0]), mesh.create_unit_cube(4, 4, 4),
                                      0.1, Cf.dtheta_rotation_3d(pufl), cell_solver="lu")  # fmt: skip


def test_micro_structure_detection():
    st = micro.detect_structure(mesh.create_unit_cube(5, 5, 5))
    assert st.n == 5 and st.vertex_order.shape == (6, 4, 3)
    # permuting the cells of the mesh does not matter, breaking the split does
    m = mesh.create_unit_square(4, 4)
    perm = np.random.default_rng(0).permutation(m.num_cells)
    st2 = micro.detect_structure(mesh.SimplexMesh(m.x, m.cells[perm], 2))
    assert st2.n == 4
    bad = m.cells.copy()
    bad[0] = [0, 1, 5]  # left-diagonal triangle
    with pytest.raises(ValueError):
        micro.detect_structure(mesh.SimplexMesh(m.x, bad, 2))


def test_nvrtc_compiles_the_cell_kernel_without_nvcc():
    """The in-process NVRTC path produces an sm_100a cubin from the same translation unit (no GPU needed)."""
    prog = __import__("cases").program(__import__("cases").BY_NAME["p2_smooth_n8"])
    image = native.compile_kernel_nvrtc(prog, 8)
    assert image[:4] == b"\x7fELF" and len(image) > 10000


def test_ctypes_structures_match_the_c_header(tmp_path):
    """include/hmx.h is the boundary: the ctypes mirrors in native.py (hmx_desc, hmx_micro_mesh) must have the layout a C
    compiler gives the header's structs, and the emulator's CellParams / MicroMesh the layout of csrc/hmx_cell_common.cuh."""
    import ctypes as C
    import subprocess
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cpu_emu"))
    import emu

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.cpp"
    src.write_text(
        '#include <cstddef>\n#include <cstdio>\n#define HMX_EMULATE 1\n#include "hmx.h"\n#include "hmx_cell_common.cuh"\n'
        "int main() {\n"
        '  printf("%zu %zu %zu %zu\\n", sizeof(hmx_desc), offsetof(hmx_desc, kernel_image), offsetof(hmx_desc, rtol), offsetof(hmx_desc, micro_mesh));\n'
        '  printf("%zu %zu %zu\\n", sizeof(hmx_micro_mesh), offsetof(hmx_micro_mesh, elem_nodes), offsetof(hmx_micro_mesh, diag));\n'
        '  printf("%zu %zu %zu\\n", sizeof(hmx::CellParams), offsetof(hmx::CellParams, nq), offsetof(hmx::CellParams, mesh));\n'
        '  printf("%zu %zu %zu\\n", sizeof(hmx::MicroMesh), offsetof(hmx::MicroMesh, elem_nodes), offsetof(hmx::MicroMesh, diag));\n'
        '  printf("%d\\n", HMX_ABI_VERSION);\n}\n'
    )
    exe = tmp_path / "layout"
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(root, "include"), "-I", native.CSRC, "-I", os.path.join(root, "tests", "cpu_emu"),
                        str(src), "-o", str(exe)], capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr[-2000:]
    got = [[int(v) for v in ln.split()] for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()]
    d, m = native.hmx_desc, native.hmx_micro_mesh
    assert got[0] == [C.sizeof(d), d.kernel_image.offset, d.rtol.offset, d.micro_mesh.offset]
    assert got[1] == [C.sizeof(m), m.elem_nodes.offset, m.diag.offset]
    assert got[2] == [C.sizeof(emu.CellParams), emu.CellParams.nq.offset, emu.CellParams.mesh.offset]
    assert got[3] == [C.sizeof(emu.MicroMesh), emu.MicroMesh.elem_nodes.offset, emu.MicroMesh.diag.offset]
    assert got[4] == [native.ABI_VERSION]
