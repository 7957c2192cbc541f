"""The unmodified CUDA kernel source, compiled by g++ against tests/cpu_emu (fibers for CUDA
threads), compared with the oracle.  Checks the kernel LOGIC without a GPU; the same cases run
on the real device in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cpu_emu"))
import cases as K  # noqa: E402
import emu  # noqa: E402
from hommx_b200 import native  # noqa: E402
from oracle import hmm_oracle as ho  # noqa: E402

LIGHT = [c for c in K.CASES if not c.heavy]


@pytest.mark.parametrize("case", LIGHT, ids=[c.name for c in LIGHT])
def test_cell_tensor_matches_oracle(case):
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, threads=case.threads)
    x = K.points(case, 2)
    Ah, it, res = s.cell_tensors(x, return_stats=True)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max(), (case.name, it, res)


DENSE = [(c, co) for c in K.CASES if c.kind == 1 and c.threads is None for co in (False, True)
         if native.dense_fits(K.program(c), c.n, native.collapse_mask(K.program(c), co))
         and (not co or native.collapse_mask(K.program(c), True)) and (not c.heavy or co)]


@pytest.mark.parametrize("case,collapse", DENSE, ids=[c.name + ("_collapsed" if co else "") for c, co in DENSE])
def test_dense_cholesky_variant_matches_oracle(case, collapse):
    """K5: the direct solve of small elasticity cells (csrc/hmx_cell_dense.cuh) -- odd and even n, 2-D and 3-D,
    with and without collapsed axes, up to the 192 unknowns of the collapsed BASELINE config 4 cell."""
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, variant=native.DENSE, collapse=collapse, grid=2)
    x = K.points(case, 3)
    Ah, it, res = s.cell_tensors(x, return_stats=True)
    assert np.all(it == 0) and np.all(res == 0.0)  # no iteration
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()


@pytest.mark.parametrize("name,collapse", [("e2_hooke_sin_strat_n7", False), ("e3_fibre_rot_n4", True), ("e3_hooke_smooth_n4", False)])
def test_dense_cholesky_correctors_and_local_matrices(name, collapse):
    """The rare paths of the direct kernel: back substitution for the correctors, fused local macro matrices."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    ref = emu.EmuSolver(prog, case.n, qp, qw, rtol=1e-12, collapse=collapse)
    s = emu.EmuSolver(prog, case.n, qp, qw, variant=native.DENSE, collapse=collapse)
    x = K.points(case, 1)
    a, b = s.correctors(x)[0], ref.correctors(x)[0]
    ax = tuple(range(2, a.ndim))
    a, b = a - a.mean(axis=ax, keepdims=True), b - b.mean(axis=ax, keepdims=True)
    assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max() + 1e-14
    cells, xyz = K.random_simplices(case.dim, 3)
    S, Sr = s.local_matrices(cells, xyz)[0], ref.local_matrices(cells, xyz)[0]
    assert np.abs(S - Sr).max() <= 1e-10 * np.abs(Sr).max()


@pytest.mark.parametrize("name", ["p2_fulltensor_strat_n9", "p3_fulltensor_shear_n5", "e2_hooke_sin_strat_n7", "e3_hooke_smooth_n4"])
def test_local_matrix_matches_oracle(name):
    """Fused mode: macro cell vertices in, S_loc out (hmm.py:334-369 end to end)."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, threads=case.threads)
    cells, xyz = K.random_simplices(case.dim, 2)
    S, Ah = s.local_matrices(cells, xyz)
    mic = K.oracle_cell(case, prog)
    for k in range(len(cells)):
        verts = xyz[cells[k]]
        Ao = K.oracle_tensor(case, mic, verts.mean(axis=0))
        So = ho.local_stiffness_from_tensor(Ao, verts, mic.kind)
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()
        assert np.abs(S[k] - So).max() <= case.tol * np.abs(So).max()


COLLAPSIBLE = ["p2_smooth_n15", "p2_laminate_wavy_n16", "p2_smooth_strat_n12", "p3_smooth_n8_c3", "p2_xonly_n7", "e2_hooke_sin_n6",
               "e2_hooke_sin_strat_n7", "e3_hooke_const_n3", "e3_fibre_rot_n4"]  # fmt: skip


@pytest.mark.parametrize("name", COLLAPSIBLE)
def test_axis_collapse_is_exact(name):
    """Coefficients that do not depend on y_a: the cell problem solved on one layer of cubes along a
    (CO::YDEP / HMX_COLL) reproduces the full-grid oracle."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    assert prog.ydep != (1 << case.dim) - 1
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, collapse=True)
    x = K.points(case, 2)
    Ah = s.cell_tensors(x)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()


@pytest.mark.parametrize("name,collapse", [("p2_fulltensor_strat_n9", False), ("p3_smooth_n4", True), ("e2_hooke_sin_strat_n7", False),
                                           ("e3_fibre_rot_n4", True), ("e3_hooke_smooth_n4", False)])  # fmt: skip
def test_correctors_match_oracle(name, collapse):
    """chi_q (hmm.py:1211-1213) up to the additive constants the periodic problem leaves free."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=1e-12, collapse=collapse)
    x = K.points(case, 1)
    chi = s.correctors(x)[0]  # (n_rhs, bs, [z,] y, x) possibly with collapsed axes of extent 1
    d, n = case.dim, case.n
    chi = np.broadcast_to(chi, chi.shape[:2] + (n,) * d)
    mic = K.oracle_cell(case, prog)
    import coefficients as Cf
    from oracle import npufl

    M = None
    if case.dtheta:
        M = np.asarray(getattr(Cf, case.dtheta)(npufl)(x[0]))[..., 0]
    _, chis = ho.cell_tensor(mic, getattr(Cf, case.coeff)(npufl), x[0], M, return_correctors=True)
    # oracle dofs: periodic node id (np.unique order of master vertices = natural order), comps interleaved
    bs = mic.bs
    scale = max(np.abs(c - c.mean()).max() for c in chis)
    for q in range(len(chis)):
        co = chis[q].reshape(-1, bs).T.reshape((bs,) + (n,) * d)  # natural order: slowest axis first
        for k in range(bs):
            a, b = chi[q, k] - chi[q, k].mean(), co[k] - co[k].mean()
            assert np.abs(a - b).max() <= 1e-8 * scale + 1e-14, (q, k)


@pytest.mark.parametrize("name", ["e3_hooke_smooth_n4", "e2_hooke_sin_strat_n7", "e3_fibre_rot_n4", "p3_fulltensor_n6", "p2_inclusion_n16"])
def test_vectors_in_l2_fallback_matches_oracle(name, monkeypatch):
    """Large elasticity cells keep p and K p in the L2 scratch instead of shared memory (HMX_VGLOB);
    forced here on small cells."""
    monkeypatch.setenv("HMX_FORCE_VGLOB", "1")
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, variant=0, threads=case.threads)
    assert "g1p" in os.path.basename(s.lib._name)
    x = K.points(case, 2)
    Ah = s.cell_tensors(x)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()


SCHED = [("p2_smooth_n8", {}), ("p3_smooth_n4", {}), ("p2_inclusion_n16", {}), ("e2_hooke_sin_n6", {}), ("e3_fibre_rot_n4", {}),
         ("e3_fibre_rot_n4", {"collapse": True}), ("e3_fibre_rot_n4_blocks", {}), ("e3_hooke_smooth_shear_n6", {}),
         ("e3_hooke_smooth_n4", {"variant": native.DENSE}), ("e2_hooke_sin_strat_n7", {"variant": native.DENSE})]  # fmt: skip


@pytest.mark.parametrize("name,kw", SCHED, ids=[n + "".join(f"_{k}{v}" for k, v in kw.items()) for n, kw in SCHED])
def test_results_do_not_depend_on_the_thread_schedule(name, kw, monkeypatch):
    """CPU-side race check: the emulator runs the fibers of a CTA between barriers in forward, reverse and random
    order (tests/cpu_emu/emu_runtime.h, HMX_EMU_ORDER).  Any read of a location another thread writes in the
    same barrier interval -- a data race on the device -- makes the result depend on that order; the kernels
    are deterministic, so the tensors, correctors and iteration counts must be bit-identical."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = emu.EmuSolver(prog, case.n, qp, qw, rtol=1e-9, threads=case.threads, grid=2, **kw)
    x = K.points(case, 3, seed=7)
    ref = None
    for order in ("forward", "reverse", "shuffle:1", "shuffle:2"):
        monkeypatch.setenv("HMX_EMU_ORDER", order)
        got = (s.cell_tensors(x, return_stats=True), s.correctors(x[:1]))
        if ref is None:
            ref = got
            continue
        assert np.array_equal(got[0][0], ref[0][0]) and np.array_equal(got[0][1], ref[0][1]), order
        assert np.array_equal(got[1], ref[1]), order


TWO_LEVEL = ["e3_fibre_rot_n4_blocks", "e3_hooke_smooth_shear_n6", "e2_hooke_sin_n6", "e3_fibre_rot_n8_c4"]


@pytest.mark.parametrize("name", TWO_LEVEL)
def test_two_level_preconditioner_matches_oracle(name, monkeypatch):
    """csrc/hmx_cell_coarse.cuh: the additive two-level PCG (Kuhn-nested level-1 space; 8^3 fibre cell: summed up
    along the fibre axis to 1 x 4 x 4 nodes) reaches the same tensors as block-Jacobi PCG and the oracle -- with and without the
    semi-coarsening step, for a cell whose sweep clears y itself (n = 6) and for 2-D -- and pays on the hard cell."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    x = K.points(case, 1 if case.heavy else 2)
    mic = K.oracle_cell(case, prog)
    its = {}
    for mode in ("jacobi", "twolevel"):
        monkeypatch.setenv("HMX_PRECOND", mode)
        s = emu.EmuSolver(prog, case.n, qp, qw, rtol=1e-9, threads=case.threads)
        assert ("p1_" if mode == "twolevel" else "p0_") in os.path.basename(s.lib._name)
        Ah, it, res = s.cell_tensors(x, return_stats=True)
        its[mode] = it
        for k in range(len(x)):
            Ao = K.oracle_tensor(case, mic, x[k])
            assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max(), (name, mode, it, res)
    if name == "e3_fibre_rot_n8_c4":  # BASELINE config 4: the coarse correction must cut the iteration count
        assert its["twolevel"].max() <= 0.7 * its["jacobi"].max(), its
