"""CPU oracle for the hommx hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference algorithm
(/root/reference/src/hommx/hmm.py:298-432, cell_problem.py:16-388).  It is the
checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it; nothing under ``hommx_b200/`` does.

Pinning status: the reference cannot be imported in this image (DOLFINx, UFL,
basix, FFCx, dolfinx_mpc, PETSc are absent and un-vendored), so the oracle is
pinned on the known answers the reference's own tests encode
(test/integration/test_integration_poisson.py:121-143, 188-240, 398-473;
test_integration_linear_elasticity.py:205-322; test/unit/test_unit.py:25-103)
and on closed forms that follow from the reference's equations -- see
tests/test_oracle_*.py.  For smooth coefficients with quadrature degree >= 2
in 3-D the basix Xiao-Gimbutas tetrahedron table is not reproducible offline:
that boundary is "parity unpinned" (quadrature tables are therefore an input).
"""
