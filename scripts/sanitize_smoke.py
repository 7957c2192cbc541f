"""Tiny run of every kernel family for compute-sanitizer (memcheck / synccheck / racecheck), one tool per call:
   compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import cases as K
import coefficients as Cf
import general_meshes as G
from hommx_b200 import codegen, micro, native, quadrature
from hommx_b200 import ufl as pufl

RUNS = [("p2_inclusion_n16", {}), ("p3_smooth_n4", {}), ("e3_fibre_rot_n4", {}), ("e3_fibre_rot_n4", {"collapse": True}),
        ("e2_hooke_sin_n6", {}), ("e3_fibre_rot_n8_c4", {}), ("e3_fibre_rot_n4", {"variant": native.DENSE, "collapse": True}),
        ("e3_fibre_rot_n4", {"variant": native.CLUSTER}), ("e3_cubic_shear_n4", {"variant": native.CLUSTER}),
        ("e3_fibre_rot_n8_c4", {"variant": native.CLUSTER})]  # fmt: skip
for name, kw in RUNS:
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-6, **kw)
    s.set_grid(2 * max(1, s.info.get("cluster", 1)))
    A, it, res = s.cell_tensors(K.points(case, 3), return_stats=True)
    print(name, kw, it.tolist(), float(np.abs(A).max()), flush=True)
    s.close()
# element-list kernel on a general periodic micro mesh
prog = codegen.build_program(Cf.hooke_sin_2d(pufl), 2, 1, Cf.dtheta_test_stratified(pufl))
tables = micro.ElementListTables(G.perturbed(2, 6, 3), *quadrature.default_rule(2, prog.degree))
s = native.CellSolver(prog, 0, None, None, rtol=1e-8, micro_tables=tables)
s.set_grid(2)
A, it, res = s.cell_tensors(np.array([[0.3, 0.2], [0.7, 0.6], [0.1, 0.9]]), return_stats=True)
print("element list", it.tolist(), float(np.abs(A).max()), flush=True)
s.close()
print("done")
