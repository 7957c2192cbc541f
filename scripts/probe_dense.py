"""Direct (dense Cholesky, native.DENSE) against PCG (native.MATRIX_FREE) on small elasticity cells:
python scripts/probe_dense.py [case[:c] ...]   (":c" = with the exact axis collapse)."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
names = sys.argv[1:] or ["e3_fibre_rot_n8_c4:c", "e3_fibre_rot_n4", "e3_hooke_smooth_n4", "e3_cubic_shear_n4", "e3_fibre_rot_n4:c",
                         "e2_hooke_sin_n6", "e2_hooke_sin_strat_n7", "e3_hooke_const_n3"]
for arg in names:
    name, _, c = arg.partition(":")
    case = K.BY_NAME[name]; prog = K.program(case); qp, qw = K.tables(case, prog)
    npts = 148 * 64
    x = K.points(case, npts, seed=3)
    xd = torch.tensor(x, device='cuda'); out = {}
    for label, var in (("pcg", native.MATRIX_FREE), ("dense", native.DENSE)):
        s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, variant=var, collapse=bool(c))
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device='cuda', dtype=torch.float64)
        it = torch.zeros(npts, device='cuda', dtype=torch.int32)
        best = 1e9
        for rep in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); s.cell_tensors_dev(npts, xd, A, it); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        out[label] = (A.clone(), best, float(it.float().mean()))
        s.close()
    d = float((out["pcg"][0] - out["dense"][0]).abs().max() / out["pcg"][0].abs().max())
    print(f"{arg}: pcg {npts/out['pcg'][1]*1e3:.0f} cells/s ({out['pcg'][2]:.0f} its)  dense {npts/out['dense'][1]*1e3:.0f} cells/s  rel diff {d:.1e}", flush=True)
