import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'scripts')
import numpy as np, torch
import cases as K
from hommx_b200 import native, micro, quadrature
import sweep_c5
jobs = [("p2_inclusion_n16", 16, 200000), ("p2_inclusion_n16", 32, 50000), ("p2_laminate_wavy_n32_c2", 32, 131072), ("p3_smooth_n8_c3", 8, 196608), ("p3_fulltensor_n6", 8, 20000)]
for name, n, npts in jobs:
    case = K.BY_NAME[name]; prog = K.program(case)
    st = micro.default_structure(case.dim, n); qp, qw = micro.quadrature_table(st, *quadrature.default_rule(case.dim, prog.degree))
    rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3))
    if case.dim==2: x[:,2]=0
    xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, case.dim, case.dim), device='cuda', dtype=torch.float64); it = torch.empty(npts, device='cuda', dtype=torch.int32)
    s = native.CellSolver(prog, n, qp, qw, rtol=1e-8)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    best=1e9
    for rep in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A, it); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(f"{name} n={n} threads {s.info['threads']} x {s.info['ctas_per_sm']}: {best:.3f} ms  {npts/best*1e3:.3e} pts/s mean its {it.float().mean().item():.1f}", flush=True)
    s.close()
