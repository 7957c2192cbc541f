import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
combos = {"p3_smooth_n8_c3": [(256,1),(256,2),(256,3),(128,2),(128,3),(128,4),(512,1)], "p2_laminate_wavy_n32_c2": [(512,1),(512,2),(256,2),(256,3),(1024,1)],
          "p2_inclusion_n16": [(128,4),(128,2),(64,4),(64,8),(256,2)]}
for name, lst in combos.items():
    case = K.BY_NAME[name]; prog = K.program(case); qp,qw = K.tables(case, prog)
    npts = 196608 if 'c3' in name else (131072 if 'c2' in name else 40000)
    rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3))
    if case.dim==2: x[:,2]=0
    xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, case.dim, case.dim), device='cuda', dtype=torch.float64)
    for nt, mb in lst:
        s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, threads=nt, min_blocks=mb)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        best=1e9
        for rep in range(4):
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record(); s.cell_tensors_dev(npts, xd, A); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
        print(f"{name} threads {nt} minb {mb} ctas/sm {s.info['ctas_per_sm']} smem {s.info['smem_bytes']}: {best:.3f} ms  {npts/best*1e3:.3e} pts/s", flush=True)
        s.close()
