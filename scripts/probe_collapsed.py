import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
case = K.BY_NAME["e3_fibre_rot_n8_c4"]; prog = K.program(case); qp,qw = K.tables(case, prog)
npts = 148*100
rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3)); x[:,1]*=0.4; x[:,2]*=0.1
xd = torch.tensor(x, device='cuda')
for mb in (1, 4, 5, 6, 8):
    try:
        s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, collapse=True, min_blocks=mb)
    except Exception as e:
        print(mb, 'failed', str(e)[:100]); continue
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    A = torch.empty((npts, 6, 6), device='cuda', dtype=torch.float64); it = torch.empty(npts, device='cuda', dtype=torch.int32)
    best = 1e9
    for rep in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A, it); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(f"min_blocks {mb}: ctas/sm {s.info['ctas_per_sm']} threads {s.info['threads']} smem {s.info['smem_bytes']}: {best:.2f} ms {npts/best*1e3:.0f} cells/s", flush=True)
    s.close()
