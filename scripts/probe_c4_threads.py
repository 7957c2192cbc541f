"""C4 cell kernel at other thread counts per right-hand side (python scripts/probe_c4_threads.py 192 384 ...)."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
case = K.BY_NAME["e3_fibre_rot_n8_c4"]; prog = K.program(case); qp, qw = K.tables(case, prog)
npts = 148 * 16
rng = np.random.default_rng(0); x = rng.uniform(0, 1, (npts, 3))
xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, 6, 6), device='cuda', dtype=torch.float64)
ref = None
for nt in [int(a) for a in sys.argv[1:]] or [384, 192]:
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, threads=nt, min_blocks=1)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    if ref is None: ref = A.clone()
    print(f"threads {nt} regs {s.info.get('regs')} smem {s.info['smem_bytes']}: {best:.2f} ms {npts/best*1e3:.0f} cells/s  max rel diff {float((A-ref).abs().max()/ref.abs().max()):.2e}", flush=True)
    s.close()
