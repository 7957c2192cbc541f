"""C4 cell kernel under experimental compile flags: python scripts/probe_c4_flags.py "" "-DHMX_BLK_NOPF" ...
(each argument is one HMX_EXTRA_NVCC setting; kernels are compiled on the fly if not in the cache)."""
import os, sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
case = K.BY_NAME["e3_fibre_rot_n8_c4"]; prog = K.program(case); qp, qw = K.tables(case, prog)
npts = 148 * 16
rng = np.random.default_rng(0); x = rng.uniform(0, 1, (npts, 3))
xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, 6, 6), device='cuda', dtype=torch.float64)
ref = None
for flags in sys.argv[1:] or [""]:
    os.environ["HMX_EXTRA_NVCC"] = flags
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    if ref is None: ref = A.clone()
    print(f"flags '{flags}': {best:.2f} ms {npts/best*1e3:.0f} cells/s  max rel diff {float((A-ref).abs().max()/ref.abs().max()):.2e}", flush=True)
    s.close()
