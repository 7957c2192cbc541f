"""A parity case at other thread counts (python scripts/probe_threads.py <case>[:n] <npts> 192 384:1 192:2 ... (threads[:min_blocks])."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
import dataclasses
name, _, nn = sys.argv[1].partition(":")
case = K.BY_NAME[name]
if nn: case = dataclasses.replace(case, n=int(nn))
prog = K.program(case); qp, qw = K.tables(case, prog)
npts = int(sys.argv[2])
rng = np.random.default_rng(0); x = rng.uniform(0, 1, (npts, 3))
if case.dim == 2: x[:, 2] = 0.0
xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device='cuda', dtype=torch.float64)
ref = None
for arg in sys.argv[3:] or ['384', '192']:
    nt, _, mb = arg.partition(':')
    nt, mb = int(nt), int(mb or 1)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, threads=nt, min_blocks=mb)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    if ref is None: ref = A.clone()
    print(f"threads {nt} minb {mb} ctas/sm {s.info.get('ctas_per_sm')} regs {s.info.get('regs')} smem {s.info['smem_bytes']}: {best:.2f} ms {npts/best*1e3:.0f} cells/s  max rel diff {float((A-ref).abs().max()/ref.abs().max()):.2e}", flush=True)
    s.close()
