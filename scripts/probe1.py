import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
print(torch.cuda.get_device_name(0))
print("peaks fp64 TF, copy GB/s:", native.measure_peaks(0))
for name, npts in [("e3_fibre_rot_n8_c4", 4320), ("e3_fibre_rot_n8_c4", 296), ("p3_smooth_n8_c3", 196608), ("p2_laminate_wavy_n32_c2", 131072), ("p2_smooth_n16_c1", 2048), ("p2_inclusion_n16", 20000)]:
    case = K.BY_NAME[name]; prog = K.program(case); qp,qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=float(sys.argv[1]) if len(sys.argv)>1 else 1e-8)
    rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3)); 
    if case.dim==2: x[:,2]=0
    xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, s.m, s.m), device='cuda', dtype=torch.float64)
    it = torch.empty(npts, device='cuda', dtype=torch.int32); res = torch.empty(npts, device='cuda', dtype=torch.float64)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    for rep in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); s.cell_tensors_dev(npts, xd, A, it, res); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)
    t=time.time(); Ah = s.cell_tensors(x); th=time.time()-t
    print(name, s.info, f"dev {ms:.3f} ms -> {npts/ms*1e3:.3e} pts/s ; host-call {th*1e3:.1f} ms; mean it {it.float().mean().item():.1f} max res {res.max().item():.2e}", np.abs(Ah-A.cpu().numpy()).max())
