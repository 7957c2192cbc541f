"""BASELINE configs[4]: synthetic micro-solve throughput sweep -- random macro points, 2-D 16^2..64^2 and
3-D 6^3..12^3 micro cells, three coefficient families, two PCG tolerances.  Writes a markdown table.

    python scripts/sweep_c5.py [--out profiles/r01_sweep_c5.md] [--max-points 1000000]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import coefficients as Cf  # noqa: E402
from hommx_b200 import codegen, micro, native, quadrature  # noqa: E402
from hommx_b200 import ufl as pufl  # noqa: E402

# (label, dim, kind, coefficient, Dtheta, micro sizes)
FAMILIES = [
    ("poisson2d smooth 1.1+x0+sin(2 pi y0)", 2, 0, "smooth_sin", None, (16, 24, 32, 48, 64)),
    ("poisson2d inclusion (inclusion.py:107-118), stratified", 2, 0, "inclusion", "dtheta_inclusion", (16, 24, 32, 48, 64)),
    ("poisson3d smooth 1.1+x0+sin(2 pi y0)", 3, 0, "smooth_sin", None, (6, 8, 10, 12)),
    ("poisson3d full tensor (all y), sheared", 3, 0, "full_tensor_3d", "dtheta_shear_3d", (6, 8, 10, 12)),
    ("elasticity3d fibre mu 100/0.001, rotated (C4)", 3, 1, "hooke_fibre_3d", "dtheta_rotation_3d", (6, 8, 10, 12)),
]


def program(f):
    A = getattr(Cf, f[3])(pufl)
    Dt = getattr(Cf, f[4])(pufl) if f[4] else None
    return codegen.build_program(A, f[1], f[2], Dt)


def fits(prog, n):
    """Every size runs: what does not fit in shared memory moves to the L2 scratch (native.vectors_in_l2)."""
    return True


def kernel_jobs():
    jobs = []
    for f in FAMILIES:
        prog = program(f)
        for n in f[5]:
            if fits(prog, n):
                jobs.append((prog, n, None))
    return jobs


def flops_per_rhs_iteration(prog, n):
    d = prog.dim
    bs = 1 if prog.kind == 0 else d
    nodes = n**d
    nnz = (7 if d == 2 else 15) * nodes * bs * bs
    return 2 * nnz + 11 * nodes * bs, nnz


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep_c5.md"))
    ap.add_argument("--max-points", type=int, default=1000000)
    ap.add_argument("--budget-ms", type=float, default=1500.0, help="skip larger point counts once a launch exceeds this")
    ap.add_argument("--only", default="", help="run only the families whose label contains this text")
    args = ap.parse_args()
    peak, _ = native.measure_peaks(0)
    lines = [
        "# C5 sweep: micro cell solves/s on one B200 (inputs resident, CUDA events, best of 3)",
        "",
        f"FP64 DFMA peak measured on this GPU: {peak:.1f} TFLOP/s.  `frac` = algorithmic FLOPs (SURVEY 8d) / time / peak.",
        "PCG atol 1e-10.  Points x ~ U([0,1]^d), numpy default_rng(0).",
        "",
        "Kernels: `pcg` = the matrix-free / stencil PCG kernel (one CTA per point; vectors in L2 beyond 8^3 elasticity), `cluster` = the",
        "assembled stencil resident in a thread-block cluster's distributed shared memory (3-D elasticity, csrc/hmx_cell_cluster.cuh).",
        "",
        "| family | micro | kernel | threads x CTAs/SM | rtol | points | ms | cell solves/s | mean its | TFLOP/s | frac |",
        "|---|---|---|---|---|---|---|---|---|---|---|",
    ]
    rng = np.random.default_rng(0)
    for f in FAMILIES:
        if args.only not in f[0]:
            continue
        prog = program(f)
        for n in f[5]:
            if not fits(prog, n):
                lines.append(f"| {f[0]} | {n}^{f[1]} | - | - | - | - | does not fit one CTA (needs the right-hand sides split over a cluster) | | | |")
                continue
            st = micro.default_structure(f[1], n)
            qp, qw = micro.quadrature_table(st, *quadrature.default_rule(f[1], prog.degree))
            kernels = [("pcg", native.MATRIX_FREE)]
            if f[2] == 1 and f[1] == 3 and n % 2 == 0 and native.cluster_size(prog, n) >= 2:
                kernels.append(("cluster", native.CLUSTER))
            per_it, nnz = flops_per_rhs_iteration(prog, n)
            for kname, variant in kernels:
              sol = native.CellSolver(prog, n, qp, qw, variant=variant)
              sol.set_stream(torch.cuda.current_stream().cuda_stream)
              for rtol in (1e-6, 1e-10):
                  sol.set_tolerances(rtol, 1e-10)
                  for npts in (10**4, 10**5, 10**6, 10**7):
                      if npts > args.max_points:
                          break
                      x = rng.uniform(0, 1, (npts, 3))
                      if f[1] == 2:
                          x[:, 2] = 0.0
                      xd = torch.as_tensor(x, device="cuda")
                      A = torch.empty((npts, sol.m, sol.m), dtype=torch.float64, device="cuda")
                      it = torch.empty(npts, dtype=torch.int32, device="cuda")
                      best = 1e30
                      for rep in range(3):
                          sol.rhs_iterations(reset=True)
                          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                          e0.record()
                          sol.cell_tensors_dev(npts, xd, A, it, None)
                          e1.record()
                          torch.cuda.synchronize()
                          best = min(best, e0.elapsed_time(e1))
                          if best > args.budget_ms:
                              break
                      rhs_its = sol.rhs_iterations(reset=True)
                      fl = rhs_its * per_it + npts * sol.m * sol.m * 2 * nnz
                      tf = fl / (best * 1e-3) / 1e12
                      lines.append(
                          f"| {f[0]} | {n}^{f[1]} | {kname}{' x' + str(sol.info['cluster']) if sol.info['cluster'] > 1 else ''} | {sol.info['threads']} x {sol.info['ctas_per_sm']} | {rtol:.0e} | {npts:.0e} | {best:.2f} | "
                          f"{npts / best * 1e3:.3e} | {it.float().mean().item():.1f} | {tf:.2f} | {tf / peak:.3f} |"
                      )
                      print(lines[-1], flush=True)
                      if best > args.budget_ms:
                          break
              sol.close()
    with open(args.out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
