#!/usr/bin/env python
"""bench.py -- the hot path of hommx on B200: micro cell solves/s and macro assembly wall-time.

A "step" is one pass of the hot path over the macro cells of the workload: what the reference does
in ``BaseHMM._assemble_stiffness`` + ``self._A.assemble()`` (/root/reference/src/hommx/hmm.py:
298-332, 442): per macro cell the periodic micro problems, A_hom, the local matrix, the scatter
into the macro CSR values and -- when cells are sharded over GPUs -- the sum of shared slots.

Default workload: BASELINE.json configs[3], the one the north-star target is quoted on
(LinearElasticityStratifiedHMM, rotated-fibre beam [0,1]x[0,0.4]x[0,0.1], 8^3 micro cell, 6
right-hand sides per point).  ``--workload c1|c2|c3`` select the Poisson configs.  Per-GPU work
is fixed as N grows (weak scaling): the macro mesh gets N slabs of cells.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU restatement of the reference
algorithm (oracle/, kind "port": DOLFINx/PETSc are not installable in this image) on the host
cores for the same workload/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import coefficients as Cf  # noqa: E402  (tests/coefficients.py: the reference's example coefficients)

# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: class, dim, micro n, coefficient, Dtheta, per-GPU macro mesh (slab count multiplies the last axis)
    "c4": dict(cls="LinearElasticityStratifiedHMM", dim=3, kind=1, n=8, coeff="hooke_fibre_3d", dtheta="dtheta_rotation_3d",
               box=((0.0, 0.0, 0.0), (1.0, 0.4, 0.1)), cells=(40, 16, 4), eps=0.01,
               desc="BASELINE configs[3]: rotated-fibre beam, Hooke mu=100/0.001 lambda=1, 8^3 micro cell, 6 RHS/point"),
    "c3": dict(cls="PoissonHMM", dim=3, kind=0, n=8, coeff="smooth_sin", dtheta=None,
               box=((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), cells=(32, 32, 32), eps=2.0**-3,
               desc="BASELINE configs[2]: PoissonHMM 3D, 32^3 macro mesh per GPU, 8^3 micro cell"),
    "c2": dict(cls="PoissonStratifiedHMM", dim=2, kind=0, n=32, coeff="laminate", dtheta="dtheta_wavy",
               box=((0.0, 0.0), (1.0, 1.0)), cells=(256, 256), eps=1e-5,
               desc="BASELINE configs[1]: PoissonStratifiedHMM wavy laminate, 256x256 macro mesh per GPU, 32x32 micro cell"),
    "c1": dict(cls="PoissonHMM", dim=2, kind=0, n=16, coeff="smooth_sin", dtheta=None,
               box=((0.0, 0.0), (1.0, 1.0)), cells=(32, 32), eps=2.0**-5,
               desc="BASELINE configs[0]: PoissonHMM 2D, 32x32 macro mesh, 16x16 micro cell"),
}  # fmt: skip
CELL_RTOL, CELL_ATOL = 1e-8, 1e-10


SHRINK = 1


def build_solver(wl, world, collapse=False, **kw):
    import hommx_b200 as hx
    from hommx_b200 import mesh
    from hommx_b200 import ufl as pufl

    w = WORKLOADS[wl]
    cells = [max(1, c // SHRINK) for c in w["cells"]]
    cells[-1] *= world
    msh = mesh.create_rectangle(*w["box"], cells) if w["dim"] == 2 else mesh.create_box(*w["box"], cells)
    mic = mesh.create_unit_square(w["n"], w["n"]) if w["dim"] == 2 else mesh.create_unit_cube(w["n"], w["n"], w["n"])
    A = getattr(Cf, w["coeff"])(pufl)
    f = (lambda x: 1.0) if w["kind"] == 0 else (lambda x: pufl.as_vector([0.0] * (w["dim"] - 1) + [-0.05 * 0.4**2]))
    opts = {"ksp_rtol": CELL_RTOL, "ksp_atol": CELL_ATOL}
    cls = getattr(hx, w["cls"])
    if w["dtheta"]:
        return cls(msh, A, f, mic, w["eps"], getattr(Cf, w["dtheta"])(pufl), petsc_options_cell_problem=opts,
                   collapse_invariant_axes=collapse, **kw)
    return cls(msh, A, f, mic, w["eps"], petsc_options_cell_problem=opts, collapse_invariant_axes=collapse, **kw)


def kernel_jobs():
    """Cell kernels the bench needs (compiled by __graft_entry__.build())."""
    from hommx_b200 import codegen
    from hommx_b200 import ufl as pufl

    jobs = []
    for w in WORKLOADS.values():
        A = getattr(Cf, w["coeff"])(pufl)
        Dt = getattr(Cf, w["dtheta"])(pufl) if w["dtheta"] else None
        prog = codegen.build_program(A, w["dim"], w["kind"], Dt)
        jobs.append((prog, w["n"], None))
        jobs.append((prog, w["n"], None, False, True, None, None, True))  # axis-collapsed variant (other_workloads)
    return jobs


def algorithmic_flops(w, n_pts, rhs_iterations):
    """F = sum_rhs I (2 nnz_K + 11 n_dof) + n_pts n_rhs^2 2 nnz_K, nnz_K the structural
    non-zeros of the assembled periodic micro stiffness matrix (7n^2 / 15n^3 stencil, x bs^2)."""
    d, n = w["dim"], w["n"]
    bs = 1 if w["kind"] == 0 else d
    nrhs = d if w["kind"] == 0 else d * (d + 1) // 2
    nodes = n**d
    nnz = (7 if d == 2 else 15) * nodes * bs * bs
    ndof = nodes * bs
    return rhs_iterations * (2 * nnz + 11 * ndof) + n_pts * nrhs * nrhs * 2 * nnz


# ---------------------------------------------------------------------------------------------
# CPU arm: the restated reference algorithm (oracle), all host cores
# ---------------------------------------------------------------------------------------------
def _oracle_worker(args):
    wl, idx = args
    from oracle import hmm_oracle as ho
    from oracle import meshes as omesh
    from oracle import npufl

    w = WORKLOADS[wl]
    n = w["n"]
    state = _oracle_worker.__dict__.setdefault("state", {})
    if wl not in state:
        mm = omesh.create_unit_square(n, n) if w["dim"] == 2 else omesh.create_unit_cube(n, n, n)
        degree = {"smooth_sin": 3}.get(w["coeff"], 0)
        macro = omesh.create_rectangle(*w["box"], list(w["cells"])) if w["dim"] == 2 else omesh.create_box(*w["box"], list(w["cells"]))
        Dt = None
        if w["dtheta"]:
            Dn = getattr(Cf, w["dtheta"])(npufl)
            Dt = lambda x: np.asarray(Dn(np.asarray(x, float)))[..., 0]  # noqa: E731
        state[wl] = (ho.MicroCell(mm, "poisson" if w["kind"] == 0 else "elasticity", degree), macro, getattr(Cf, w["coeff"])(npufl), Dt)
    mic, macro, A, Dt = state[wl]
    out = []
    for c in idx:
        verts = macro.x[macro.cells[c]]
        out.append(ho.local_stiffness_literal(mic, A, verts, w["eps"], Dt))  # hmm.py:334-369, literally
    return len(out)


def cpu_reference_rate(wl, budget_s, procs=None):
    """macro cells/s of the restated reference algorithm on `procs` host cores over a bounded sample."""
    import multiprocessing as mp

    procs = procs or (os.cpu_count() or 1)
    w = WORKLOADS[wl]
    n_total = int(np.prod(w["cells"])) * (2 if w["dim"] == 2 else 6)
    rng = np.random.default_rng(0)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        # calibrate on one cell per process (also builds the per-process micro cell once)
        t0 = time.perf_counter()
        pool.map(_oracle_worker, [(wl, [int(c)]) for c in rng.integers(0, n_total, procs)])
        t1 = time.perf_counter() - t0
        per_proc = max(1, int(budget_s / max(t1, 1e-3)))
        chunks = [(wl, [int(c) for c in rng.integers(0, n_total, per_proc)]) for _ in range(procs)]
        t0 = time.perf_counter()
        done = sum(pool.map(_oracle_worker, chunks))
        dt = time.perf_counter() - t0
    return done / dt, procs, f"{done} macro cells of {wl} drawn at random, literal n_b-corrector algorithm (sparse LU), {dt:.1f} s"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}  # fmt: skip


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    budget = 20.0
    rates = []
    t_all = time.perf_counter()
    for s in range(args.warmup + args.steps):
        r, cores, sample = cpu_reference_rate(args.workload, budget / max(1, args.steps + args.warmup))
        if s >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    n_cells = int(np.prod(w["cells"])) * (2 if w["dim"] == 2 else 6) * world
    line = {
        "impl": "reference", "metric": "micro cell solves/sec (FP64)", "value": value, "unit": "cell solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_cells / value * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "macro_cells": n_cells,
                   "note": "ms_per_step extrapolated from the bounded sample to the whole workload"},
        "cpu_baseline": {"value": value, "unit": "cell solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from hommx_b200 import native

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL may print its version banner on fd 1 when the communicator comes up: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    w = WORKLOADS[args.workload]
    hmm = build_solver(args.workload, world, device=local_rank)
    hmm._ensure_solver()
    sol, d = hmm._solver, hmm._dev
    stream = torch.cuda.current_stream()
    sol.set_stream(stream.cuda_stream)
    n_local = d["hi"] - d["lo"]
    n_total = hmm._msh.num_cells
    nnz = hmm._pattern.nnz
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = {"n": 0}

    def step_resident(ev=None):
        """inputs already in HBM: cell kernel -> gather -> (halo sum)"""
        if ev:
            ev[0].record(stream)
        sol.assemble_macro_dev(n_local, d["cells"], hmm._msh.num_nodes, d["xyz"], 0, None, None, None, d["S"], d["it"], d["res"])
        if ev:
            ev[1].record(stream)
        sol.gather_csr_dev(nnz, d["ptr"], d["src"], d["S"], d["vals"])
        launches["n"] += 2
        if world > 1:
            hmm._halo_sum()
            launches["n"] += 2
        if ev:
            ev[2].record(stream)

    # ---- resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
        flush.zero_()
    barrier()
    sol.rhs_iterations(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches["n"] = 0
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_resident(evs[s])
        flush.zero_()  # L2 flush between timed steps (outside the event pairs)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [e[0].elapsed_time(e[2]) for e in evs]
    cell_ms = [e[0].elapsed_time(e[1]) for e in evs]
    rhs_its = sol.rhs_iterations(reset=True)
    tot_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = tot_ms.item() / args.steps
    value = n_total / (ms_per_step * 1e-3)
    n_launch = launches["n"]
    mean_it = float(d["it"].float().mean().item())
    max_res = float(d["res"].max().item())

    # ---- end to end: host buffers in, host buffers out ------------------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_cells, h_xyz = pin(hmm._msh.cells[d["lo"] : d["hi"]].astype(np.int32)), pin(hmm._msh.x)
    h_ptr, h_src = pin(d["ptr"].cpu().numpy()), pin(d["src"].cpu().numpy())
    h_vals = torch.empty(nnz, dtype=torch.float64).pin_memory()
    h_it = torch.empty(n_local, dtype=torch.int32).pin_memory()
    h_res = torch.empty(n_local, dtype=torch.float64).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_cells, h_xyz, h_ptr, h_src))
    d2h = sum(t.numel() * t.element_size() for t in (h_vals, h_it, h_res))

    def step_e2e():
        if world == 1:  # the host-buffer C-ABI entry point: copies in, kernels, copies out, sync
            sol._check(sol.lib.hmx_assemble_macro(sol._h, n_local, h_cells.data_ptr(), hmm._msh.num_nodes, h_xyz.data_ptr(), nnz, h_ptr.data_ptr(),
                                       h_src.data_ptr(), h_vals.data_ptr(), None, h_it.data_ptr(), h_res.data_ptr()))  # fmt: skip
        else:  # sharded: the halo sum sits between the kernels and the copy-out
            d["cells"].copy_(h_cells, non_blocking=True)
            d["xyz"].copy_(h_xyz, non_blocking=True)
            d["ptr"].copy_(h_ptr, non_blocking=True)
            d["src"].copy_(h_src, non_blocking=True)
            step_resident()
            h_vals.copy_(d["vals"], non_blocking=True)
            h_it.copy_(d["it"], non_blocking=True)
            h_res.copy_(d["res"], non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * args.steps / e2e_s.item()
    if world == 1:  # the two paths agree bit for bit
        assert np.array_equal(h_vals.numpy(), d["vals"].cpu().numpy())

    # ---- roofline of the dominant kernel (the cell kernel), this rank ---------------------------
    flops = algorithmic_flops(w, n_local * args.steps, rhs_its) / args.steps
    cell_avg_ms = float(np.mean(cell_ms))
    fp64_peak, copy_gbs = native.measure_peaks(local_rank)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)

    # secondary workloads: parity-test configs and the axis-collapsed form, reported for context.
    # At N > 1 only the collapsed headline config is repeated (every rank takes part).
    other = {}
    if not args.no_extra:
        # cell_solver "auto": small hard elasticity cells (the collapsed C4 cell) are factorised directly (K5)
        todo = ((("c2", False, "auto"), ("c3", False, "auto"), ("c4", True, "auto"), ("c4", True, "pcg"), ("c3", True, "auto"),
                 ("c2", True, "auto")) if world == 1 else ((args.workload, True, "auto"),))  # fmt: skip
        for name, collapse, how in todo:
            if name == args.workload and not collapse:
                continue
            key = name + ("_axis_collapsed" if collapse else "") + ("_pcg" if how == "pcg" else "")
            try:
                other[key] = quick_rate(name, local_rank, collapse, world, how)
            except Exception as e:  # never lose the headline line
                other[key] = {"error": str(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # CPU baseline: bounded sample on the host cores
    cpu = None
    if not args.no_cpu and world == 1:  # rank 0 at N = 1 only (the other ranks' host cores idle otherwise)
        r, cores, sample = cpu_reference_rate(args.workload, 15.0)
        cpu = {"value": r, "unit": "cell solves/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": "micro cell solves/sec (FP64)", "value": value, "unit": "cell solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "macro_cells": n_total, "macro_cells_per_gpu": n_local,
                   "macro_nnz": nnz, "axis_collapse": False, "cell_rtol": CELL_RTOL, "cell_atol": CELL_ATOL, "mean_pcg_iterations": mean_it,
                   "max_rel_residual": max_res, "macro_assembly_wall_ms": ms_per_step,
                   "l2": "flushed (256 MiB write) between timed steps", "parallelism": f"macro cells sharded over {world} GPU(s)"},
        "clocks": clocks, "gpu_launches": n_launch,
        "e2e": {"value": e2e_value, "unit": "cell solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": {"bound": "fp64", "kernel": "hmx_cell", "achieved": flops / (cell_avg_ms * 1e-3) / 1e12, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": flops / (cell_avg_ms * 1e-3) / 1e12 / fp64_peak, "traffic": traffic,
                     "peak_source": "FP64 DFMA microbenchmark of libhmx on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
                     "cell_kernel_ms": cell_avg_ms, "cell_kernel_share_of_step": cell_avg_ms / float(np.mean(step_ms)),
                     "algorithmic_flops_per_launch": flops, "copy_gbs_measured": copy_gbs},
        "cpu_baseline": cpu, "other_workloads": other, "wall_s_timed_region": wall,
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def quick_rate(name, device, collapse=False, world=1, cell_solver="auto"):
    import torch
    import torch.distributed as dist

    hmm = build_solver(name, world, collapse=collapse, device=device, cell_solver=cell_solver)
    hmm._ensure_solver()
    sol, d = hmm._solver, hmm._dev
    sol.set_stream(torch.cuda.current_stream().cuda_stream)
    n = d["hi"] - d["lo"]
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sol.assemble_macro_dev(n, d["cells"], hmm._msh.num_nodes, d["xyz"], hmm._pattern.nnz, d["ptr"], d["src"], d["vals"], d["S"],
                               d["it"], d["res"])  # fmt: skip
        if world > 1:
            hmm._halo_sum()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item()
    n_all = hmm._msh.num_cells
    note = ("micro axes the coefficient does not depend on solved on one layer of cubes (exact symmetry reduction, "
            "same A_hom to 1e-10; DESIGN.md 4)") if collapse else "full n^d micro cell"
    return {"desc": WORKLOADS[name]["desc"], "micro_problem": note, "macro_cells": n_all, "n_gpus": world, "ms_per_step": best,
            "cell_solves_per_s": n_all / (best * 1e-3), "cell_solver": hmm.cell_solver_used,
            "mean_pcg_iterations": float(d["it"].float().mean().item())}  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--shrink", type=int, default=1, help="divide every macro mesh axis by this (ncu --set full captures)")
    args = ap.parse_args()
    global SHRINK
    SHRINK = max(1, args.shrink)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]  # fmt: skip
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
