#!/usr/bin/env python
"""bench.py -- the hot path of hommx on B200: micro cell solves/s and macro assembly wall-time.

A "step" is one pass of the hot path over the macro cells of the workload: what the reference does
in ``BaseHMM._assemble_stiffness`` + ``self._A.assemble()`` (/root/reference/src/hommx/hmm.py:
298-332, 442): per macro cell the periodic micro problems, A_hom, the local matrix, the scatter
into the macro CSR values and -- when cells are sharded over GPUs -- the sum of shared slots.

Default workload: BASELINE.json configs[3], the one the north-star target is quoted on
(LinearElasticityStratifiedHMM, rotated-fibre beam [0,1]x[0,0.4]x[0,0.1], 8^3 micro cell, 6
right-hand sides per point).  ``--workload c1|c2|c3`` select the Poisson configs.  Per-GPU work
is fixed as N grows (weak scaling): the macro mesh gets N slabs of cells; ``--scaling strong``
keeps the named mesh size (C4 80x32x8, C3 32^3) at every N, and every run reports those
strong-scaling wall times under ``other_workloads`` (``*_strong``).

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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hommx_b200 import workloads as W  # noqa: E402  (the reference's example problems, BASELINE configs[0..3])

WORKLOADS = W.WORKLOADS
CELL_RTOL, CELL_ATOL = 1e-8, 1e-10
SHRINK = 1
SCALING = "weak"
REF_SAMPLE = 256  # macro cells of the CPU arm's fixed sample


def build_solver(wl, world, collapse=False, **kw):
    return W.build_solver(wl, world, collapse=collapse, shrink=SHRINK, scaling=SCALING, rtol=CELL_RTOL, atol=CELL_ATOL, **kw)


def kernel_jobs():
    """Cell kernels the bench needs (compiled by __graft_entry__.build())."""
    return W.kernel_jobs()


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
    from oracle import npufl, ufldegree

    w = WORKLOADS[wl]
    n = w["n"]
    state = _oracle_worker.__dict__.setdefault("state", {})
    if wl not in state:
        mm = omesh.create_unit_square(n, n) if w["dim"] == 2 else omesh.create_unit_cube(n, n, n)
        degree = ufldegree.form_degree(W.coefficient(wl, ufldegree)[0], w["dim"])
        macro = omesh.create_rectangle(*w["box"], list(w["cells"])) if w["dim"] == 2 else omesh.create_box(*w["box"], list(w["cells"]))
        A, Dn = W.coefficient(wl, npufl)
        Dt = None
        if Dn is not None:
            Dt = lambda x: np.asarray(Dn(np.asarray(x, float)))[..., 0]  # noqa: E731
        state[wl] = (ho.MicroCell(mm, "poisson" if w["kind"] == 0 else "elasticity", degree), macro, A, Dt)
    mic, macro, A, Dt = state[wl]
    out = []
    for c in idx:
        verts = macro.x[macro.cells[c]]
        out.append(ho.local_stiffness_literal(mic, A, verts, w["eps"], Dt))  # hmm.py:334-369, literally
    return len(out)


class CpuArm:
    """The restated reference algorithm (oracle, kind "port") on all host cores over a FIXED sample of macro cells:
    drawn once with a fixed seed, the same cells every step, split evenly over the worker processes."""

    def __init__(self, wl, n_sample=REF_SAMPLE, procs=None):
        import multiprocessing as mp

        self.wl, self.procs = wl, procs or (os.cpu_count() or 1)
        w = WORKLOADS[wl]
        n_total = int(np.prod(w["cells"])) * (2 if w["dim"] == 2 else 6)
        n_sample = max(self.procs, min(n_sample, n_total))
        cells = np.random.default_rng(0).choice(n_total, size=n_sample, replace=False)
        self.chunks = [(wl, [int(c) for c in cells[p :: self.procs]]) for p in range(self.procs)]
        self.n_sample = n_sample
        self.pool = mp.get_context("fork").Pool(self.procs)
        self.pool.map(_oracle_worker, [(wl, [int(cells[0])])] * self.procs)  # builds the micro cell once per process

    def step(self):
        """cells/s of one pass over the sample"""
        t0 = time.perf_counter()
        done = sum(self.pool.map(_oracle_worker, self.chunks))
        return done / (time.perf_counter() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()

    def sample(self, rates):
        spread = (max(rates) - min(rates)) / np.mean(rates) if len(rates) > 1 else 0.0
        return (f"{self.n_sample} macro cells of {self.wl} (seed 0, the same cells every step), literal n_b-corrector algorithm "
                f"of hmm.py:334-369 with sparse LU, {len(rates)} timed pass(es), spread (max-min)/mean {spread:.1%}")


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
    t_all = time.perf_counter()
    arm = CpuArm(args.workload)
    for _ in range(min(args.warmup, 1)):  # one untimed pass warms caches and the page cache; more would only cost minutes
        arm.step()
    rates = [arm.step() for _ in range(args.steps)]
    arm.close()
    value = float(np.mean(rates))
    n_cells = int(np.prod(W.macro_cells(args.workload, world, SHRINK, SCALING))) * (2 if w["dim"] == 2 else 6)
    line = {
        "impl": "reference", "metric": "micro cell solves/sec (FP64)", "value": value, "unit": "cell solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_cells / value * 1e3, "higher_is_better": True, "scaling": SCALING,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "macro_cells": n_cells,
                   "note": "each step = one pass over the fixed sample; ms_per_step extrapolated from the sample to the whole workload",
                   "rates_per_step": rates},
        "cpu_baseline": {"value": value, "unit": "cell solves/s", "cores": arm.procs, "kind": "port", "sample": arm.sample(rates)},
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
    nnz, n_nodes = d["nnz"], d["n_nodes"]  # this rank's CSR slots and macro nodes (local numbering)
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
        sol.assemble_macro_dev(n_local, d["cells"], n_nodes, d["xyz"], 0, None, None, None, d["S"], d["it"], d["res"])
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
    h_cells, h_xyz = pin(d["shard"].cells), pin(hmm._msh.x[d["shard"].nodes])
    h_ptr, h_src = pin(d["ptr"].cpu().numpy()), pin(d["src"].cpu().numpy())
    h_vals = torch.empty(nnz, dtype=torch.float64).pin_memory()
    h_it = torch.empty(n_local, dtype=torch.int32).pin_memory()
    h_res = torch.empty(n_local, dtype=torch.float64).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_cells, h_xyz, h_ptr, h_src))
    d2h = sum(t.numel() * t.element_size() for t in (h_vals, h_it, h_res))

    def step_e2e():
        if world == 1:  # the host-buffer C-ABI entry point: copies in, kernels, copies out, sync
            sol._check(sol.lib.hmx_assemble_macro(sol._h, n_local, h_cells.data_ptr(), n_nodes, h_xyz.data_ptr(), nnz, h_ptr.data_ptr(),
                                       h_src.data_ptr(), h_vals.data_ptr(), None, h_it.data_ptr(), h_res.data_ptr()))  # fmt: skip
        else:  # sharded: the halo sum sits between the kernels and the copy-out
            d["cells"].copy_(h_cells, non_blocking=True)
            d["xyz"].copy_(h_xyz, non_blocking=True)
            d["ptr"].copy_(h_ptr, non_blocking=True)
            d["src"].copy_(h_src, non_blocking=True)
            step_resident()
            h_vals.copy_(d["vals"][:nnz], non_blocking=True)
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
        assert np.array_equal(h_vals.numpy(), d["vals"][:nnz].cpu().numpy())

    # ---- roofline of the dominant kernel (the cell kernel), this rank ---------------------------
    flops = algorithmic_flops(w, n_local * args.steps, rhs_its) / args.steps
    cell_avg_ms = float(np.mean(cell_ms))
    fp64_peak, copy_gbs = native.measure_peaks(local_rank)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)

    # sharded == unsharded, proven inside the run: shared-row slots recomputed from ALL contributing cells on rank 0
    halo_check = halo_self_check(hmm, local_rank) if world > 1 else None

    # secondary workloads: parity-test configs, a coefficient that varies along all micro axes, the axis-collapsed
    # form, and the STRONG-scaling wall times (the named mesh sizes of BASELINE configs[2..3] whatever N is).
    # At N > 1 only the sharded entries are repeated (every rank takes part).
    other = {}
    if not args.no_extra:
        # cell_solver "auto": small hard elasticity cells (the collapsed C4 cell) are factorised directly (K5)
        todo = [("c4", False, "auto", "strong"), ("c3", False, "auto", "strong"), (args.workload, True, "auto", "weak")]
        if world == 1:
            todo += [("c2", False, "auto", "weak"), ("c3", False, "auto", "weak"), ("c4s", False, "auto", "weak"),
                     ("c4", True, "pcg", "weak"), ("c3", True, "auto", "weak"), ("c2", True, "auto", "weak"),
                     # the cluster-resident assembled stencil (csrc/hmx_cell_cluster.cuh): opt-in on the 8^3 cell, the
                     # default where the cell exceeds one SM (10^3), next to the matrix-free kernel on the same cell
                     ("c4", False, "cluster", "weak"), ("c4n10", False, "auto", "weak"), ("c4n10", False, "pcg", "weak"),
                     ("c4s", False, "pcg", "weak")]  # fmt: skip
        for name, collapse, how, scaling in todo:
            if name == args.workload and not collapse and scaling == SCALING and how == "auto":
                continue
            key = (name + ("_axis_collapsed" if collapse else "") + ("_pcg" if how == "pcg" else "") + ("_cluster" if how == "cluster" else "")
                   + ("_strong" if scaling == "strong" else ""))  # fmt: skip
            try:
                other[key] = quick_rate(name, local_rank, collapse, world, how, scaling)
            except Exception as e:  # never lose the headline line
                other[key] = {"error": str(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # CPU baseline: bounded sample on the host cores
    cpu = None
    if not args.no_cpu and world == 1:  # rank 0 at N = 1 only (the other ranks' host cores idle otherwise)
        arm = CpuArm(args.workload)
        rates = [arm.step()]
        arm.close()
        cpu = {"value": rates[0], "unit": "cell solves/s", "cores": arm.procs, "kind": "port", "sample": arm.sample(rates)}
    line = {
        "metric": "micro cell solves/sec (FP64)", "value": value, "unit": "cell solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
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
                     "algorithmic_flops_per_launch": flops, "copy_gbs_measured": copy_gbs,
                     "preconditioner": precond_note(w),
                     "frac_incl_coarse_solve": (flops + coarse_flops(w, n_local, rhs_its / args.steps)) / (cell_avg_ms * 1e-3) / 1e12 / fp64_peak},
        "cpu_baseline": cpu, "halo_check": halo_check, "other_workloads": other, "wall_s_timed_region": wall,
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def coarse_dofs(w):
    """Unknowns of the coarse space of the two-level PCG for workload ``w`` (0 = block Jacobi)."""
    from hommx_b200 import codegen, native
    from hommx_b200 import ufl as pufl

    name = next(k for k, v in WORKLOADS.items() if v is w)
    A, Dt = W.coefficient(name, pufl)
    return native.coarse_dofs(codegen.build_program(A, w["dim"], w["kind"], Dt), w["n"])


def coarse_flops(w, n_pts, rhs_iterations):
    """FLOPs of the dense coarse solves of the two-level preconditioner: one symmetric NCD x NCD matrix-vector product
    per right-hand side and iteration, NCD^3 for the inversion per macro point (reported separately from the SURVEY 8d
    count, which knows a Jacobi apply only)."""
    ncd = coarse_dofs(w)
    return rhs_iterations * 2.0 * ncd * ncd + n_pts * float(ncd) ** 3


def precond_note(w):
    ncd = coarse_dofs(w)
    return ("block Jacobi" if ncd == 0 else
            f"block Jacobi + additive coarse correction, {ncd} coarse unknowns inverted exactly per macro point (csrc/hmx_cell_coarse.cuh)")


def halo_self_check(hmm, device, n_slots=256):
    """Rank 0 recomputes a sample of the CSR slots it shares with other ranks from ALL contributing macro cells
    (its own and its neighbours', solved again on this GPU) and compares with what the sharded assembly + halo sum
    left in its value array: the parity proof of the multi-GPU path, inside the bench run itself."""
    import torch

    d = hmm._dev
    sh = d["shard"]
    if hmm._rank != 0:
        return None
    shared_local = sh.shared[sh.shared < sh.nnz]
    if len(shared_local) == 0:
        return {"slots": 0, "max_rel_diff": 0.0}
    pick = shared_local[:: max(1, len(shared_local) // n_slots)][:n_slots]
    gslots = sh.slots[pick]
    sm = hmm._pattern.slot_map
    cells = np.nonzero(np.isin(sm, gslots).any(axis=1))[0]
    nodes, inv = np.unique(hmm._msh.cells[cells], return_inverse=True)
    tdev = d["vals"].device
    t_cells = torch.as_tensor(inv.reshape(len(cells), -1).astype(np.int32), device=tdev)
    t_xyz = torch.as_tensor(np.ascontiguousarray(hmm._msh.x[nodes]), device=tdev)
    nb2 = hmm._num_basis_functions_per_cell**2
    S = torch.zeros((len(cells), nb2), dtype=torch.float64, device=tdev)
    hmm._solver.assemble_macro_dev(len(cells), t_cells, len(nodes), t_xyz, 0, None, None, None, S, None, None)
    hmm._solver.sync()
    S = S.cpu().numpy()
    want = np.zeros(len(gslots))
    pos = {int(g): k for k, g in enumerate(gslots)}
    for ci, c in enumerate(cells):
        for e, g in enumerate(sm[c]):
            k = pos.get(int(g))
            if k is not None:
                want[k] += S[ci, e]
    got = d["vals"][torch.as_tensor(pick, device=tdev)].cpu().numpy()
    scale = np.abs(want).max()
    return {"slots": int(len(gslots)), "contributing_cells": int(len(cells)),
            "cells_of_other_ranks": int(((cells < sh.lo) | (cells >= sh.hi)).sum()),
            "max_rel_diff": float(np.abs(got - want).max() / scale)}


def quick_rate(name, device, collapse=False, world=1, cell_solver="auto", scaling="weak"):
    import torch
    import torch.distributed as dist

    global SCALING
    saved, SCALING = SCALING, scaling
    try:
        hmm = build_solver(name, world, collapse=collapse, device=device, cell_solver=cell_solver)
    finally:
        SCALING = saved
    hmm._ensure_solver()
    sol, d = hmm._solver, hmm._dev
    sol.set_stream(torch.cuda.current_stream().cuda_stream)
    n = d["hi"] - d["lo"]
    best = 1e30
    for _ in range(3 if scaling == "strong" else 4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sol.assemble_macro_dev(n, d["cells"], d["n_nodes"], d["xyz"], d["nnz"], d["ptr"], d["src"], d["vals"], d["S"],
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
    return {"desc": WORKLOADS[name]["desc"], "micro_problem": note, "macro_cells": n_all, "n_gpus": world, "scaling": scaling,
            "macro_assembly_wall_ms": best, "ms_per_step": best,
            "cell_solves_per_s": n_all / (best * 1e-3), "cell_solver": hmm.cell_solver_used,
            "ctas_per_cluster": sol.info["cluster"], "resident_clusters": sol.info["resident_clusters"],
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one slab of macro cells per GPU (default); strong: the named mesh size of the config at every N")
    args = ap.parse_args()
    global SHRINK, SCALING
    SHRINK = max(1, args.shrink)
    SCALING = args.scaling
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
