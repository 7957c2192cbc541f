"""Summarise an .ncu-rep (raw page) into the few metrics the design notes cite."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]  # fmt: skip
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"# kernel {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for h, u, v in zip(hdr, units, r):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.2):
            print(f"{h} [{u}] {v}")
if len(sys.argv) > 2:  # dynamic opcode mix from the source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ii = hdr.index("Instructions Executed")
    import collections

    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ii:
            continue
        s = r[1].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        ops[s.split()[0].split(".")[0]] += int(r[ii])
    tot = sum(ops.values())
    print(f"# dynamic warp-instruction mix (total {tot})")
    for op, n in ops.most_common(12):
        print(f"{op} {n} {100 * n / tot:.1f}%")
