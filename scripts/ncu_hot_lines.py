"""Hot source lines of a kernel: warp-stall samples of the SASS page of an .ncu-rep attributed to the source lines of
the cubin's line table (nvdisasm -g); then the shared-memory wavefronts and the local-memory (spill) instructions per
source line.  usage: ncu_hot_lines.py report.ncu-rep kernel.cubin [top]"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, in_fn = {}, None, False
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][0], 16)
agg, inst, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= si or not r[si].isdigit():
        continue
    off = int(r[0], 16) - base
    key = line_of.get(off, (None, ""))[0]
    agg[key] += int(r[si])
    inst[key] += int(r[ii])
    tot += int(r[si])
print(f"total samples {tot}")
files = {}
for (key, n) in agg.most_common(top):
    text = ""
    if key:
        import glob
        import os

        if key[0] not in files:
            c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hommx_b200", "csrc", key[0]))
            files[key[0]] = open(c[0]).read().splitlines() if c else []
        if 0 < key[1] <= len(files[key[0]]):
            text = files[key[0]][key[1] - 1].strip()[:110]
    print(f"{100 * n / tot:5.1f}%  {inst[key]:>11d} inst  {key}  {text}")

wi, wid = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
wf, ideal, lcl = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= wi:
        continue
    key = line_of.get(int(r[0], 16) - base, (None, ""))[0]
    try:
        wf[key] += int(r[wi])
        ideal[key] += int(r[wid])
    except ValueError:
        pass
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    if op.startswith(("LDL", "STL")) and r[ii].isdigit():
        lcl[key] += int(r[ii])
twf = sum(wf.values())
print(f"shared-memory wavefronts: {twf} total")
for key, n in wf.most_common(12):
    print(f"{100 * n / max(twf, 1):5.1f}%  {n:>12d} (ideal {ideal[key]:>12d})  {key}")
print("local-memory instructions (spills) by line:")
for key, n in lcl.most_common(8):
    print(f"{n:>12d}  {key}")
