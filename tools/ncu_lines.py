"""Per-source-line executed-instruction and stall-sample totals of one kernel: joins the SASS page of an .ncu-rep with the
line table of the object file (developer tool).   python tools/ncu_lines.py REPORT KERNEL_SUBSTR OBJECT [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kname, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
execd, samples = {}, {}
first = None
for r in rows[hdr_i + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        if r and r[0] == "Kernel Name":
            break  # next launch of the same kernel
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    first = a if first is None else first
    execd[a - first] = int(r[ie])
    samples[a - first] = int(r[isamp]) if r[isamp].isdigit() else 0
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur, infunc = None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m:
        infunc = kname in m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
    if m and infunc:
        line_of[int(m.group(1), 16)] = cur
agg, sagg = collections.Counter(), collections.Counter()
for a, n in execd.items():
    agg[line_of.get(a)] += n
    sagg[line_of.get(a)] += samples[a]
tot, stot = sum(agg.values()), max(sum(sagg.values()), 1)
print(f"{kname}: {tot} warp instructions, {stot} samples")
for key, n in agg.most_common(top):
    print(f"  {100 * n / tot:5.1f}% inst  {100 * sagg[key] / stot:5.1f}% samples  {key}")
