"""Aggregate an ncu SASS source page (per-instruction counts) by CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin.sass from nvdisasm -g -c> [launch_index]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, sass = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else -1
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = secs[which]
e = secs[secs.index(s) + 1] if secs.index(s) + 1 < len(secs) else len(rows)
kname = rows[s][1]
hdr = rows[s + 1]
ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = [r for r in rows[s + 2:e] if len(r) > ii]
base = int(data[0][ia], 16)
# map function offsets -> source line from nvdisasm -g output
fn = re.search(r"ke_\w+_kernel\w*", kname).group(0)
tmpl = re.findall(r"\((?:int|bool)\)(\d+)", kname)
if not tmpl and "<" in kname:  # newer ncu prints "kernel<3, 1, 1, 16>(...)": ints and bools alike
    tmpl = re.findall(r"\d+", kname[kname.index(fn) + len(fn):].split(">")[0])
line_of = {}
cur = None
infn = False
want = None
for ln in open(sass):
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        name = m.group(1)
        infn = (fn + "I" in name or fn + "E" in name) and re.search("".join(f"L[ib]{t}E" for t in tmpl) + "E", name) is not None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0])
tot = 0
for r in data:
    off = int(r[ia], 16) - base
    n = int(r[ii] or 0)
    sm = int(r[isamp] or 0)
    key = line_of.get(off, ("?", 0))
    agg[key][0] += n
    agg[key][1] += sm
    tot += n
tots = sum(v[1] for v in agg.values())
print(f"{kname}: {tot} warp instructions, {tots} samples")
for key, (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{n:>13d} {100 * n / tot:5.1f}%  samples {100 * sm / max(tots, 1):5.1f}%  {key[0]}:{key[1]}")
