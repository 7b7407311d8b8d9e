"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of a bench run: the kernels of the last
resident step (from its 70 000-image K1 launch to the next one) and every kernel's total over the run.

    python tools/launches_summary.py gpurun_out/launches.csv "<the command that was profiled>" > profiles/rN_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows:
    if r is hdr or len(r) <= iv or "gpu__time_duration" not in ",".join(r):
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].strip(), 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "").strip()
    launches.append((name, v))


def short(n):
    return n if len(n) <= 110 else n[:110]


print(cmd)
print("(per-launch device times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n")
big = [i for i, (n, v) in enumerate(launches) if "phash_v5" in n and v > 5.0]  # K1 over a whole 70 000-image bank
if len(big) >= 2:
    lo, hi = big[-2], big[-1]
    step = defaultdict(float)
    for n, v in launches[lo:hi]:
        step[n if n.startswith("ke_") else "torch helper kernels (sort, index, copy ...)"] += v
    tot = sum(step.values())
    print(f"{len(big)} resident steps found (warm-up + timed); the last complete one:")
    for n, v in sorted(step.items(), key=lambda kv: -kv[1]):
        print(f"  {v:8.3f} ms  {100 * v / tot:5.1f} %  {short(n)}")
    print(f"  {tot:8.3f} ms  total kernel time of the step\n")
tot = defaultdict(lambda: [0.0, 0])
for n, v in launches:
    tot[n][0] += v
    tot[n][1] += 1
print("all launches of the run by kernel:")
for n, (v, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {v:9.3f} ms  {c:4d}x  {short(n)}")
