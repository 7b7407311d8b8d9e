"""Top SASS instructions of a kernel by stall samples, with the dominant stall reasons.

    python tools/ncu_hot.py <source-page.csv | report.ncu-rep> <kernel-regex> [top]
"""
import csv
import subprocess
import sys

src, kre = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if src.endswith(".ncu-rep"):
    txt = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
else:
    txt = open(src).read()
rows = list(csv.reader(txt.splitlines()))
s = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"][0]
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[s + 2:] if len(r) > ix["stall_wait"] and r[ix["Address"]].startswith("0x") or (len(r) > 5 and r[0][:1].isdigit())]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"total samples {tot}, instructions {len(data)}")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    sm = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{i:5d} {100 * sm / tot:5.1f}% n={int(r[ix['Instructions Executed']] or 0):>10d}  {r[ix['Source']][:70]:70s} " +
          " ".join(f"{h}:{v}" for v, h in why if v))
