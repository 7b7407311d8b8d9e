"""Config C5 (10 M hashes, 5e13 pairs) through the SCANNER-LEVEL API of the drop-in, in ONE process over every visible
GPU — what a Qt worker thread of the reference would call (DupViewModel(scanner_factory=DuplicateScanner)):

    python tools/check_scanner_c5.py [--n 10000000] [--json out.json]

`DuplicateScanner.build_clusters_from_columns(file_id, phash, size)` runs ke_scan_table_host on the process-wide
multi-device context: bucket statistics, the join fanned over all devices, gates, union-find on the device; Python builds
cluster objects for the members only.  Reported: wall time of the whole call, of the library call inside it (the join
dominates), and of the Python part; the edges of three row stripes are checked against the CPU oracle (all pairs within
the threshold that share a band), and the components against a host union-find of the returned edges."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "kobato-eyes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--json", default=None)
args = ap.parse_args()

import oracle
from kobato_b200 import _native as nat
from kobato_b200 import ops, synth
from kobato_b200.dup import scanner as kscanner

n = args.n
h = synth.synth_hashes(n)
ids = np.arange(1, n + 1, dtype=np.int64) * 3  # distinct, ascending (ORDER BY f.id)
rng = np.random.default_rng(1)
sizes = rng.integers(10_000, 8_000_000, n).astype(np.int64)
group = nat.group()
print(f"devices: {group.devices}", flush=True)
warm = min(n, 1_500_000)  # enough pairs for the fan to touch EVERY device: module load, pools and scratch are first-call costs
ops.scan_table(h[:warm].view(np.int64), ids[:warm], sizes[:warm])

lib_s = []
orig = ops.scan_table


def timed_scan(*a, **kw):
    t0 = time.perf_counter()
    out = orig(*a, want_edges=True, **{k: v for k, v in kw.items() if k != "want_edges"})
    lib_s.append(time.perf_counter() - t0)
    return out


made = []


def make_file(row):
    made.append(row)
    return kscanner.DuplicateFile(file_id=int(ids[row]), path=Path(f"d/{row}.jpg"), size=int(sizes[row]), width=64, height=64,
                                  phash=int(h[row]))


cfg = kscanner.DuplicateScanConfig(hamming_threshold=8, size_ratio=0.5)
# (1) the array-level result: no per-row Python object at all
t0 = time.perf_counter()
table = kscanner.DuplicateScanner(cfg, scan_table=timed_scan).scan_columns(ids, h.view(np.int64), sizes)
first_wall = time.perf_counter() - t0  # includes growing the scratch buffers to this table size
lib_s.pop()
t0 = time.perf_counter()
table = kscanner.DuplicateScanner(cfg, scan_table=timed_scan).scan_columns(ids, h.view(np.int64), sizes)
table_wall = time.perf_counter() - t0
table_lib = lib_s.pop()
t0 = time.perf_counter()
first = [table.cluster(c, make_file) for c in range(min(1000, len(table)))]  # what a view pages in
page_s = time.perf_counter() - t0
made.clear()
# (2) the reference's return type: every cluster materialised
keep = {}
sc = kscanner.DuplicateScanner(cfg, scan_table=lambda *a, **kw: keep.setdefault("scan", timed_scan(*a, **kw)))
t0 = time.perf_counter()
clusters = sc.build_clusters_from_columns(ids, h.view(np.int64), sizes, make_file=make_file)
wall = time.perf_counter() - t0
scan = keep["scan"]
ei, ej, ed = scan["edges"]

oracle.build()
ok = True
checked = 0
for lo in (0, n // 2, n - 150_000):
    wi, wj, wd = oracle.hamming_join(h, 8, require_band=True, threads=os.cpu_count() or 1, row_begin=lo, row_end=lo + 128)
    sa, sb = sizes[wi.astype(np.int64)], sizes[wj.astype(np.int64)]
    gate = np.minimum(sa, sb) / np.maximum(sa, sb) >= 0.5
    wi, wj, wd = wi[gate], wj[gate], wd[gate]
    sel = (ei >= lo) & (ei < lo + 128)
    ok = ok and np.array_equal(ei[sel], wi) and np.array_equal(ej[sel], wj) and np.array_equal(ed[sel], wd)
    checked += len(wi)
members, offsets = ops.cluster_pairs_csr(ei.astype(np.int64), ej.astype(np.int64))
comp_ok = np.array_equal(members, scan["index"]) and np.array_equal(offsets, scan["offsets"])
pairs = n * (n - 1) // 2
report = {"n": n, "devices": len(group.devices),
          "scan_columns": {"first_call_wall_s": round(first_wall, 3), "wall_s": round(table_wall, 3), "library_s": round(table_lib, 3),
                           "python_s": round(table_wall - table_lib, 3), "pairs_per_s": pairs / table_lib,
                           "clusters": len(table), "first_1000_clusters_materialised_s": round(page_s, 3)},
          "build_clusters_from_columns": {"wall_s": round(wall, 3), "library_s": round(lib_s[0], 3),
                                          "python_s": round(wall - lib_s[0], 3)},
          "stats": scan["stats"],
          "clusters": len(clusters), "cluster_objects_built_for_rows": len(made),
          "edge_stripes_equal_oracle": bool(ok), "edges_in_stripes": int(checked), "components_equal_host_union_find": bool(comp_ok)}
print(json.dumps(report))
if args.json:
    Path(args.json).write_text(json.dumps(report, indent=1))
sys.exit(0 if ok and comp_ok else 1)
