"""Time the 70k-hash join (the C2 step's K2) alone and right after a K1 launch."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
import numpy as np, torch
from kobato_b200 import ops, synth

torch.cuda.set_device(0)
h = torch.from_numpy(synth.synth_hashes(70000).view(np.int64)).cuda()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bank = torch.empty((N, 512, 512, 3), dtype=torch.uint8, device="cuda")
for lo in range(0, N, 8192):
    c = min(8192, N - lo)
    ops.synth_images_device(lo, c, 512, 512, 3, n_set=N, out=bank[lo:lo + c])

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for rep in range(4):
    a = ev(); r = ops.hamming_join_device(h, 8, require_band=True, capacity=140000); b = ev(); b.synchronize()
    print("join alone ms", a.elapsed_time(b), "hits", r[0].numel())
for rep in range(4):
    a0 = ev(); ops.phash_dhash_batch(bank); a = ev()
    t0 = time.perf_counter(); r = ops.hamming_join_device(h, 8, require_band=True, capacity=140000); t1 = time.perf_counter(); b = ev(); b.synchronize()
    print("k1 ms", a0.elapsed_time(a), "join after k1 ms", a.elapsed_time(b), "host s", t1 - t0)
for rep in range(3):
    a0 = ev(); p, d = ops.phash_dhash_batch(bank); a = ev()
    r = ops.hamming_join_device(p, 8, require_band=True, capacity=140000); b = ev(); b.synchronize()
    print("k1 ms", a0.elapsed_time(a), "join of k1 hashes ms", a.elapsed_time(b), "hits", r[0].numel())
from kobato_b200 import pipeline, _native as nat
mode = int(os.environ.get('KE_PROBE_JOIN_MODE', '0'))
nat.context(0).set_option(nat.KE_OPT_JOIN_MODE, mode)
for rep in range(4):
    t0 = time.perf_counter(); out = pipeline.scan(bank, threshold=8, ssim_threshold=0.9); t1 = time.perf_counter()
    print("scan wall ms", (t1 - t0) * 1e3, {k: round(v, 2) for k, v in out.stage_ms.items()})
