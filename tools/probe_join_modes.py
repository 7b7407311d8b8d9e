"""Time the join at a few table sizes with the POPC kernel (mode 1) and the hybrid (mode 2)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import numpy as np, torch
from kobato_b200 import _native as nat, ops, synth
torch.cuda.set_device(0)
ctx = nat.context(0)
for n in (70000, 140000, 280000, 560000):
    h = torch.from_numpy(synth.synth_hashes(n).view(np.int64)).cuda()
    for mode in (0, 1, 2):
        ctx.set_option(nat.KE_OPT_JOIN_MODE, mode)
        ops.hamming_join_device(h, 8, require_band=True, capacity=4 * n)
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = ops.hamming_join_device(h, 8, require_band=True, capacity=4 * n); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"n={n} mode={mode}: {min(ts):.3f} ms -> {n * (n - 1) / 2 / min(ts) / 1e9:.2f}e12 pairs/s, hits {r[0].numel()}")
ctx.set_option(nat.KE_OPT_JOIN_MODE, 0)
