"""K2 at the table sizes of the scan steps: 70 k (1-GPU step), a rank's share of 140 k / 280 k / 560 k (2 / 4 / 8-GPU steps),
1 M (C3).  Random tables with 5 % planted near-duplicates, T = 8, band predicate, best of 5."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import numpy as np
import torch

from kobato_b200 import ops, synth

for n, parts in ((70_000, 1), (140_000, 2), (280_000, 4), (560_000, 8), (1_000_000, 1)):
    h = torch.from_numpy(synth.synth_hashes(n).view(np.int64)).cuda()
    ops.hamming_join_device(h, 8, require_band=True, part_index=0, part_count=parts, capacity=1 << 22)
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.hamming_join_device(h, 8, require_band=True, part_index=0, part_count=parts, capacity=1 << 22)
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    pairs = n * (n - 1) / 2 / parts
    print(f"n={n} part 0/{parts}: {best:.3f} ms -> {pairs / (best * 1e-3):.3e} pairs/s")
