"""ke_phash_batch_host (the call behind core.fastsig.compute_signatures_mp) from PAGEABLE and from page-locked host memory:
images/s and GB/s of the whole call (host->device copies, kernels, hashes back), next to the bare pinned copy.

    python tools/probe_host_path.py [n_images=12288]      # KE_STAGE_THREADS: default max(2, min(8, cores / 2))
"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import numpy as np
import torch

from kobato_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
bank = ops.synth_images_device(0, n, 512, 512, 3, n_set=n)
pageable = bank.cpu().numpy().copy()
pinned_t = torch.empty(bank.shape, dtype=torch.uint8, pin_memory=True)
pinned_t.copy_(bank)
pinned = pinned_t.numpy()
want = ops.phash_dhash_batch(bank)[0].cpu().numpy()
gb = pageable.nbytes / 1e9


def run(arr, label):
    ops.phash_dhash_batch(arr[:256])
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ph, _ = ops.phash_dhash_batch(arr)
        best = min(best, time.perf_counter() - t0)
    assert np.array_equal(ph, want)
    print(f"{label}: {n} images in {best * 1e3:.1f} ms -> {n / best:.0f} img/s, {gb / best:.1f} GB/s "
          f"(KE_STAGE_THREADS={os.environ.get('KE_STAGE_THREADS', 'default')})", flush=True)


a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
bank.copy_(pinned_t, non_blocking=True)
b.record()
b.synchronize()
print(f"bare pinned H2D: {gb / (a.elapsed_time(b) * 1e-3):.1f} GB/s")
run(pinned, "page-locked source (DMA in place)")
run(pageable, "pageable source (staged through pinned buffers)")
