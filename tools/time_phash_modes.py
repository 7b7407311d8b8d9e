"""Time K1 on 'L', RGB and RGBA banks of 512x512 images (CUDA events, best of 3)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import torch
from kobato_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for c in (1, 3, 4):
    bank = ops.synth_images_device(0, n, 512, 512, c, n_set=n)
    ops.phash_dhash_batch(bank[:64])
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.phash_dhash_batch(bank); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    print(f"c={c}: {n} images {t:.3f} ms -> {n / t * 1e3:.3e} img/s, {n * 512 * 512 * c / t / 1e6:.0f} GB/s")
    del bank
