"""Soak test: K1 v5 / the N1 resize against the generic kernels on random batch sizes and geometries."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import numpy as np
import torch

from kobato_b200 import _native as nat
from kobato_b200 import ops

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(123)
ctx = nat.context(0)
shapes = [(512, 512, 3), (512, 512, 1), (512, 512, 4), (96, 160, 3), (300, 256, 3), (47, 512, 4), (1000, 64, 3), (33, 512, 1)]
bad = 0
for it in range(rounds):
    h, w, c = shapes[it % len(shapes)]
    n = int(rng.integers(1, 1200 if h * w <= 512 * 512 else 200))
    imgs = ops.synth_images_device(int(rng.integers(0, 1 << 20)), n, h, w, c, n_set=1 << 30)
    got = ops.phash_dhash_batch(imgs, want_planes=True)
    ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 1)
    try:
        ref = ops.phash_dhash_batch(imgs, want_planes=True)
    finally:
        ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 0)
    ok = torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and torch.equal(got[2][0], ref[2][0]) and torch.equal(got[2][1], ref[2][1])
    side = (32, 64, 128)[it % 3]
    a = ops.gray_resize_batch(imgs, side, side, "bilinear")
    ctx.set_option(nat.KE_OPT_RESIZE_GENERIC, 1)
    try:
        b = ops.gray_resize_batch(imgs, side, side, "bilinear")
    finally:
        ctx.set_option(nat.KE_OPT_RESIZE_GENERIC, 0)
    ok2 = torch.equal(a, b)
    if not (ok and ok2):
        bad += 1
        print("MISMATCH", it, (h, w, c), n, ok, ok2)
print(f"soak: {rounds} rounds, {bad} mismatches")
sys.exit(1 if bad else 0)
