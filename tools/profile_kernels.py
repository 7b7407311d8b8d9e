"""Launch each hot-path kernel a few times at moderate size (for ncu / timing sweeps).

    python tools/profile_kernels.py [--images 4096] [--hashes 262144] [--pairs 100000] [--reps 2]
Prints CUDA-event times per kernel.  Used under `ncu` to produce profiles/*.csv.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kobato_b200 import ops, synth  # noqa: E402


def timed(fn, reps):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        out.append(a.elapsed_time(b))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=4096)
    ap.add_argument("--hashes", type=int, default=262144)
    ap.add_argument("--pairs", type=int, default=100000)
    ap.add_argument("--bank", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    only = set(args.only.split(",")) if args.only else {"phash", "join", "ssim", "n1"}
    if "phash" in only:
        bank = ops.synth_images_device(0, args.images, 512, 512, 3, n_set=args.images)
        ops.phash_dhash_batch(bank[:64])
        t = timed(lambda: ops.phash_dhash_batch(bank), args.reps)
        gbs = args.images * 786448 / (min(t) * 1e-3) / 1e9
        print(f"phash: {args.images} images 512x512x3: ms={t} -> {args.images / (min(t) * 1e-3):.3e} img/s, {gbs:.1f} GB/s")
        del bank
    if "phash_large" in only:  # the streaming kernel at one CTA per SM, resample fragments in registers (1024 / 2048 px)
        for (h, w) in ((1024, 1024), (1536, 2048)):
            n = max(64, int(3e9) // (h * w * 3))
            bank = ops.synth_images_device(0, n, h, w, 3, n_set=n)
            ops.phash_dhash_batch(bank[:8])
            t = timed(lambda: ops.phash_dhash_batch(bank), args.reps)
            gbs = n * (h * w * 3 + 16) / (min(t) * 1e-3) / 1e9
            print(f"phash: {n} images {w}x{h}x3: ms={t} -> {n / (min(t) * 1e-3):.3e} img/s, {gbs:.1f} GB/s")
            del bank
    if "n1" in only:
        bank = ops.synth_images_device(0, args.images, 512, 512, 3, n_set=args.images)
        for side, grid, tile in ((64, 8, 8), (32, 4, 8), (128, 16, 8)):
            def run():
                planes = ops.gray_resize_batch(bank, side, side, "bilinear")
                return ops.tile_ahash_bits(planes, grid, tile)
            run()
            t = timed(run, args.reps)
            gbs = args.images * 786432 / (min(t) * 1e-3) / 1e9
            print(f"n1 tile-aHash {grid}x{tile} ({side}x{side}): {args.images} images 512x512x3: ms={t} -> "
                  f"{args.images / (min(t) * 1e-3):.3e} img/s, {gbs:.1f} GB/s")
        planes = ops.gray_resize_batch(bank, 128, 128, "bilinear")
        g = torch.Generator().manual_seed(2)
        ia = torch.randint(0, args.images, (200000,), generator=g).cuda()
        ib = torch.randint(0, args.images, (200000,), generator=g).cuda()
        t = timed(lambda: ops.plane_sad_pairs(planes, ia, ib), args.reps)
        print(f"n1 pixel SAD 128x128: 200000 pairs: ms={t} -> {200000 / (min(t) * 1e-3):.3e} pairs/s, "
              f"{200000 * 32768 / (min(t) * 1e-3) / 1e9:.1f} GB/s")
        del bank, planes
    if "join" in only:
        h = torch.from_numpy(synth.synth_hashes(args.hashes).view(np.int64)).cuda()
        ops.hamming_join(h[:4096], 8)
        t = timed(lambda: ops.hamming_join(h, 8), args.reps)
        pairs = args.hashes * (args.hashes - 1) // 2
        print(f"join: {args.hashes} hashes T=8: ms={t} -> {pairs / (min(t) * 1e-3):.3e} pairs/s (includes D2H of hits)")
    if "ssim" in only:
        crops = ops.synth_images_device(0, args.bank, 256, 256, 1, n_set=args.bank, planted=0.5)
        g = torch.Generator().manual_seed(1)
        ia = torch.randint(0, args.bank, (args.pairs,), generator=g).cuda()
        ib = torch.randint(0, args.bank, (args.pairs,), generator=g).cuda()
        ops.ssim_batch(crops, ia[:256], ib[:256])
        t = timed(lambda: ops.ssim_batch(crops, ia, ib), args.reps)
        gbs = args.pairs * 131080 / (min(t) * 1e-3) / 1e9
        print(f"ssim: {args.pairs} pairs 256x256: ms={t} -> {args.pairs / (min(t) * 1e-3):.3e} pairs/s, {gbs:.1f} GB/s")


if __name__ == "__main__":
    main()
