"""Small invocations of every hand-rolled kernel, for `compute-sanitizer --tool memcheck|racecheck|synccheck`:

    compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python tools/sanitize_targets.py [names...]

Sizes are tiny (the sanitizer slows kernels by 10-100x) but every kernel takes the code path the benchmarks use:
K1 v5 (TMA raw ring + mbarrier hand-offs + setmaxnreg roles) on RGB / L / RGBA, the N1 streaming resize, K3 bulk-fed and
cp.async-staged, the fused join (both roles + the warp-cooperative slow path), the Gaussian SSIM, the table scan's
union-find.  Results are checked against the generic kernels / the oracle so a silent corruption also fails the run."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import oracle
from kobato_b200 import _native as nat
from kobato_b200 import ops, synth
from oracle import ref_py

ctx = nat.context(0)
which = set(sys.argv[1:]) or {"phash", "resize", "ssim", "join", "gauss", "scan", "luma"}


def generic(fn, opt):
    ctx.set_option(opt, 1)
    try:
        return fn()
    finally:
        ctx.set_option(opt, 0)


if "phash" in which:
    for (h, w, c, n) in ((96, 160, 3, 40), (64, 512, 3, 12), (130, 256, 1, 9), (48, 64, 4, 20), (33, 48, 3, 7)):
        imgs = ops.synth_images_device(0, n, h, w, c, n_set=n)
        got = ops.phash_dhash_batch(imgs, want_planes=True)
        ref = generic(lambda: ops.phash_dhash_batch(imgs, want_planes=True), nat.KE_OPT_PHASH_GENERIC)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and torch.equal(got[2][0], ref[2][0]), (h, w, c)
    print("phash ok")
if "resize" in which:
    for (h, w, c, n, side) in ((96, 160, 3, 20, 64), (128, 256, 3, 9, 32), (64, 64, 1, 5, 128)):
        imgs = ops.synth_images_device(0, n, h, w, c, n_set=n)
        a = ops.gray_resize_batch(imgs, side, side, "bilinear")
        b = generic(lambda: ops.gray_resize_batch(imgs, side, side, "bilinear"), nat.KE_OPT_RESIZE_GENERIC)
        assert torch.equal(a, b), (h, w, c, side)
        ops.tile_ahash_bits(a, side // 8, 8)
    print("resize ok")
if "ssim" in which:
    for (h, w, c) in ((64, 64, 1), (40, 80, 3), (30, 300, 1), (24, 48, 4), (9, 21, 1)):
        imgs = synth.synth_images(0, 8, h, w, c, n_set=8, planted=0.5)
        bank = torch.from_numpy(imgs).cuda()
        got = ops.ssim_batch(bank, [0, 1, 2, 3], [4, 5, 6, 7]).cpu().numpy()
        v1 = generic(lambda: ops.ssim_batch(bank, [0, 1, 2, 3], [4, 5, 6, 7]).cpu().numpy(), nat.KE_OPT_SSIM_V1)
        assert np.abs(got - v1).max() <= 2e-6, (h, w, c)
    print("ssim ok")
if "gauss" in which:
    imgs = synth.synth_images(0, 6, 40, 50, 1, n_set=6, planted=0.5)
    got = ops.ssim_batch(torch.from_numpy(imgs).cuda(), [0, 1, 2], [3, 4, 5], gaussian=True).cpu().numpy()
    for k in range(3):
        assert abs(got[k] - ref_py.ssim_gaussian_of_planes(imgs[k], imgs[k + 3])) <= 1e-5
    print("gauss ok")
if "join" in which:
    h = synth.synth_hashes(9000, seed=4, planted=0.3)
    want = oracle.hamming_join(h, 8, require_band=True, threads=4)
    for mode in (1, 2, 3):
        ctx.set_option(nat.KE_OPT_JOIN_MODE, mode)
        try:
            got = ops.hamming_join(torch.from_numpy(h.view(np.int64)).cuda(), 8, require_band=True)
        finally:
            ctx.set_option(nat.KE_OPT_JOIN_MODE, 0)
        assert all(np.array_equal(g, w) for g, w in zip(got, want)), mode
    print("join ok")
if "scan" in which:
    h = synth.synth_hashes(4000, seed=9, planted=0.3)
    got = ops.scan_table(h.view(np.int64), np.arange(4000) * 3 + 1, np.arange(4000) % 977 * 1000, size_ratio=0.5)
    want = ref_py.scan_table(h, np.arange(4000) * 3 + 1, np.arange(4000) % 977 * 1000, size_ratio=0.5)
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["label"], want["label"])
    big = ops.scan_table(np.full(3000, 77, np.int64), threshold=0)
    assert big["stats"]["clusters"] == 1 and np.all(big["label"] == 0)
    print("scan ok")
if "luma" in which:
    imgs = synth.synth_images(0, 5, 32, 48, 3, n_set=5)
    got = ops.luma_planes(torch.from_numpy(imgs).cuda(), [4, 0, 2]).cpu().numpy()
    assert all(np.array_equal(got[k], oracle.to_l(imgs[i])) for k, i in enumerate((4, 0, 2)))
    print("luma ok")
torch.cuda.synchronize()
print("sanitize targets done")
