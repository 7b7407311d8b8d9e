"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI of
libkobato_b200.so and is compared with the CPU oracle on the same seeded inputs.
Bars: bit-exact for hashes, planes and candidate pairs; |delta| <= 1e-5 for SSIM."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from kobato_b200 import ops, synth
from oracle import ref_py

pytestmark = pytest.mark.gpu
U64 = (1 << 64) - 1
SSIM_TOL = 1e-5  # north_star: SSIM within 1e-5 absolute


def _torch():
    import torch

    return torch


def _u64(a):
    return np.asarray(a).astype(np.int64).view(np.uint64)


# --------------------------------------------------------------------------------- K2 join


@pytest.mark.parametrize("threshold", [0, 4, 8, 12])
@pytest.mark.parametrize("require_band", [False, True])
def test_join_matches_oracle(threshold, require_band):
    h = synth.synth_hashes(6007, seed=42 + threshold, planted=0.2)
    want = oracle.hamming_join(h, threshold, require_band=require_band, threads=4)
    got_host = ops.hamming_join(h, threshold, require_band=require_band)
    torch = _torch()
    got_dev = ops.hamming_join(torch.from_numpy(h.view(np.int64)).cuda(), threshold, require_band=require_band)
    for got in (got_host, got_dev):
        assert len(got[0]) == len(want[0])
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    if threshold >= 4:
        assert len(want[0]) > 100  # the planted near-duplicates are actually found


def test_join_threshold_64_emits_every_pair():
    h = synth.synth_hashes(700, seed=7)
    i, j, d = ops.hamming_join(h, 64)
    assert len(i) == 700 * 699 // 2
    wi, wj, wd = oracle.hamming_join(h, 64)
    assert np.array_equal(i, wi) and np.array_equal(j, wj) and np.array_equal(d, wd)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 511, 513, 2049])
def test_join_edge_sizes(n):
    h = synth.synth_hashes(max(n, 1), seed=99, planted=0.3)[:n]
    if n >= 2:
        h[-1] = h[0]  # a guaranteed hit across the whole range
    got = ops.hamming_join(h, 6)
    want = oracle.hamming_join(h, 6)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_join_other_band_geometries_and_allow_mask():
    h = synth.synth_hashes(5000, seed=5, planted=0.25)
    for bits, count in ((8, 8), (12, 5), (16, 2), (64, 1), (1, 3)):
        got = ops.hamming_join(h, 10, require_band=True, band_bits=bits, band_count=count)
        want = oracle.hamming_join(h, 10, require_band=True, band_bits=bits, band_count=count, threads=4)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), (bits, count)
    # allow mask: forbid band 0 for even items, band 1 for everyone
    allow = np.full(5000, 0b1101, np.uint64)
    allow[::2] &= np.uint64(0b1100)
    got = ops.hamming_join(h, 10, require_band=True, band_allow=allow)
    wi, wj, wd = oracle.hamming_join(h, 10, require_band=False, threads=4)
    keep = []
    for a, b in zip(wi, wj):
        x = int(h[a]) ^ int(h[b])
        ok = any(((x >> (16 * band)) & 0xFFFF) == 0 and (int(allow[a]) >> band) & 1 and (int(allow[b]) >> band) & 1
                 for band in range(4))
        keep.append(ok)
    keep = np.array(keep, bool)
    assert np.array_equal(got[0], wi[keep]) and np.array_equal(got[1], wj[keep]) and np.array_equal(got[2], wd[keep])


def test_join_partitions_cover_the_triangle_exactly_once():
    h = synth.synth_hashes(9001, seed=3, planted=0.2)
    full = ops.hamming_join(h, 8)
    for parts in (2, 3, 8):
        pieces = [ops.hamming_join(h, 8, part_index=p, part_count=parts) for p in range(parts)]
        i = np.concatenate([p[0] for p in pieces])
        j = np.concatenate([p[1] for p in pieces])
        d = np.concatenate([p[2] for p in pieces])
        order = np.lexsort((j, i))
        assert np.array_equal(i[order], full[0]) and np.array_equal(j[order], full[1]) and np.array_equal(d[order], full[2])


def test_join_capacity_regrows_without_truncation():
    h = np.zeros(3000, np.uint64)  # every pair is a hit: 4.5 M pairs
    i, j, d = ops.hamming_join(h, 0, capacity=1000)
    assert len(i) == 3000 * 2999 // 2 and d.max() == 0
    assert np.all(i < j)


def test_join_rejects_bad_arguments():
    h = synth.synth_hashes(100)
    with pytest.raises(ValueError):
        ops.hamming_join(h, 65)
    with pytest.raises(ValueError):
        ops.hamming_join(h, 8, require_band=True, band_bits=16, band_count=5)
    with pytest.raises(ValueError):
        ops.hamming_join(h, 8, require_band=True, band_bits=0, band_count=4)


def test_join_large_scale_properties():
    """1 M hashes (config C3): properties that need no CPU all-pairs — every emitted pair is
    correct, planted duplicates are all found, and an oracle stripe of rows matches exactly."""
    n = 1_000_000
    h = synth.synth_hashes(n, planted=0.05)
    i, j, d = ops.hamming_join(h, 8)
    assert np.all(i < j)
    x = h[i] ^ h[j]
    pc = np.array([int(v).bit_count() for v in x[:200000]], np.uint8)
    assert np.array_equal(pc, d[:200000]) and d.max() <= 8
    # stripes of rows against the CPU oracle
    for lo in (0, 499_000, 999_000):
        wi, wj, wd = oracle.hamming_join(h, 8, row_begin=lo, row_end=lo + 1000, threads=8)
        sel = (i >= lo) & (i < lo + 1000)
        assert np.array_equal(i[sel], wi) and np.array_equal(j[sel], wj) and np.array_equal(d[sel], wd)
    # planted copies with <= 8 flips must all be present (as (src, copy) pairs)
    n_base = n - (n * 50) // 1000
    t = np.arange(n_base, n, dtype=np.uint64)
    src = (synth._mix(synth.SEED, t, 1) % np.uint64(n_base)).astype(np.int64)
    dist = np.array([int(a ^ b).bit_count() for a, b in zip(h[src[:5000]], h[n_base:n_base + 5000])])
    found = set(zip(i.tolist(), j.tolist()))
    for k in np.flatnonzero(dist <= 8):
        assert (int(src[k]), int(n_base + k)) in found


# --------------------------------------------------------------------------------- K1 pHash


def _hashes_of(arrs):
    """phash/dhash per image through both the device-tensor and the host-buffer entry points."""
    torch = _torch()
    batch = np.stack(arrs)
    t = torch.from_numpy(batch).cuda()
    ph, dh, mg, (p32, p98) = ops.phash_dhash_batch(t, want_margin=True, want_planes=True)
    hph, hdh = ops.phash_dhash_batch(batch)
    assert np.array_equal(ph.cpu().numpy(), hph) and np.array_equal(dh.cpu().numpy(), hdh)
    return _u64(ph.cpu().numpy()), _u64(dh.cpu().numpy()), mg.cpu().numpy(), p32.cpu().numpy(), p98.cpu().numpy()


def test_phash_matches_reference_golden_vectors(golden_phash, golden_planes):
    groups = {}
    for case in golden_phash["cases"]:
        groups.setdefault((case["h"], case["w"], case["c"]), []).append(case)
    mismatched_bits = 0
    for (h, w, c), cases in groups.items():
        arrs = [synth.synth_image(cs["index"], h, w, c, n_set=1 << 30) for cs in cases]
        ph, dh, mg, p32, p98 = _hashes_of(arrs)
        for k, cs in enumerate(cases):
            assert f"{dh[k]:016x}" == cs["dhash"], cs
            key = f"p32_{cs['index']}"
            if key in golden_planes:
                assert np.array_equal(p32[k], golden_planes[key]), cs
                assert np.array_equal(p98[k], golden_planes[f"p98_{cs['index']}"]), cs
            if f"{ph[k]:016x}" != cs["phash"]:
                # only a near-tie of the reference's float32 DCT may differ
                assert mg[k] < 1e-3, (cs, mg[k])
                mismatched_bits += (int(ph[k]) ^ int(cs["phash"], 16)).bit_count()
    assert mismatched_bits == 0


@pytest.mark.parametrize("shape", [(512, 512, 3), (64, 64, 3), (33, 47, 3), (100, 33, 1), (48, 64, 4), (32, 32, 3),
                                   (8, 9, 1), (7, 5, 3), (31, 29, 3), (200, 1000, 3), (1000, 200, 1), (1, 1, 3),
                                   (16, 500, 4), (257, 255, 3), (2048, 1536, 3)])
def test_phash_planes_and_hashes_match_oracle(shape):
    """Against the CPU oracle: planes byte-identical to Pillow's arithmetic, dHash exact, pHash
    equal to the reference (cv2 float32 DCT) except at flagged near-ties."""
    from PIL import Image

    h, w, c = shape
    rng = np.random.default_rng(h * 7 + w)
    n = 3 if h * w > 1 << 20 else 6
    arrs = []
    for k in range(n):
        if k % 2:
            a = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        else:
            a = synth.synth_image(300 + k, h, w, c)
            a = a if a.ndim == 3 else a[..., None]
        arrs.append(a)
    ph, dh, mg, p32, p98 = _hashes_of(arrs)
    for k, a in enumerate(arrs):
        img = a[..., 0] if c == 1 else a
        oph, odh, omm, o32, o98 = oracle.signature(img)
        assert np.array_equal(p32[k], o32) and np.array_equal(p98[k], o98)
        assert dh[k] == odh
        pil = Image.fromarray(img, {1: "L", 3: "RGB", 4: "RGBA"}[c])
        ref = ref_py.phash(pil) & U64
        assert ref_py.dhash(pil) & U64 == dh[k]
        if ph[k] != ref:
            assert mg[k] < 1e-3, (shape, k, mg[k], hex(int(ph[k]) ^ ref))
        # and the double-precision restatement agrees unless it is a tie itself
        assert ph[k] == oph or omm < 1e-9


def test_phash_strided_and_gray_inputs():
    torch = _torch()
    base = torch.from_numpy(synth.synth_images(0, 4, 96, 160, 3)).cuda()
    ph0, dh0 = ops.phash_dhash_batch(base)
    # a view with padded rows / images (non-contiguous strides)
    padded = torch.zeros((4, 100, 170, 3), dtype=torch.uint8, device="cuda")
    padded[:, :96, :160] = base
    ph1, dh1 = ops.phash_dhash_batch(padded[:, :96, :160])
    assert torch.equal(ph0, ph1) and torch.equal(dh0, dh1)
    gray = torch.from_numpy(synth.synth_images(0, 4, 96, 160, 1)).cuda()
    ph2, _ = ops.phash_dhash_batch(gray)
    ph3, _ = ops.phash_dhash_batch(gray.unsqueeze(-1))
    assert torch.equal(ph2, ph3)
    assert ops.phash_dhash_batch(torch.empty((0, 8, 8, 3), dtype=torch.uint8, device="cuda"))[0].numel() == 0


def test_phash_big_batch_is_permutation_consistent():
    """Size-independent property at bench scale: hashing is per-image, so a shuffled batch gives
    shuffled hashes, and the GPU generator equals the NumPy generator byte for byte."""
    torch = _torch()
    n = 2048
    imgs = ops.synth_images_device(0, n, 512, 512, 3, n_set=n)
    for k in (0, 1, n - 1, n - 50):
        assert np.array_equal(imgs[k].cpu().numpy(), synth.synth_image(k, 512, 512, 3, n_set=n))
    ph, dh = ops.phash_dhash_batch(imgs)
    perm = torch.randperm(n, device="cuda")
    ph2, dh2 = ops.phash_dhash_batch(imgs[perm])
    assert torch.equal(ph[perm], ph2) and torch.equal(dh[perm], dh2)
    # spot-check against the reference path
    from PIL import Image

    for k in (3, 777, n - 1):
        pil = Image.fromarray(imgs[k].cpu().numpy(), "RGB")
        assert ref_py.phash(pil) == int(ph[k]) and ref_py.dhash(pil) == int(dh[k])
    # planted near-duplicates land within the scan threshold of their source
    close = 0
    for k in range(n - (n * 50) // 1000, n):
        src, _ = synth.image_source(k, n)
        close += ref_py.hamming64(int(ph[k]), int(ph[src])) <= 8
    assert close >= 0.9 * ((n * 50) // 1000)


# --------------------------------------------------------------------------------- K3 SSIM


def _pair(kind, h, w, rng, k):
    a = synth.synth_image(8000 + k, h, w, 1)
    if kind == 0:
        b = np.clip(a.astype(int) + rng.integers(-4, 5, a.shape), 0, 255).astype(np.uint8)
    elif kind == 1:
        b = synth.synth_image(8500 + k, h, w, 1)
    elif kind == 2:
        a = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        b = np.clip(a.astype(int) + int(rng.integers(-3, 4)), 0, 255).astype(np.uint8)
    elif kind == 3:
        b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    else:
        b = a.copy()
    return a, b


@pytest.mark.parametrize("shape", [(7, 7), (8, 31), (64, 64), (100, 37), (256, 256), (37, 300), (512, 512), (29, 263),
                                   (300, 1030)])
def test_ssim_matches_oracle(shape):
    torch = _torch()
    h, w = shape
    rng = np.random.default_rng(h + w)
    pairs = [_pair(k % 5, h, w, rng, k) for k in range(10)]
    bank = np.stack([p[0] for p in pairs] + [p[1] for p in pairs])
    ia = np.arange(10)
    ib = np.arange(10) + 10
    got = ops.ssim_batch(torch.from_numpy(bank).cuda(), ia, ib).cpu().numpy()
    got_host = ops.ssim_pairs(bank[:10], bank[10:])
    for k, (a, b) in enumerate(pairs):
        want = ref_py.ssim_of_planes(a, b)  # restated reference (float32 maps)
        exact = oracle.ssim_u8(a, b, exact=True)
        assert abs(got[k] - want) <= SSIM_TOL, (shape, k, got[k], want)
        assert abs(got[k] - exact) <= 2e-6, (shape, k, got[k], exact)
        assert abs(got_host[k] - got[k]) <= 1e-12
        for thr in (0.9, 0.92):  # identical accept/reject unless the reference itself is borderline
            if abs(want - thr) > SSIM_TOL:
                assert (got[k] >= thr) == (want >= thr)
    assert abs(got[4 if len(pairs) > 4 else 0] - 1.0) < 1e-6  # identical images


def test_ssim_rgb_bank_applies_pillow_luma():
    torch = _torch()
    imgs = synth.synth_images(0, 6, 80, 120, 3, n_set=6, planted=0.5)
    got = ops.ssim_batch(torch.from_numpy(imgs).cuda(), [0, 1, 2, 0], [3, 4, 5, 1]).cpu().numpy()
    for k, (i, j) in enumerate(((0, 3), (1, 4), (2, 5), (0, 1))):
        want = ref_py.ssim_of_planes(oracle.to_l(imgs[i]), oracle.to_l(imgs[j]))
        assert abs(got[k] - want) <= SSIM_TOL


def test_ssim_reference_behavioural_pins_and_errors():
    from PIL import Image, ImageEnhance

    base = Image.new("RGB", (64, 64), color=(200, 10, 10))
    var = ImageEnhance.Brightness(base).enhance(1.02)
    pa, pb = ref_py.ssim_planes(base, var)
    s = float(ops.ssim_pairs(pa, pb)[0])
    assert s > 0.95 and abs(s - ref_py.compute_ssim(base, var)) <= SSIM_TOL
    with pytest.raises(ValueError):
        ops.ssim_pairs(np.zeros((6, 20), np.uint8), np.zeros((6, 20), np.uint8))
    with pytest.raises(ValueError):
        ops.ssim_pairs(np.zeros((8, 20), np.uint8), np.zeros((8, 21), np.uint8))


def test_ssim_scale_properties_at_bench_size():
    """Config C4 shape (256x256 'L' crops): symmetry, identity, and an oracle sample."""
    torch = _torch()
    m = 4096
    bank = ops.synth_images_device(0, m, 256, 256, 1, n_set=m, planted=0.5)
    rng = np.random.default_rng(1)
    ia = rng.integers(0, m, 20000)
    ib = np.where(rng.random(20000) < 0.5, rng.integers(0, m, 20000), ia)
    s_ab = ops.ssim_batch(bank, ia, ib)
    s_ba = ops.ssim_batch(bank, ib, ia)
    assert torch.allclose(s_ab, s_ba, atol=1e-12, rtol=0)
    same = torch.from_numpy(ia == ib).cuda()
    assert torch.all((s_ab[same] - 1.0).abs() < 1e-6)
    host = bank.cpu().numpy()
    for k in range(0, 20000, 997):
        want = ref_py.ssim_of_planes(host[ia[k]], host[ib[k]])
        assert abs(float(s_ab[k]) - want) <= SSIM_TOL


def test_ssim_both_kernels_agree_with_the_oracle():
    """K3 has two kernels (v1: one output column per thread; v2: four, packed FP32); either may serve any shape."""
    from kobato_b200 import _native as nat

    torch = _torch()
    for (h, w, c) in ((256, 256, 1), (64, 80, 1), (7, 7, 1), (300, 520, 1), (96, 128, 3), (512, 512, 3), (40, 271, 4), (33, 1030, 1)):
        n = 12
        imgs = synth.synth_images(0, n, h, w, c, n_set=n, planted=0.5)
        bank = torch.from_numpy(imgs).cuda()
        ia, ib = list(range(0, n, 2)) + [0], list(range(1, n, 2)) + [0]
        got = {}
        ctx = nat.context(torch.cuda.current_device())
        for kernel in ("v1", "v2"):
            ctx.set_option(nat.KE_OPT_SSIM_V1, 1 if kernel == "v1" else 0)
            try:
                got[kernel] = ops.ssim_batch(bank, ia, ib).cpu().numpy()
            finally:
                ctx.set_option(nat.KE_OPT_SSIM_V1, 0)
        assert np.abs(got["v1"] - got["v2"]).max() <= 2e-6, (h, w, c)
        for k, (i, j) in enumerate(zip(ia, ib)):
            want = ref_py.ssim_of_planes(oracle.to_l(imgs[i]), oracle.to_l(imgs[j]))
            assert abs(got["v2"][k] - want) <= SSIM_TOL and abs(got["v1"][k] - want) <= SSIM_TOL, (h, w, c, k)
        assert abs(got["v2"][-1] - 1.0) < 1e-6


def test_phash_streaming_and_generic_kernels_agree():
    """K1 has two kernels (v5: the streaming tensor-pipe kernel, every batch of contiguous rows; generic: strided rows
    and the reference inside the library): identical planes, hashes and margins on every geometry — widths that are not a
    multiple of 16 or 4, rows / images off the 16-byte grid, bands whose fragments live in registers (two CTAs per SM
    up to ~512 px, one up to ~2200 px) and behind pointers (3000, 4096 px), upscales, one-row and one-column-block images."""
    torch = _torch()
    from kobato_b200 import _native as nat

    ctx = nat.context(torch.cuda.current_device())
    for (h, w, c, n) in ((512, 512, 3, 300), (96, 160, 3, 40), (200, 64, 4, 20), (130, 256, 1, 20), (1100, 1024, 3, 6),
                         (70, 100, 3, 9), (512, 512, 1, 64), (512, 512, 4, 64), (300, 256, 3, 40), (31, 48, 3, 17), (640, 480, 3, 20), (8, 16, 3, 5), (17, 512, 3, 300), (33, 512, 1, 40), (512, 16, 4, 33), (100, 496, 3, 7),
                         (1, 32, 3, 3), (47, 512, 4, 1), (2048, 32, 3, 3),
                         (33, 47, 3, 11), (100, 33, 1, 7), (61, 501, 3, 9), (45, 250, 4, 5), (64, 7, 3, 4), (9, 9, 1, 3),
                         (300, 1536, 3, 5), (256, 2048, 3, 4), (130, 1000, 1, 6), (96, 4096, 1, 2), (77, 1201, 3, 3),
                         (64, 3000, 4, 2), (1536, 2048, 3, 2)):
        imgs = ops.synth_images_device(0, n, h, w, c, n_set=n)
        got = ops.phash_dhash_batch(imgs, want_margin=True, want_planes=True)
        if w > 2048:  # beyond the generic kernel's shared-memory reach: the CPU oracle is the reference
            host = imgs.cpu().numpy()
            for k in range(n):
                wp, wd, wm, p32, p98 = oracle.signature(host[k])
                assert np.array_equal(got[3][0][k].cpu().numpy(), p32) and np.array_equal(got[3][1][k].cpu().numpy(), p98), (h, w, c)
                assert int(got[0][k].item()) & U64 == wp and int(got[1][k].item()) & U64 == wd, (h, w, c)
            continue
        ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 1)
        try:
            gen = ops.phash_dhash_batch(imgs, want_margin=True, want_planes=True)
        finally:
            ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 0)
        assert torch.equal(got[3][0], gen[3][0]) and torch.equal(got[3][1], gen[3][1]), (h, w, c)
        assert torch.equal(got[0], gen[0]) and torch.equal(got[1], gen[1]) and torch.equal(got[2], gen[2]), (h, w, c)
        if n > 2:  # a batch that starts off the 16-byte grid (a slice of a larger tensor)
            flat = torch.empty(imgs.numel() + 64, dtype=torch.uint8, device=imgs.device)
            for shift in (1, 4, 13):
                view = flat[shift:shift + imgs.numel()].view(imgs.shape)
                view.copy_(imgs)
                off = ops.phash_dhash_batch(view, want_planes=True)
                assert torch.equal(off[0], gen[0]) and torch.equal(off[1], gen[1]), (h, w, c, shift)
                assert torch.equal(off[2][0], gen[3][0]) and torch.equal(off[2][1], gen[3][1]), (h, w, c, shift)


def test_phash_widths_around_the_kernel_class_boundaries():
    """K1 picks its kernel by the width: resample bands of <= 8 k-steps (two CTAs per SM), <= 16 and <= 32 (one CTA per SM,
    wide-target fragments in registers, narrow-target warps split by K with <= 9 / 18 k-steps each), beyond (pointer-fed).
    Widths on both sides of every boundary — aligned and not, heights that end in the middle of a ring buffer — must give
    the generic kernel's (w <= 2048) or the CPU oracle's planes and hashes."""
    torch = _torch()
    from kobato_b200 import _native as nat

    ctx = nat.context(torch.cuda.current_device())
    widths = (513, 520, 528, 544, 576, 600, 1039, 1040, 1056, 1088, 1104, 1119, 1120, 1136, 1152, 1153, 1168, 1200,
              2047, 2049, 2175, 2176, 2208, 2239, 2240, 2272, 2303, 2304, 2305, 2336, 2400)
    for i, w in enumerate(widths):
        h, c, n = (41, 70, 97, 130)[i % 4], (3, 1, 4)[i % 3], 3
        imgs = ops.synth_images_device(0, n, h, w, c, n_set=n, seed=100 + i)
        got = ops.phash_dhash_batch(imgs, want_planes=True)
        if w > 2048:
            host = imgs.cpu().numpy()
            for k in range(n):
                wp, wd, _, p32, p98 = oracle.signature(host[k])
                assert np.array_equal(got[2][0][k].cpu().numpy(), p32) and np.array_equal(got[2][1][k].cpu().numpy(), p98), (h, w, c)
                assert int(got[0][k].item()) & U64 == wp and int(got[1][k].item()) & U64 == wd, (h, w, c)
            continue
        ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 1)
        try:
            gen = ops.phash_dhash_batch(imgs, want_planes=True)
        finally:
            ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 0)
        assert torch.equal(got[2][0], gen[2][0]) and torch.equal(got[2][1], gen[2][1]), (h, w, c)
        assert torch.equal(got[0], gen[0]) and torch.equal(got[1], gen[1]), (h, w, c)


def test_phash_every_staging_configuration_gives_the_same_hashes():
    """The streaming K1 kernel under pinned staging configurations (KE_OPT_PHASH_CFG: sub-chunk rows, raw slots, luma
    ring buffers of 32 or 16 rows, resample fragments in registers / shared memory / L2): every configuration that fits a geometry must give
    the automatic choice's planes and hashes, on short and long rows, aligned or not."""
    torch = _torch()
    from kobato_b200 import _native as nat

    ctx = nat.context(torch.cuda.current_device())
    for (h, w, c, n) in ((200, 512, 3, 40), (150, 1024, 3, 12), (97, 2048, 3, 6), (130, 1000, 1, 9), (77, 1201, 3, 5), (64, 640, 4, 7)):
        imgs = ops.synth_images_device(0, n, h, w, c, n_set=n)
        want = ops.phash_dhash_batch(imgs, want_planes=True)
        tried = 0
        for cr16 in (0, 1):
            for place in (0, 1, 2, 3):
                for bufs in (1, 2, 3):
                    for sub in (16, 4, 1):
                        value = sub | 1 << 8 | bufs << 12 | place << 16 | cr16 << 20
                        ctx.set_option(nat.KE_OPT_PHASH_CFG, value)
                        try:
                            got = ops.phash_dhash_batch(imgs, want_planes=True)
                        except nat.KobatoNativeError:
                            continue  # this configuration does not fit the geometry
                        finally:
                            ctx.set_option(nat.KE_OPT_PHASH_CFG, 0)
                        tried += 1
                        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), (h, w, c, hex(value))
                        assert torch.equal(got[2][0], want[2][0]) and torch.equal(got[2][1], want[2][1]), (h, w, c, hex(value))
        assert tried >= 12, (h, w, c, tried)


@pytest.mark.parametrize("mode", [2, 3])
def test_join_hybrid_and_bit_sliced_kernels_match_oracle(mode):
    """K2 has a second, LOP3-only (bit-sliced carry-save) kernel that runs concurrently with the POPC
    kernel on large tables (mode 2) or alone (mode 3); both must give the oracle's pair set."""
    torch = _torch()
    from kobato_b200 import _native as nat

    ctx = nat.context(torch.cuda.current_device())
    ctx.set_option(nat.KE_OPT_JOIN_MODE, mode)
    try:
        h = synth.synth_hashes(20011, seed=17, planted=0.25, max_flips=16)
        for threshold, band in ((0, False), (5, True), (8, False), (15, True), (16, False)):
            want = oracle.hamming_join(h, threshold, require_band=band, threads=8)
            got = ops.hamming_join(torch.from_numpy(h.view(np.int64)).cuda(), threshold, require_band=band)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), (mode, threshold, band)
        for n in (2, 31, 32, 33, 2047, 2049, 4097):
            hh = synth.synth_hashes(n, seed=3, planted=0.3)
            hh[-1] = hh[0]
            for g, w in zip(ops.hamming_join(hh, 6), oracle.hamming_join(hh, 6)):
                assert np.array_equal(g, w), (mode, n)
        pieces = [ops.hamming_join(h, 8, part_index=p, part_count=3) for p in range(3)]
        full = oracle.hamming_join(h, 8, threads=8)
        i = np.concatenate([p[0] for p in pieces]); j = np.concatenate([p[1] for p in pieces])
        order = np.lexsort((j, i))
        assert np.array_equal(i[order], full[0]) and np.array_equal(j[order], full[1])
    finally:
        ctx.set_option(nat.KE_OPT_JOIN_MODE, 0)


def test_join_config_c5_ten_million_hashes():
    """Config C5 (10 M hashes, 5e13 pairs, T=8), run as 8 tile partitions on this GPU exactly as the
    8-GPU split does: the union must contain every planted pair, nothing out of range, and equal the
    CPU oracle on sampled row stripes."""
    torch = _torch()
    n = 10_000_000
    h = synth.synth_hashes(n, planted=0.05)
    table = torch.from_numpy(h.view(np.int64)).cuda()
    parts = [ops.hamming_join_device(table, 8, part_index=p, part_count=8, capacity=1 << 22) for p in range(8)]
    i = torch.cat([p[0] for p in parts]).to(torch.int64) & 0xFFFFFFFF
    j = torch.cat([p[1] for p in parts]).to(torch.int64) & 0xFFFFFFFF
    d = torch.cat([p[2] for p in parts])
    order = torch.argsort((i << 32) | j)
    i, j, d = i[order].cpu().numpy(), j[order].cpu().numpy(), d[order].cpu().numpy()
    assert len(np.unique((i << 32) | j)) == len(i), "a pair was emitted by two partitions"
    assert np.all(i < j) and j.max() < n and d.max() <= 8
    x = h[i] ^ h[j]
    assert np.array_equal(np.bitwise_count(x).astype(np.uint8), d)
    for lo in (0, 4_999_900, 9_999_700):
        wi, wj, wd = oracle.hamming_join(h, 8, row_begin=lo, row_end=lo + 200, threads=8)
        sel = (i >= lo) & (i < lo + 200)
        assert np.array_equal(i[sel], wi) and np.array_equal(j[sel], wj) and np.array_equal(d[sel], wd)
    n_base = n - (n * 50) // 1000
    t = np.arange(n_base, n, dtype=np.uint64)
    src = (synth._mix(synth.SEED, t, 1) % np.uint64(n_base)).astype(np.int64)
    dist = np.bitwise_count(h[src] ^ h[n_base:])
    want = set(zip(src[dist <= 8].tolist(), (np.flatnonzero(dist <= 8) + n_base).tolist()))
    got = set(zip(i.tolist(), j.tolist()))
    assert want <= got and len(want) > 300_000


def test_streaming_kernels_random_batches_match_the_generic_kernels():
    """K1 v5 and the N1 streaming resize on random batch sizes (ring / slot hand-off paths with n below, at and
    above the persistent grid) against the generic kernels: hashes, planes and resized planes identical."""
    import os

    torch = _torch()
    from kobato_b200 import _native as nat

    ctx = nat.context(torch.cuda.current_device())
    rng = np.random.default_rng(77)
    shapes = [(512, 512, 3), (512, 512, 1), (512, 512, 4), (96, 160, 3), (300, 256, 3), (47, 512, 4), (33, 512, 1)]
    for it in range(21):
        h, w, c = shapes[it % len(shapes)]
        n = int(rng.integers(1, 700))
        imgs = ops.synth_images_device(int(rng.integers(0, 1 << 20)), n, h, w, c, n_set=1 << 30)
        got = ops.phash_dhash_batch(imgs, want_planes=True)
        ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 1)
        try:
            ref = ops.phash_dhash_batch(imgs, want_planes=True)
        finally:
            ctx.set_option(nat.KE_OPT_PHASH_GENERIC, 0)
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), (it, h, w, c, n)
        assert torch.equal(got[2][0], ref[2][0]) and torch.equal(got[2][1], ref[2][1]), (it, h, w, c, n)
        side = (32, 64, 128)[it % 3]
        fast = ops.gray_resize_batch(imgs, side, side, "bilinear")
        ctx.set_option(nat.KE_OPT_RESIZE_GENERIC, 1)
        try:
            slow = ops.gray_resize_batch(imgs, side, side, "bilinear")
        finally:
            ctx.set_option(nat.KE_OPT_RESIZE_GENERIC, 0)
        assert torch.equal(fast, slow), (it, h, w, c, n, side)
