"""GPU parity tests of the round-2 entry points (run with -m gpu on the B200 box), all through the C ABI:
the independent SSIM pins on the kernel itself, the optional Gaussian window, the table-level scan with the device
union-find (N3), the multi-device context, the luma-plane gather, the staged host copies and — with two or more GPUs —
the multi-process pipeline with duplicates planted across shards."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from conftest import ROOT, normalise_clusters

import oracle
from kobato_b200 import _native as nat
from kobato_b200 import ops, synth
from kobato_b200.dup import scanner as kscanner
from oracle import ref_py
from oracle import ssim_independent as ind

pytestmark = pytest.mark.gpu
SSIM_TOL = 1e-5


def _torch():
    import torch

    return torch


# --------------------------------------------------------------------------------- K3 pinned independently


def test_ssim_kernel_agrees_with_the_independent_oracle_and_closed_forms():
    """K3 against things the builder's restatement did not generate: the float64 49-tap evaluation of
    the published definition (oracle/ssim_independent.py, no uniform_filter, no code shared with ref_py) on the C4 shape
    and the edge shapes, and the closed forms (constants, identical, inverted checkerboard).  Restatement, brute force
    and kernel must agree within 1e-5."""
    rng = np.random.default_rng(17)
    for (h, w) in ((256, 256), (7, 7), (8, 31), (100, 37), (29, 263), (64, 520)):
        a = synth.synth_image(5000 + h, h, w, 1)
        variants = [np.clip(a.astype(np.int64) * 261 // 256 + 1, 0, 255).astype(np.uint8),
                    synth.synth_image(6000 + w, h, w, 1),
                    np.clip(a.astype(np.int64) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8), a.copy()]
        got = ops.ssim_pairs(np.stack([a] * len(variants)), np.stack(variants))
        for k, b in enumerate(variants):
            brute = ind.mssim_bruteforce(a, b)
            restated = ref_py.ssim_of_planes(a, b)
            assert abs(got[k] - brute) <= SSIM_TOL, (h, w, k, got[k], brute)
            assert abs(restated - brute) <= SSIM_TOL and abs(got[k] - restated) <= SSIM_TOL
            for thr in (0.9, 0.92):
                if abs(brute - thr) > SSIM_TOL:
                    assert (got[k] >= thr) == (brute >= thr)
    for la, lb, shape in ((0, 0, (7, 7)), (255, 255, (9, 13)), (200, 10, (64, 64)), (17, 18, (31, 8)), (0, 255, (16, 16)),
                          (128, 128, (256, 256))):
        got = float(ops.ssim_pairs(np.full(shape, la, np.uint8), np.full(shape, lb, np.uint8))[0])
        assert abs(got - ind.closed_form_constants(la, lb)) <= 1e-6, (la, lb, got)
    for (h, w, p, q) in ((16, 16, 200, 50), (31, 8, 255, 0), (64, 40, 130, 120), (7, 7, 10, 240), (256, 256, 90, 91)):
        a = ind.checkerboard(h, w, p, q)
        b = (255 - a.astype(np.int32)).astype(np.uint8)
        got = float(ops.ssim_pairs(a, b)[0])
        assert abs(got - ind.closed_form_inverted_checkerboard(h, w, p, q)) <= SSIM_TOL, (h, w, p, q, got)


def test_ssim_wide_planes_are_bit_reproducible():
    """Planes wider than one column block used to be reduced with atomicAdd(double): now per-block partials are summed
    in a fixed order, so repeated runs (and the v1 / v2 kernels' own repeats) give identical bits."""
    torch = _torch()
    imgs = synth.synth_images(0, 24, 96, 1100, 1, n_set=24, planted=0.5)
    bank = torch.from_numpy(imgs).cuda()
    ia, ib = np.arange(0, 24, 2), np.arange(1, 24, 2)
    runs = [ops.ssim_batch(bank, ia, ib).cpu().numpy() for _ in range(6)]
    assert all(np.array_equal(runs[0], r) for r in runs[1:])
    for k in range(len(ia)):
        assert abs(runs[0][k] - ref_py.ssim_of_planes(imgs[ia[k]], imgs[ib[k]])) <= SSIM_TOL
    rgb = synth.synth_images(0, 8, 40, 600, 3, n_set=8, planted=0.5)
    bank3 = torch.from_numpy(rgb).cuda()
    r3 = [ops.ssim_batch(bank3, [0, 2, 4, 6], [1, 3, 5, 7]).cpu().numpy() for _ in range(4)]
    assert all(np.array_equal(r3[0], r) for r in r3[1:])


@pytest.mark.parametrize("shape", [(11, 11, 1), (64, 64, 1), (40, 33, 1), (256, 256, 1), (50, 70, 3), (37, 300, 4)])
def test_ssim_gaussian_window_matches_the_restated_skimage_variant(shape):
    """ke_ssim_batch(gaussian=1) = skimage's gaussian_weights=True (sigma 1.5, 11 taps, crop 5): NOT the reference's
    path, offered because north_star words kernel 3 that way.  Compared with the scipy.ndimage.gaussian_filter
    restatement at the same 1e-5 bar; the kernel mirrors scipy's FP64-accumulate / float32-store passes, so constant
    and identical inputs give exactly the float32 path's values."""
    torch = _torch()
    h, w, c = shape
    n = 10
    imgs = synth.synth_images(0, n, h, w, c, n_set=n, planted=0.5)
    bank = torch.from_numpy(imgs).cuda()
    ia, ib = [0, 1, 2, 3, 4, 0], [5, 6, 7, 8, 9, 0]
    got = ops.ssim_batch(bank, ia, ib, gaussian=True).cpu().numpy()
    planes = [oracle.to_l(x) if c > 1 else (x if x.ndim == 2 else x[..., 0]) for x in imgs]
    for k, (i, j) in enumerate(zip(ia, ib)):
        want = ref_py.ssim_gaussian_of_planes(planes[i], planes[j])
        assert abs(got[k] - want) <= SSIM_TOL, (shape, k, got[k], want)
    assert abs(got[-1] - 1.0) <= 1e-6
    if c == 1:
        host = ops.ssim_pairs(np.stack([planes[i] for i in ia]), np.stack([planes[j] for j in ib]), gaussian=True)
        assert np.array_equal(host, got)
    flat = ops.ssim_pairs(np.full((20, 30), 200, np.uint8), np.full((20, 30), 10, np.uint8), gaussian=True)[0]
    assert abs(flat - ref_py.ssim_gaussian_of_planes(np.full((20, 30), 200, np.uint8), np.full((20, 30), 10, np.uint8))) <= 1e-7
    with pytest.raises(ValueError):
        ops.ssim_pairs(np.zeros((10, 30), np.uint8), np.zeros((10, 30), np.uint8), gaussian=True)


# --------------------------------------------------------------------------------- N3 table scan


def _columns(n, seed, *, planted=0.1, sizes=True):
    h = synth.synth_hashes(n, seed=seed, planted=planted)
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(np.arange(1, 10 * n), n, replace=False)).astype(np.int64)
    sz = rng.integers(0, 5_000_000, n).astype(np.int64) if sizes else None
    if sz is not None:
        sz[rng.random(n) < 0.05] = 0  # COALESCE(size, 0): unknown sizes pass the gate
    return h, ids, sz


@pytest.mark.parametrize("cfg", [dict(threshold=8), dict(threshold=4, size_ratio=0.5), dict(threshold=12, size_ratio=0.9),
                                 dict(threshold=8, band_bits=8, band_count=8), dict(threshold=10, band_bits=12, band_count=5),
                                 dict(threshold=8, pair_cap=3), dict(threshold=0), dict(threshold=6, band_bits=32, band_count=2)])
def test_scan_table_matches_the_restated_scanner(cfg):
    """ke_scan_table_host against the oracle's restatement of the reference scan (scan_edges + DSU): the same members,
    component labels, best_hamming and surviving edges, bit for bit."""
    h, ids, sz = _columns(3000, seed=31 + cfg["threshold"], planted=0.25)
    got = ops.scan_table(h.view(np.int64), ids, sz, want_edges=True, **cfg)
    want = ref_py.scan_table(h, ids, sz, **cfg)
    for key in ("index", "label", "best", "offsets"):
        assert np.array_equal(got[key], want[key]), (cfg, key)
    for g, w in zip(got["edges"], want["edges"]):
        assert np.array_equal(g, w)
    for key in ("edges", "members", "clusters"):
        assert got["stats"][key] == want["stats"][key]
    if cfg["threshold"] >= 4 and "pair_cap" not in cfg:
        assert got["stats"]["clusters"] > 50


def test_scan_table_edge_cases():
    assert ops.scan_table(np.zeros(0, np.int64))["stats"]["members"] == 0
    assert ops.scan_table(np.array([5], np.int64))["stats"]["members"] == 0
    two = ops.scan_table(np.array([5, 5], np.int64), np.array([1, 2]), None)
    assert two["index"].tolist() == [0, 1] and two["label"].tolist() == [0, 0] and two["best"].tolist() == [0, 0]
    same_id = ops.scan_table(np.array([5, 5], np.int64), np.array([7, 7]), None)  # one file listed twice: no edge
    assert same_id["stats"]["members"] == 0 and same_id["stats"]["candidates"] == 1
    # all-distinct bands: the reference stops at "no bucket has 2+ items"
    h = (np.arange(1, 200, dtype=np.uint64) * np.uint64(0x0001000100010001)).view(np.int64)
    st = ops.scan_table(h)["stats"]
    assert st["buckets_ge2"] == 0 and st["members"] == 0
    with pytest.raises(ValueError):
        ops.scan_table(np.zeros(4, np.int64), threshold=65)
    with pytest.raises(ValueError):
        ops.scan_table(np.zeros(4, np.int64), band_bits=16, band_count=5)
    # one giant component: 5000 copies of one hash -> a single cluster whatever the scheduling of the unions
    big = ops.scan_table(np.full(5000, 0x1234, np.int64), threshold=0)
    assert big["stats"]["clusters"] == 1 and big["stats"]["members"] == 5000 and np.all(big["label"] == 0)
    assert big["stats"]["candidates"] == 5000 * 4999 // 2


def test_scanner_table_path_on_gpu_matches_golden_and_scales(golden_scanner, monkeypatch):
    """DuplicateScanner's default path now IS the table path: the live reference's golden clusters, then 1 M rows
    through build_clusters_from_columns (no per-row objects) checked against the legacy host path on the same join."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for case in golden_scanner["cases"]:
        if "pair_cap" in case:
            monkeypatch.setenv("KE_DUP_BUCKET_PAIR_CAP", str(case["pair_cap"]))
        else:
            monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)
        recs = mod.make_files(case["n"], case["seed"], case.get("ids_dupe", False))
        files = [kscanner.DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                                        height=f["height"], phash=f["phash"]) for f in recs]
        clusters = kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(**case["cfg"])).build_clusters(files)
        got = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in clusters]
        assert normalise_clusters(got) == normalise_clusters(case["clusters"]), case["name"]
    monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)

    n = 1_000_000
    h, ids, sz = _columns(n, seed=5, planted=0.05)
    made = []

    def make_file(row):
        made.append(row)
        return kscanner.DuplicateFile(file_id=int(ids[row]), path=Path(f"d/f{row}.jpg"), size=int(sz[row]), width=64,
                                      height=64, phash=int(h[row]))

    cfg = kscanner.DuplicateScanConfig(hamming_threshold=8, size_ratio=0.5)
    clusters = kscanner.DuplicateScanner(cfg).build_clusters_from_columns(ids, h.view(np.int64), sz, make_file=make_file)
    assert len(clusters) > 10_000 and len(made) < n // 5  # objects exist for members only
    # the legacy host path over the same candidates (join -> Python filters -> dict DSU) must give the same clusters
    ii, jj, dd = ops.hamming_join(h, 8, require_band=True)
    member_rows = sorted(set(made))
    sub = [make_file(r) for r in member_rows]
    pos = {r: k for k, r in enumerate(member_rows)}
    keep = np.isin(ii, member_rows) & np.isin(jj, member_rows)
    sub_join = (np.array([pos[r] for r in ii[keep]], np.uint32), np.array([pos[r] for r in jj[keep]], np.uint32), dd[keep])
    legacy = kscanner.DuplicateScanner(cfg, join=lambda hashes, c, allow: sub_join)
    want = legacy.build_clusters(sub)
    norm = lambda cl: sorted((c.keeper_id, tuple((e.file.file_id, e.best_hamming) for e in c.files)) for c in cl)  # noqa: E731
    assert norm(clusters) == norm(want)


def test_cluster_pairs_device_matches_the_host_union_find():
    torch = _torch()
    rng = np.random.default_rng(2)
    n = 50_000
    a = rng.integers(0, n, 30_000)
    b = rng.integers(0, n, 30_000)
    label = ops.cluster_pairs_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), n).cpu().numpy()
    members, offsets = ops.cluster_pairs_csr(a, b)
    want = np.full(n, -1, np.int64)
    for lo, hi in zip(offsets[:-1], offsets[1:]):
        want[members[lo:hi]] = members[lo]
    assert np.array_equal(label, want)


# --------------------------------------------------------------------------------- N2 ORB matcher


def test_orb_matcher_equals_opencv_cross_check():
    """ke_orb_match_pairs against the LIVE cv2.BFMatcher(NORM_HAMMING, crossCheck=True) — match lists, distances and
    counts — on tie-heavy random descriptor sets, on real ORB descriptors of synthetic images, and through the drop-in
    dup.refine._compute_orb_ratio / refine_pairs_batch (the reference's ratio, src/dup/refine.py:55-68)."""
    import cv2
    from PIL import Image

    from kobato_b200.dup import refine as krefine

    rng = np.random.default_rng(1)
    A, B = [], []
    for trial in range(200):
        na, nb = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        bits = int(rng.integers(1, 9))
        da = rng.integers(0, 1 << bits, (na, 32)).astype(np.uint8)
        db = rng.integers(0, 1 << bits, (nb, 32)).astype(np.uint8)
        if trial % 3 == 0:
            db[: min(na, nb)] = da[: min(na, nb)]
        A.append(da)
        B.append(db)
    A += [None, np.zeros((0, 32), np.uint8)]
    B += [B[0], B[1]]
    counts, matches = ops.orb_match_pairs(A, B, want_matches=True)
    for p in range(200):
        want = sorted((m.queryIdx, m.trainIdx, int(m.distance)) for m in
                      cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(A[p], B[p]))
        assert matches[p] == want and counts[p] == len(want), p
    assert counts[200] == 0 and counts[201] == 0
    assert np.array_equal(ops.orb_match_pairs(A, B), counts)

    def live_ratio(ia, ib):  # the reference's function body, on the host
        orb = cv2.ORB_create()
        kpa, da = orb.detectAndCompute(np.asarray(ia.convert("L")), None)
        kpb, db = orb.detectAndCompute(np.asarray(ib.convert("L")), None)
        if da is None or db is None or not kpa or not kpb:
            return 0.0
        m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db)
        return float(len(m) / min(len(kpa), len(kpb))) if m else 0.0

    imgs = [Image.fromarray(synth.synth_image(40 + k, 240, 320, 3)) for k in range(4)]
    imgs.append(Image.fromarray(np.roll(np.asarray(imgs[0]), 5, axis=1)))
    imgs.append(Image.new("RGB", (64, 64), (10, 200, 10)))
    some = 0
    for a, b in ((0, 4), (0, 1), (2, 3), (1, 5), (4, 0)):
        got, want = krefine._compute_orb_ratio(imgs[a], imgs[b]), live_ratio(imgs[a], imgs[b])
        assert got == want, (a, b, got, want)
        some += want > 0
    assert some >= 2


# --------------------------------------------------------------------------------- plumbing


def test_luma_planes_are_pillow_luma():
    torch = _torch()
    for (h, w, c) in ((64, 64, 3), (33, 47, 3), (20, 24, 4), (16, 16, 1), (512, 512, 3)):
        imgs = synth.synth_images(0, 6, h, w, c, n_set=6)
        bank = torch.from_numpy(imgs).cuda()
        idx = [5, 0, 3, 3]
        got = ops.luma_planes(bank, idx).cpu().numpy()
        for k, i in enumerate(idx):
            want = oracle.to_l(imgs[i]) if c > 1 else (imgs[i] if imgs[i].ndim == 2 else imgs[i][..., 0])
            assert np.array_equal(got[k], want), (h, w, c, k)
    assert ops.luma_planes(bank, []).shape == (0, 512, 512)


def test_host_entry_points_take_pageable_and_pinned_sources():
    """ke_phash_batch_host / ke_ssim_pairs_host / ke_hamming_join_host: pageable numpy memory goes through the pinned
    staging buffers (several 32 MB pieces here), page-locked memory by DMA in place; identical results either way and
    equal to the device-resident path."""
    torch = _torch()
    n, h, w = 700, 256, 256  # 137 MB: five staging pieces, two chunk buffers... and an odd image size for the 2-D path
    bank = ops.synth_images_device(0, n, h, w, 3, n_set=n)
    host = bank.cpu()
    ph_dev, dh_dev = ops.phash_dhash_batch(bank)
    ph_page, dh_page = ops.phash_dhash_batch(host.numpy())
    pinned = host.pin_memory()
    ph_pin, dh_pin = ops.phash_dhash_batch(pinned.numpy())
    for ph, dh in ((ph_page, dh_page), (ph_pin, dh_pin)):
        assert np.array_equal(ph, ph_dev.cpu().numpy()) and np.array_equal(dh, dh_dev.cpu().numpy())
    odd = synth.synth_images(0, 40, 33, 47, 3, n_set=40)  # 4653 B per image: padded to 16 B on the device
    a, b = ops.phash_dhash_batch(odd)
    c, d = ops.phash_dhash_batch(torch.from_numpy(odd).cuda())
    assert np.array_equal(a, c.cpu().numpy()) and np.array_equal(b, d.cpu().numpy())


def test_multi_device_context_splits_the_host_entry_points():
    """ke_ctx_create_multi over every visible device (one device: the degenerate fan of one): results of the fanned host
    entry points equal the single-device ones; with >= 2 GPUs each of them really launches kernels."""
    torch = _torch()
    ndev = torch.cuda.device_count()
    ctx = nat.Context(list(range(ndev)))
    try:
        assert ctx.n_devices == ndev and ctx.devices == list(range(ndev))
        lib = nat.load()
        import ctypes as C

        n = 300_000  # 4.5e10 pairs: enough for the fan to use every device (one per 5e9 pairs)
        h = synth.synth_hashes(n, seed=3, planted=0.05)
        cap = 1 << 20
        ii, jj, dd = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.uint8)
        total = C.c_int64(0)
        before = [nat.Context(_borrowed=lib.ke_ctx_child(ctx.handle, k)).launches for k in range(ndev)]
        nat.check(lib.ke_hamming_join_host(ctx.handle, h.ctypes.data, n, 8, 1, 16, 4, None, 0, 1, ii.ctypes.data,
                                           jj.ctypes.data, dd.ctypes.data, cap, C.byref(total)), "join")
        t = total.value
        order = np.lexsort((jj[:t], ii[:t]))
        wi, wj, wd = oracle.hamming_join(h, 8, require_band=True, threads=8)
        assert np.array_equal(ii[:t][order], wi) and np.array_equal(jj[:t][order], wj) and np.array_equal(dd[:t][order], wd)
        if ndev > 1:
            for k in range(1, ndev):
                child = nat.Context(_borrowed=lib.ke_ctx_child(ctx.handle, k))
                assert child.device == k and child.launches > before[k], f"device {k} took no tiles"
        imgs = synth.synth_images(0, 64, 96, 160, 3, n_set=64)
        big = np.concatenate([imgs] * 24)  # 70 MB: above the per-device minimum, so two devices share it when present
        ph, dh = np.empty(len(big), np.int64), np.empty(len(big), np.int64)
        nat.check(lib.ke_phash_batch_host(ctx.handle, big.ctypes.data, len(big), 96, 160, 3, ph.ctypes.data, dh.ctypes.data,
                                          None), "phash")
        sp, sd = ops.phash_dhash_batch(torch.from_numpy(imgs).cuda())
        assert np.array_equal(ph, np.tile(sp.cpu().numpy(), 24)) and np.array_equal(dh, np.tile(sd.cpu().numpy(), 24))
    finally:
        ctx.close()


def _spawn_multigpu(world: int, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(ROOT / "tools" / "check_multigpu.py"), *extra]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    return subprocess.run(cmd, capture_output=True, text=True, timeout=840, env=env)


@pytest.mark.timeout(900)
def test_pipeline_scan_world_size_2_matches_the_oracle_with_cross_shard_duplicates():
    """pipeline.scan as two processes on two GPUs (NCCL), duplicates planted over the GLOBAL index space and UNEQUAL
    shards: hashes, candidates, every SSIM score, the accept/reject decisions and the clusters equal the oracle's on the
    whole set (tools/check_multigpu.py does the comparison on rank 0 and exits non-zero on any difference)."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = _spawn_multigpu(2, ["--images", "260", "--uneven"])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "multigpu check ok" in res.stdout


def test_pipeline_scan_clusters_on_the_device_equal_the_host_union_find():
    """pipeline.scan keeps the candidate list on the GPU; long lists are unioned there (device union-find + sort),
    short ones by the library's host union-find: the two must give the same components in the same order."""
    import torch

    from kobato_b200 import pipeline

    n = 600
    bank = ops.synth_images_device(0, n, 64, 64, 3, n_set=n, planted=0.5)
    a = pipeline.scan(bank, cluster_on_device=False)
    b = pipeline.scan(bank, cluster_on_device=True)
    assert len(a.cand_i) > 50 and a.counts["accepted"] > 10
    assert np.array_equal(a.cand_i, b.cand_i) and np.array_equal(a.cand_j, b.cand_j) and np.array_equal(a.ssim, b.ssim)
    assert a.clusters.as_list() == b.clusters.as_list() and len(a.clusters) > 3
    assert np.array_equal(a.clusters.members, b.clusters.members) and np.array_equal(a.clusters.offsets, b.clusters.offsets)
