"""Pin the CPU oracle (oracle/) before trusting it: against the golden vectors produced by the
live reference (tests/golden/make_golden.py) and, when /root/reference is mounted, against the
reference itself on fresh seeded inputs.  CPU only."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import pytest
import sys

from conftest import REFERENCE_SRC, ROOT, normalise_clusters

import oracle
from kobato_b200 import synth
from oracle import ref_py

U64 = (1 << 64) - 1
MODES = {1: "L", 3: "RGB", 4: "RGBA"}


def _pil(arr, c):
    from PIL import Image

    return Image.fromarray(arr, MODES[c])


# ------------------------------------------------------------------ pHash / dHash


def test_python_restatement_matches_reference_golden_hashes(golden_phash):
    for case in golden_phash["cases"]:
        arr = synth.synth_image(case["index"], case["h"], case["w"], case["c"], n_set=1 << 30)
        im = _pil(arr, case["c"])
        assert f"{ref_py.phash(im) & U64:016x}" == case["phash"], case
        assert f"{ref_py.dhash(im) & U64:016x}" == case["dhash"], case


def test_c_restatement_planes_are_byte_identical_to_pillow(golden_phash, golden_planes):
    checked = 0
    for case in golden_phash["cases"]:
        key = f"p32_{case['index']}"
        if key not in golden_planes:
            continue
        arr = synth.synth_image(case["index"], case["h"], case["w"], case["c"], n_set=1 << 30)
        _, _, _, p32, p98 = oracle.signature(arr)
        assert np.array_equal(p32, golden_planes[key]), case
        assert np.array_equal(p98, golden_planes[f"p98_{case['index']}"]), case
        checked += 1
    assert checked >= 20


def test_c_restatement_hashes_match_reference_golden(golden_phash):
    """dHash must be exact; the f64-DCT pHash may differ from cv2's float32 DCT only at a near-tie."""
    near_ties = 0
    for case in golden_phash["cases"]:
        arr = synth.synth_image(case["index"], case["h"], case["w"], case["c"], n_set=1 << 30)
        ph, dh, margin, _, _ = oracle.signature(arr)
        assert f"{dh:016x}" == case["dhash"], case
        if f"{ph:016x}" != case["phash"]:
            assert margin < 1e-3, (case, margin)
            near_ties += 1
    assert near_ties == 0  # none on the committed vectors


def test_signed_wrap():
    assert ref_py.to_signed64(0) == 0
    assert ref_py.to_signed64(U64) == -1
    assert ref_py.to_signed64(1 << 63) == -(1 << 63)
    assert ref_py.to_signed64((1 << 63) - 1) == (1 << 63) - 1


def test_resample_table_shape_and_normalisation():
    kk, bd = oracle.resample_table(512, 32)
    assert kk.shape == (32, 97) and bd.shape == (32, 2)
    # 22-bit fixed point: every row sums to ~2^22
    assert np.all(np.abs(kk.sum(axis=1) - (1 << 22)) < 64)
    kk9, bd9 = oracle.resample_table(512, 9)
    assert kk9.shape[1] == 343
    assert bd9[:, 0].min() == 0 and (bd9[:, 0] + bd9[:, 1]).max() == 512


@pytest.mark.reference
def test_live_reference_hashes_on_fresh_images(reference_modules):
    """Fresh seeds (not the committed ones), several sizes/modes, against the live reference."""
    ph_mod = reference_modules["phash"]
    rng = np.random.default_rng(77)
    for k in range(24):
        h, w = int(rng.integers(7, 700)), int(rng.integers(7, 700))
        c = (1, 3, 4)[k % 3]
        arr = synth.synth_image(50_000 + k, h, w, c, n_set=1 << 30) if k % 2 else \
            rng.integers(0, 256, (h, w, c) if c > 1 else (h, w), dtype=np.uint8)
        im = _pil(arr, c)
        rp, rd = ph_mod.phash(im) & U64, ph_mod.dhash(im) & U64
        assert ref_py.phash(im) & U64 == rp and ref_py.dhash(im) & U64 == rd
        cp, cd, margin, _, _ = oracle.signature(arr)
        assert cd == rd
        assert cp == rp or margin < 1e-3
    for a, b in ((0, U64), (-1, 1), (123456789, 987654321), (-(1 << 63), (1 << 63) - 1)):
        assert ref_py.hamming64(a, b) == ph_mod.hamming64(a, b)


# ------------------------------------------------------------------ scanner


def _files(case):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.make_files(case["n"], case["seed"], case.get("ids_dupe", False))


def test_scanner_restatement_matches_reference_golden(golden_scanner):
    for case in golden_scanner["cases"]:
        files = [ref_py.FileRec(f["file_id"], f["path"], f["size"], f["width"], f["height"], f["phash"])
                 for f in _files(case)]
        got = ref_py.build_clusters(files, bucket_pair_cap=case.get("pair_cap"), **case["cfg"])
        assert normalise_clusters(got) == normalise_clusters(case["clusters"]), case["name"]


def test_reference_edges_equal_allpairs_with_band_predicate(golden_scanner):
    """The set identity the GPU join relies on: reference edges == {i<j: ham<=T and a band equal}."""
    for case in golden_scanner["cases"]:
        if "pair_cap" in case or case.get("ids_dupe") or "size_ratio" in case["cfg"]:
            continue
        files = _files(case)
        cfg = case["cfg"]
        recs = [ref_py.FileRec(f["file_id"], f["path"], f["size"], f["width"], f["height"], f["phash"]) for f in files]
        edges = ref_py.scan_edges(recs, **cfg)
        h = np.array([f["phash"] for f in files], dtype=np.uint64)
        oi, oj, od = oracle.hamming_join(h, cfg["hamming_threshold"], require_band=True,
                                         band_bits=cfg.get("band_bits", 16), band_count=cfg.get("band_count", 4))
        got = {(files[i]["file_id"], files[j]["file_id"]): int(d) for i, j, d in zip(oi, oj, od)}
        assert got == edges, case["name"]
        # and without the predicate the all-pairs set is a superset
        ai, aj, _ = oracle.hamming_join(h, cfg["hamming_threshold"], require_band=False)
        assert len(ai) >= len(oi)


@pytest.mark.reference
def test_live_reference_scanner_on_fresh_sets(reference_modules, monkeypatch):
    sc = reference_modules["scanner"]
    for seed, cfg, cap in ((101, {"hamming_threshold": 6}, None), (102, {"hamming_threshold": 9, "size_ratio": 0.7}, None),
                           (103, {"hamming_threshold": 8, "band_bits": 8, "band_count": 8}, 40)):
        case = {"n": 2500, "seed": seed}
        files = _files(case)
        if cap is None:
            monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)
        else:
            monkeypatch.setenv("KE_DUP_BUCKET_PAIR_CAP", str(cap))
        dfs = [sc.DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                                height=f["height"], phash=f["phash"]) for f in files]
        ref = sc.DuplicateScanner(sc.DuplicateScanConfig(**cfg)).build_clusters(dfs)
        ref_n = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in ref]
        recs = [ref_py.FileRec(f["file_id"], f["path"], f["size"], f["width"], f["height"], f["phash"]) for f in files]
        got = ref_py.build_clusters(recs, bucket_pair_cap=cap, **cfg)
        assert normalise_clusters(got) == normalise_clusters(ref_n)


# ------------------------------------------------------------------ SSIM


def _ssim_pair(cid, h, w):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ssim_pair(cid, h, w)


def test_ssim_restatements_agree_and_match_golden(golden_ssim):
    for case in golden_ssim["cases"]:
        if not isinstance(case["case"], int):
            continue
        a, b = _ssim_pair(case["case"], case["h"], case["w"])
        py = ref_py.ssim_of_planes(a, b)
        assert abs(py - case["ssim"]) < 1e-7  # scipy-based restatement is stable
        c32 = oracle.ssim_u8(a, b)
        cex = oracle.ssim_u8(a, b, exact=True)
        assert abs(c32 - py) < 2e-6, (case, c32, py)  # C float32 restatement == scipy restatement
        assert abs(cex - py) < 1e-5, (case, cex, py)  # exact arithmetic stays inside the parity bar


def test_ssim_restatement_is_pinned_by_an_independent_oracle_and_closed_forms(golden_ssim):
    """scikit-image itself is not installable here, so the restatement (ref_py, on scipy's uniform_filter) and the C
    oracle are pinned by things they did not generate: a float64 49-tap evaluation of the published definition that
    shares no code with them (oracle/ssim_independent.py), and closed forms that need no filter at all."""
    from oracle import ssim_independent as ind

    def all_three(a, b):
        return ref_py.ssim_of_planes(a, b), oracle.ssim_u8(a, b), oracle.ssim_u8(a, b, exact=True)

    # (1) brute force on the golden pairs (incl. the C4 shape 256x256 and the edge shapes 7x7, 8x31, 100x37)
    for case in golden_ssim["cases"]:
        if "h" not in case:
            continue
        a, b = _ssim_pair(case["case"], case["h"], case["w"])
        want = ind.mssim_bruteforce(a, b)
        for got in all_three(a, b):
            assert abs(got - want) <= 1e-5, (case, got, want)
        assert abs(case["ssim"] - want) <= 1e-5  # the committed golden value too
    # (2) constant vs constant: S = (2ab + C1) / (a^2 + b^2 + C1) everywhere, variances exactly 0
    for la, lb, shape in ((0, 0, (7, 7)), (255, 255, (9, 13)), (200, 10, (64, 64)), (17, 18, (31, 8)), (0, 255, (16, 16)),
                          (128, 128, (256, 256))):
        a = np.full(shape, la, np.uint8)
        b = np.full(shape, lb, np.uint8)
        want = ind.closed_form_constants(la, lb)
        for got in all_three(a, b) + (ind.mssim_bruteforce(a, b),):
            assert abs(got - want) <= 1e-6, (la, lb, got, want)
    # (3) identical images: exactly the definition's 1
    rng = np.random.default_rng(3)
    for shape in ((7, 7), (40, 23), (256, 256)):
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        for got in all_three(a, a) + (ind.mssim_bruteforce(a, a),):
            assert abs(got - 1.0) <= 1e-6
    # (4) inverted two-level checkerboard: mu_b = 1 - mu_a, var_a = var_b = (p - q)^2 * 25/98, cov = -var
    for (h, w, p, q) in ((16, 16, 200, 50), (31, 8, 255, 0), (64, 40, 130, 120), (7, 7, 10, 240)):
        a = ind.checkerboard(h, w, p, q)
        b = (255 - a.astype(np.int32)).astype(np.uint8)
        want = ind.closed_form_inverted_checkerboard(h, w, p, q)
        assert abs(ind.mssim_bruteforce(a, b) - want) <= 1e-12
        for got in all_three(a, b):
            assert abs(got - want) <= 1e-5, (h, w, p, q, got, want)
    # (5) the reference's own behavioural values (tests/dup/test_refine.py:24-46) from the independent oracle
    from PIL import Image, ImageEnhance

    base = Image.new("RGB", (64, 64), color=(200, 10, 10))
    pa, pb = ref_py.ssim_planes(base, ImageEnhance.Brightness(base).enhance(1.02))
    assert ind.mssim_bruteforce(pa, pb) > 0.95
    pa, pb = ref_py.ssim_planes(Image.new("RGB", (64, 64), (0, 255, 0)), Image.new("RGB", (64, 64), (0, 0, 255)))
    assert ind.mssim_bruteforce(pa, pb) < 0.95


def test_gaussian_window_restatement_sanity():
    """The optional Gaussian window (skimage gaussian_weights=True; NOT the reference's path): definitional checks of
    the restatement the CUDA flag is compared with."""
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (40, 33), dtype=np.uint8)
    assert abs(ref_py.ssim_gaussian_of_planes(a, a) - 1.0) <= 1e-6
    c, d = np.full((16, 20), 200, np.uint8), np.full((16, 20), 10, np.uint8)
    want = (2 * (200 / 255) * (10 / 255) + 1e-4) / ((200 / 255) ** 2 + (10 / 255) ** 2 + 1e-4)
    assert abs(ref_py.ssim_gaussian_of_planes(c, d) - want) <= 1e-6
    with pytest.raises(ValueError):
        ref_py.ssim_gaussian_of_planes(a[:10, :10], a[:10, :10])
    b = np.clip(a.astype(np.int32) + rng.integers(-6, 7, a.shape), 0, 255).astype(np.uint8)
    g, u = ref_py.ssim_gaussian_of_planes(a, b), ref_py.ssim_of_planes(a, b)
    assert 0.0 < g < 1.0 and 0.0 < u < 1.0 and abs(g - u) < 0.2


def test_ssim_reference_behavioural_pins():
    """tests/dup/test_refine.py:24-46 of the reference, through the restatement."""
    from PIL import Image, ImageEnhance

    base = Image.new("RGB", (64, 64), color=(200, 10, 10))
    var = ImageEnhance.Brightness(base).enhance(1.02)
    s = ref_py.compute_ssim(base, var)
    assert s > 0.95
    assert ref_py.refine_decision(s, None) == (True, "ssim>=0.9")
    s2 = ref_py.compute_ssim(Image.new("RGB", (64, 64), (0, 255, 0)), Image.new("RGB", (64, 64), (0, 0, 255)))
    assert ref_py.refine_decision(s2, 0.0, ssim_thr=0.95, orb_thr=0.5) == (False, "below thresholds")
    assert ref_py.refine_decision(None, None, errors=("ssim unavailable", "orb unavailable")) == \
        (False, "ssim unavailable, orb unavailable")
    pa, pb = ref_py.ssim_planes(base, var)
    assert abs(oracle.ssim_u8(pa, pb, exact=True) - s) < 1e-5
    with pytest.raises(ValueError):
        ref_py.structural_similarity(np.zeros((6, 10), np.float32), np.zeros((6, 10), np.float32))


def test_ssim_exact_vs_float32_noise_over_varied_pairs():
    """How far exact arithmetic sits from the reference's float32 path (budget: 1e-5)."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for k in range(40):
        h, w = int(rng.integers(7, 96)), int(rng.integers(7, 96))
        a = synth.synth_image(7000 + k, h, w, 1)
        mode = k % 4
        if mode == 0:
            b = np.clip(a.astype(int) + rng.integers(-4, 5, a.shape), 0, 255).astype(np.uint8)
        elif mode == 1:
            b = synth.synth_image(7500 + k, h, w, 1)
        elif mode == 2:
            a = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
            b = np.clip(a.astype(int) + int(rng.integers(-3, 4)), 0, 255).astype(np.uint8)
        else:
            b = rng.integers(0, 256, (h, w), dtype=np.uint8)
        worst = max(worst, abs(oracle.ssim_u8(a, b, exact=True) - ref_py.ssim_of_planes(a, b)))
    assert worst < 5e-6, worst


def test_cluster_matches_restatement():
    ms = [(1, 2, True), (2, 3, True), (4, 5, True), (3, 5, False)]
    assert ref_py.cluster_matches(ms) == [(1, [1, 2, 3]), (4, [4, 5])]


# ------------------------------------------------------------------ the live reference's dup.refine / dup.cluster
# (importable only through oracle/skimage_shim: scikit-image is absent from this image)


def test_orb_cross_check_restatement_equals_live_opencv():
    """N2's oracle: the restated cross-check (mutual nearest neighbours, first index wins ties) equals
    cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match on descriptor sets with many ties, and on real ORB descriptors."""
    import cv2

    rng = np.random.default_rng(0)
    for trial in range(120):
        na, nb = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        bits = int(rng.integers(1, 4))  # few distinct byte values -> many equal distances
        da = rng.integers(0, 1 << bits, (na, 32)).astype(np.uint8)
        db = rng.integers(0, 1 << bits, (nb, 32)).astype(np.uint8)
        if trial % 3 == 0:
            db[: min(na, nb)] = da[: min(na, nb)]
        want = sorted((m.queryIdx, m.trainIdx, int(m.distance)) for m in
                      cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db))
        assert ref_py.orb_cross_check(da, db) == want, trial
    from kobato_b200 import synth

    orb = cv2.ORB_create()
    a = synth.synth_image(3, 256, 256, 1)
    b = np.roll(a, 3, axis=1)
    _, da = orb.detectAndCompute(a, None)
    _, db = orb.detectAndCompute(b, None)
    if da is not None and db is not None:
        want = sorted((m.queryIdx, m.trainIdx, int(m.distance)) for m in
                      cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db))
        assert ref_py.orb_cross_check(da, db) == want and len(want) > 5


@pytest.mark.reference
def test_reference_refine_and_cluster_tests_run_through_the_skimage_shim():
    """The reference's OWN tests for the SSIM / clustering half of the path (tests/dup/test_refine.py,
    tests/dup/test_cluster.py) pass with ``skimage.metrics.structural_similarity`` answered by the oracle restatement."""
    import os
    import subprocess

    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(ROOT / "oracle" / "skimage_shim"), str(ROOT), str(REFERENCE_SRC)])
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
                          str(REFERENCE_SRC.parent / "tests" / "dup" / "test_refine.py"),
                          str(REFERENCE_SRC.parent / "tests" / "dup" / "test_cluster.py")],
                         capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout


@pytest.mark.reference
def test_dropin_refine_and_cluster_equal_the_live_reference(tmp_path, monkeypatch):
    """Same files through the live ``dup.refine.refine_pair`` / ``dup.cluster.ClusterBuilder`` and through the drop-ins
    (SSIM arithmetic of the drop-in stubbed by the oracle here; the CUDA kernel is compared with it in the GPU tests):
    identical RefinedMatch fields, reasons and clusters."""
    from PIL import Image, ImageEnhance

    from kobato_b200.dup import cluster as kcluster
    from kobato_b200.dup import refine as krefine

    shim = str(ROOT / "oracle" / "skimage_shim")
    for p in (shim, str(REFERENCE_SRC)):
        if p not in sys.path:
            sys.path.append(p)
    import importlib

    ref_refine = importlib.import_module("dup.refine")
    ref_cluster = importlib.import_module("dup.cluster")
    monkeypatch.setattr(krefine, "_compute_ssim", lambda a, b: ref_py.compute_ssim(a, b))
    from kobato_b200 import ops as kops

    monkeypatch.setattr(kops, "orb_match_pairs", lambda a, b, **kw: ref_py.orb_match_counts(a, b))  # GPU matcher stand-in

    rng = np.random.default_rng(9)
    paths = []
    base = (rng.random((72, 96, 3)) * 255).astype(np.uint8)
    for k in range(8):
        arr = base.copy() if k % 2 == 0 else (rng.random((72, 96, 3)) * 255).astype(np.uint8)
        img = Image.fromarray(arr)
        if k in (2, 4):
            img = ImageEnhance.Brightness(img).enhance(1.0 + 0.01 * k)
        if k == 6:
            img = img.resize((80, 60))
        p = tmp_path / f"f{k}.png"
        img.save(p)
        paths.append(p)
    broken = tmp_path / "broken.png"
    broken.write_bytes(b"nope")
    tiny = tmp_path / "tiny.png"
    Image.new("RGB", (5, 5), (9, 9, 9)).save(tiny)
    cases = [(0, 2), (0, 4), (0, 1), (1, 3), (2, 4), (0, 6), (3, 5)]
    got_all, want_all = [], []
    for thr in (ref_refine.RefinementThresholds(), ref_refine.RefinementThresholds(ssim=0.5, orb=0.9)):
        kthr = krefine.RefinementThresholds(ssim=thr.ssim, orb=thr.orb)
        for a, b in cases:
            want = ref_refine.refine_pair(a, b, paths[a], paths[b], thresholds=thr)
            got = krefine.refine_pair(a, b, paths[a], paths[b], thresholds=kthr)
            assert (got.file_id_a, got.file_id_b, got.is_duplicate, got.reason) == \
                (want.file_id_a, want.file_id_b, want.is_duplicate, want.reason), (a, b)
            assert (got.ssim is None) == (want.ssim is None) and (got.orb_ratio is None) == (want.orb_ratio is None)
            if want.ssim is not None:
                assert abs(got.ssim - want.ssim) <= 1e-9
            if want.orb_ratio is not None:
                assert got.orb_ratio == want.orb_ratio
            got_all.append(got)
            want_all.append(want)
    assert ref_refine.refine_pair(1, 2, paths[0], broken) is None and krefine.refine_pair(1, 2, paths[0], broken) is None
    w, g = ref_refine.refine_pair(1, 2, tiny, tiny), krefine.refine_pair(1, 2, tiny, tiny)
    assert (g.ssim, g.is_duplicate, g.reason) == (w.ssim, w.is_duplicate, w.reason)  # side < 7: "ssim unavailable, ..."
    want_c = ref_cluster.ClusterBuilder().build(want_all)
    got_c = kcluster.ClusterBuilder().build(got_all)
    assert [(c.representative, c.members) for c in got_c] == [(c.representative, c.members) for c in want_c]
    assert [(c.representative, c.members) for c in want_c] == \
        ref_py.cluster_matches((m.file_id_a, m.file_id_b, m.is_duplicate) for m in want_all)
