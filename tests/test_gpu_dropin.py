"""GPU tests of the drop-in modules (reference-facing API): same calls a user of the reference
makes on sig.phash / core.fastsig / core.signature / dup.scanner / dup.refine, answered by the CUDA
kernels and compared with the oracle / the reference's golden vectors."""
from __future__ import annotations

import sqlite3
from pathlib import Path

import numpy as np
import pytest
from conftest import normalise_clusters

import oracle
from kobato_b200 import pipeline, synth
from kobato_b200.core import fastsig, signature
from kobato_b200.dup import cluster as kcluster
from kobato_b200.dup import refine as krefine
from kobato_b200.dup import scanner as kscanner
from kobato_b200.sig import phash as kphash
from oracle import ref_py

pytestmark = pytest.mark.gpu
U64 = (1 << 64) - 1
MODES = {1: "L", 3: "RGB", 4: "RGBA"}


def test_sig_phash_matches_reference_golden(golden_phash):
    from PIL import Image

    images, want = [], []
    for case in golden_phash["cases"]:
        arr = synth.synth_image(case["index"], case["h"], case["w"], case["c"], n_set=1 << 30)
        images.append(Image.fromarray(arr, MODES[case["c"]]))
        want.append((int(case["phash"], 16), int(case["dhash"], 16)))
    got = kphash.phash_dhash_many(images)  # mixed geometries in one call
    for (p, d), (wp, wd), case in zip(got, want, golden_phash["cases"]):
        assert (p & U64, d & U64) == (wp, wd), case
        assert -(1 << 63) <= p < (1 << 63)
    im = images[0]
    assert kphash.phash(im) & U64 == want[0][0] and kphash.dhash(im) & U64 == want[0][1]
    # modes the kernel does not take directly go through the reference's own convert("L")
    for mode in ("P", "1", "CMYK", "I;16", "LA", "F"):
        try:
            conv = im.convert(mode)
        except (ValueError, OSError):
            continue
        assert kphash.phash(conv) == ref_py.phash(conv), mode
        assert kphash.dhash(conv) == ref_py.dhash(conv), mode


def test_reference_signature_tests(tmp_path):
    """tests/core/test_image_signature.py:30-55 of the reference against the CUDA path."""
    from PIL import Image

    conn = sqlite3.connect(":memory:")
    conn.row_factory = sqlite3.Row
    conn.execute("CREATE TABLE signatures (file_id INTEGER PRIMARY KEY, phash_u64 INTEGER NOT NULL, dhash_u64 INTEGER NOT NULL)")
    for seed in range(1, 21):
        arr = (np.random.default_rng(seed).random((64, 64, 3)) * 255).astype("uint8")
        img = Image.fromarray(arr)
        p, d = signature.compute_signatures_from_image(img)
        assert (p, d) == (ref_py.phash(img), ref_py.dhash(img))
        assert signature.ensure_signatures(conn, seed, image=img, upsert=signature._upsert_signatures) is True
    assert conn.execute("SELECT COUNT(*) FROM signatures").fetchone()[0] == 20


def test_compute_signatures_mp_on_real_files(tmp_path):
    from PIL import Image

    tasks = []
    for k in range(40):
        h, w = (96, 128) if k % 3 else (77, 50 + k)
        arr = synth.synth_image(k, h, w, 3)
        path = tmp_path / f"f{k:03d}.{'png' if k % 2 else 'bmp'}"
        Image.fromarray(arr).save(path)
        tasks.append((100 + k, str(path)))
    (tmp_path / "bad.png").write_bytes(b"garbage")
    tasks.insert(7, (999, str(tmp_path / "bad.png")))
    seen = []
    got = fastsig.compute_signatures_mp(tasks, max_workers=4, chunksize=4, progress=lambda d, t: seen.append((d, t)))
    assert [g[0] for g in got] == [100 + k for k in range(40)]
    assert seen[-1] == (41, 41)
    for fid, ph, dh in got:
        im = Image.open(tmp_path / next(Path(p).name for f, p in tasks if f == fid))
        assert (ph, dh) == (ref_py.phash(im), ref_py.dhash(im))
    db = tmp_path / "s.db"
    with sqlite3.connect(db) as c:
        c.execute("CREATE TABLE signatures (file_id INTEGER PRIMARY KEY, phash_u64 INTEGER NOT NULL, dhash_u64 INTEGER NOT NULL)")
    out = fastsig.fast_fill_missing_signatures(str(db), tasks[:5])
    assert sqlite3.connect(db).execute("SELECT COUNT(*) FROM signatures").fetchone()[0] == len(out) == 5


def _golden_files(case):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.make_files(case["n"], case["seed"], case.get("ids_dupe", False))


def test_scanner_matches_reference_golden_clusters(golden_scanner, monkeypatch):
    for case in golden_scanner["cases"]:
        if "pair_cap" in case:
            monkeypatch.setenv("KE_DUP_BUCKET_PAIR_CAP", str(case["pair_cap"]))
        else:
            monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)
        files = [kscanner.DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                                        height=f["height"], phash=f["phash"]) for f in _golden_files(case)]
        clusters = kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(**case["cfg"])).build_clusters(files)
        got = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in clusters]
        assert normalise_clusters(got) == normalise_clusters(case["clusters"]), case["name"]


def test_scanner_at_70k_equals_the_lsh_restatement():
    """README scale (C2's hash count): cluster membership equals the reference's LSH scan."""
    n = 70_000
    h = synth.synth_hashes(n, seed=77, planted=0.05)
    rng = np.random.default_rng(0)
    sizes = rng.integers(1000, 5_000_000, n)
    files = [kscanner.DuplicateFile(file_id=i + 1, path=Path(f"d/f{i}.jpg"), size=int(sizes[i]), width=100, height=100,
                                    phash=int(h[i])) for i in range(n)]
    got = kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(hamming_threshold=8)).build_clusters(files)
    recs = [ref_py.FileRec(f.file_id, str(f.path), f.size, f.width, f.height, f.phash) for f in files]
    want = ref_py.build_clusters(recs, hamming_threshold=8)
    got_n = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in got]
    assert normalise_clusters(got_n) == normalise_clusters(want)
    assert len(got) > 1000


def test_refine_reference_tests_on_gpu(tmp_path):
    """tests/dup/test_refine.py:24-64 of the reference against the CUDA SSIM."""
    from PIL import Image, ImageEnhance

    pa, pb = tmp_path / "a.png", tmp_path / "b.png"
    Image.new("RGB", (64, 64), color=(200, 10, 10)).save(pa)
    ImageEnhance.Brightness(Image.open(pa).convert("RGB")).enhance(1.02).save(pb)
    r = krefine.refine_pair(1, 2, pa, pb)
    assert isinstance(r, krefine.RefinedMatch) and r.is_duplicate and r.ssim > 0.95
    assert abs(r.ssim - ref_py.compute_ssim(Image.open(pa), Image.open(pb))) <= 1e-5
    g, b = tmp_path / "g.png", tmp_path / "bl.png"
    Image.new("RGB", (64, 64), (0, 255, 0)).save(g)
    Image.new("RGB", (64, 64), (0, 0, 255)).save(b)
    r = krefine.refine_pair(1, 3, g, b, thresholds=krefine.RefinementThresholds(ssim=0.95, orb=0.5))
    assert not r.is_duplicate and r.reason == "below thresholds"
    tiny_a, tiny_b = tmp_path / "ta.png", tmp_path / "tb.png"
    Image.new("RGB", (5, 40), (1, 2, 3)).save(tiny_a)
    Image.new("RGB", (5, 40), (1, 2, 3)).save(tiny_b)
    r = krefine.refine_pair(1, 2, tiny_a, tiny_b)  # side < 7: skimage raises -> "ssim unavailable"
    assert r.ssim is None and not r.is_duplicate and r.reason == "ssim unavailable"


def test_refine_pairs_batch_matches_per_pair_and_oracle(tmp_path):
    from PIL import Image

    paths = []
    for k in range(12):
        h, w = ((90, 120), (64, 64), (120, 90))[k % 3]
        arr = synth.synth_image(k, h, w, 3, n_set=12, planted=0.5)
        p = tmp_path / f"i{k}.png"
        Image.fromarray(arr).save(p)
        paths.append(p)
    (tmp_path / "bad.png").write_bytes(b"zz")
    pairs = [(a, b, paths[a], paths[b]) for a in range(12) for b in range(a + 1, 12) if (a + b) % 4 == 0]
    pairs.append((50, 51, paths[0], tmp_path / "bad.png"))
    batch = krefine.refine_pairs_batch(pairs, max_workers=4)
    assert batch[-1] is None
    for rec, (a, b, pa, pb) in zip(batch[:-1], pairs[:-1]):
        want = ref_py.compute_ssim(Image.open(pa).convert("RGB"), Image.open(pb).convert("RGB"))
        assert abs(rec.ssim - want) <= 1e-5
        single = krefine.refine_pair(a, b, pa, pb)
        assert abs(single.ssim - rec.ssim) <= 1e-12 and single.is_duplicate == rec.is_duplicate
        assert single.reason == rec.reason and single.orb_ratio == rec.orb_ratio
    clusters = kcluster.ClusterBuilder().build([m for m in batch if m is not None])
    assert clusters == sorted(clusters, key=lambda c: c.representative)


def test_pipeline_scan_matches_oracle_end_to_end():
    import torch

    n, h, w = 300, 96, 96
    host = synth.synth_images(0, n, h, w, 3, n_set=n, planted=0.2)
    bank = torch.from_numpy(host).cuda()
    pinned = torch.from_numpy(host).pin_memory()
    a = pipeline.scan(bank, threshold=8, ssim_threshold=0.9)
    b = pipeline.scan(torch.empty_like(bank), host_images=pinned, threshold=8, ssim_threshold=0.9, chunk_images=64)
    ph = a.phash.cpu().numpy().view(np.uint64)
    assert np.array_equal(ph, b.phash_host.view(np.uint64)) and np.array_equal(a.cand_i, b.cand_i)
    assert np.allclose(a.ssim, b.ssim, atol=1e-12, rtol=0)
    wi, wj, wd = oracle.hamming_join(ph, 8, require_band=True)
    assert np.array_equal(a.cand_i, wi) and np.array_equal(a.cand_j, wj) and np.array_equal(a.cand_d, wd)
    want = np.array([ref_py.ssim_of_planes(oracle.to_l(host[i]), oracle.to_l(host[j])) for i, j in zip(wi, wj)])
    assert np.all(np.abs(a.ssim - want) <= 1e-5)
    safe = np.abs(want - 0.9) > 1e-5
    assert np.array_equal(a.accepted[safe], (want >= 0.9)[safe])
    assert a.clusters.as_list() == ref_py.cluster_matches(zip(wi.tolist(), wj.tolist(), a.accepted.tolist()))
    assert b.bytes_h2d == host.nbytes and b.bytes_d2h > 16 * n
