"""CPU tests of the host side: the C ABI library loads and exports what the header declares, the
host-only entry points work without a GPU, and the drop-in modules keep the reference's contracts
(the GPU calls are replaced by the oracle through the modules' injection seams)."""
from __future__ import annotations

import ctypes as C
import re
import sqlite3
from pathlib import Path

import numpy as np
import pytest
from conftest import ROOT, normalise_clusters

import oracle
from kobato_b200 import _native as nat
from kobato_b200 import synth
from kobato_b200.core import fastsig, signature
from kobato_b200.dup import cluster as kcluster
from kobato_b200.dup import refine as krefine
from kobato_b200.dup import scanner as kscanner
from kobato_b200.sig import phash as kphash
from oracle import ref_py

U64 = (1 << 64) - 1


# ------------------------------------------------------------------ C ABI


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "kobato_b200.h").read_text()
    declared = set(re.findall(r"\b(ke_[a-z0-9_]+)\s*\(", header))
    lib = nat.load()
    assert lib.ke_abi_version() == nat.KE_ABI_VERSION == 2
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_host_resample_tables_equal_the_pinned_oracle():
    lib = nat.load()
    for in_size, out_size in ((512, 32), (512, 9), (512, 8), (33, 32), (7, 32), (32, 32), (9, 9), (1000, 32), (1, 32),
                              (4096, 9)):
        ks = lib.ke_resample_ksize(in_size, out_size)
        kk = np.zeros((out_size, ks), np.int32)
        bd = np.zeros((out_size, 2), np.int32)
        assert lib.ke_resample_table(in_size, out_size, kk.ctypes.data, bd.ctypes.data, ks) == 0
        okk, obd = oracle.resample_table(in_size, out_size)
        assert np.array_equal(kk, okk) and np.array_equal(bd, obd), (in_size, out_size)
    assert lib.ke_resample_ksize(0, 32) < 0
    assert lib.ke_resample_table(512, 32, None, None, 97) == nat.KE_E_INVALID
    assert b"bad arguments" in lib.ke_last_error()


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(nat.KobatoNativeError, match="no CPU fallback"):
        nat.Context(0)
    with pytest.raises(nat.KobatoNativeError):
        kphash.phash(np.zeros((16, 16, 3), np.uint8))


# ------------------------------------------------------------------ sig / core drop-ins


def _oracle_many(images):
    """Stand-in for the GPU call: the reference's own path (PIL + cv2) on the decoded arrays."""
    from PIL import Image

    out = []
    for im in images:
        arr = kphash._decoded_array(im)
        pil = Image.fromarray(arr if arr.ndim == 2 or arr.shape[2] > 1 else arr[..., 0])
        out.append((ref_py.phash(pil), ref_py.dhash(pil)))
    return out


def test_hamming64_and_signed_wrap_follow_the_reference_tables():
    assert kphash.hamming64(0, U64) == 64
    assert kphash.hamming64(-1, 1) == 63
    assert kphash.hamming64(-(1 << 63), (1 << 63) - 1) == 64
    for value, expected in ((0, 0), (U64, -1), (1 << 63, -(1 << 63)), ((1 << 63) - 1, (1 << 63) - 1)):
        assert fastsig._to_signed64(value) == expected
        assert signature._to_signed64(value) == expected
        assert kphash._to_signed(value) == expected


def test_decoded_array_modes():
    from PIL import Image

    rgb = Image.fromarray(synth.synth_image(1, 20, 30, 3))
    assert kphash._decoded_array(rgb).shape == (20, 30, 3)
    assert kphash._decoded_array(rgb.convert("RGBA")).shape == (20, 30, 4)
    assert kphash._decoded_array(rgb.convert("L")).shape == (20, 30)
    pal = rgb.convert("P")
    assert np.array_equal(kphash._decoded_array(pal), np.asarray(pal.convert("L")))  # host convert like the reference
    with pytest.raises(ValueError):
        kphash._decoded_array(np.zeros((4, 4, 2), np.uint8))


def _write_images(tmp_path, count, size=(48, 40)):
    from PIL import Image

    tasks = []
    for k in range(count):
        p = tmp_path / f"img{k:03d}.png"
        Image.fromarray(synth.synth_image(k, size[1], size[0] + (k % 3), 3)).save(p)
        tasks.append((k + 1, str(p)))
    return tasks


def test_compute_signatures_contract(tmp_path, monkeypatch):
    """Ordered results, failures dropped, progress every 200 + at the end (src/core/fastsig.py:65-99)."""
    from PIL import Image

    monkeypatch.setattr(fastsig, "phash_dhash_many", _oracle_many)
    tasks = _write_images(tmp_path, 7)
    (tmp_path / "broken.png").write_bytes(b"not an image")
    tasks.insert(3, (99, str(tmp_path / "broken.png")))
    tasks.insert(5, (98, str(tmp_path / "missing.png")))
    tasks.append((97, str(tmp_path)))  # a directory
    seen = []
    res = fastsig.compute_signatures_mp(tasks, max_workers=2, chunksize=2, progress=lambda d, t: seen.append((d, t)))
    assert [r[0] for r in res] == [1, 2, 3, 4, 5, 6, 7]
    assert seen == [(len(tasks), len(tasks))]
    for fid, ph, dh in res:
        im = Image.open(tasks[[t[0] for t in tasks].index(fid)][1])
        assert ph == ref_py.phash(im) and dh == ref_py.dhash(im)
        assert -(1 << 63) <= ph < (1 << 63)
    assert fastsig.compute_signatures_mp([]) == []


def test_compute_signatures_progress_and_cancel(tmp_path, monkeypatch):
    monkeypatch.setattr(fastsig, "phash_dhash_many", lambda imgs: [(1, 2)] * len(list(imgs)))
    monkeypatch.setattr(fastsig, "_decode_worker", lambda task: (task[0], np.zeros((8, 8), np.uint8)))
    tasks = [(k, f"x{k}") for k in range(450)]
    seen = []
    res = fastsig.compute_signatures_mp(tasks, max_workers=2, chunksize=10, progress=lambda d, t: seen.append(d))
    assert len(res) == 450 and seen == [200, 400, 450]
    calls = {"n": 0}

    def cancel():
        calls["n"] += 1
        return calls["n"] > 25

    part = fastsig.compute_signatures_mp(tasks, max_workers=2, chunksize=5, cancel_fn=cancel)
    assert 0 < len(part) < 450
    assert [r[0] for r in part] == list(range(len(part)))  # an ordered prefix


def test_bulk_upsert_and_fast_fill(tmp_path, monkeypatch):
    db = tmp_path / "t.db"
    conn = sqlite3.connect(db)
    conn.execute("CREATE TABLE signatures (file_id INTEGER PRIMARY KEY, phash_u64 INTEGER NOT NULL, dhash_u64 INTEGER NOT NULL)")
    assert fastsig.bulk_upsert_signatures(conn, []) == 0
    fastsig.bulk_upsert_signatures(conn, [(1, U64, 5), (2, 1 << 63, 7)])
    fastsig.bulk_upsert_signatures(conn, [(1, 3, 4)])
    assert conn.execute("SELECT * FROM signatures ORDER BY file_id").fetchall() == [(1, 3, 4), (2, -(1 << 63), 7)]
    conn.close()
    monkeypatch.setattr(fastsig, "compute_signatures_mp", lambda items, **kw: [(5, -1, 2)])
    assert fastsig.fast_fill_missing_signatures(str(db), [(5, "x")]) == [(5, -1, 2)]
    assert sqlite3.connect(db).execute("SELECT phash_u64 FROM signatures WHERE file_id=5").fetchone() == (-1,)
    assert fastsig.fast_fill_missing_signatures(str(db), [(6, "x")], apply_to_db=False) == [(5, -1, 2)]
    assert sqlite3.connect(db).execute("SELECT COUNT(*) FROM signatures").fetchone() == (3,)


def test_ensure_signatures_contract(monkeypatch):
    """tests/core/test_image_signature.py:30-55 of the reference, hashing stubbed by the oracle."""
    from PIL import Image

    monkeypatch.setattr(signature, "phash_dhash_many", _oracle_many)
    conn = sqlite3.connect(":memory:")
    conn.row_factory = sqlite3.Row
    conn.execute("CREATE TABLE signatures (file_id INTEGER PRIMARY KEY, phash_u64 INTEGER NOT NULL, dhash_u64 INTEGER NOT NULL)")
    rng = np.random.default_rng(1)
    for fid in range(1, 21):
        img = Image.fromarray((rng.random((64, 64, 3)) * 255).astype("uint8"))
        assert signature.ensure_signatures(conn, fid, image=img, upsert=signature._upsert_signatures) is True
        p, d = signature.compute_signatures_from_image(img)
        assert -(1 << 63) <= p <= (1 << 63) - 1 and -(1 << 63) <= d <= (1 << 63) - 1
        row = conn.execute("SELECT phash_u64 FROM signatures WHERE file_id=?", (fid,)).fetchone()
        assert kscanner._parse_phash_any(row["phash_u64"]) == p & U64
    assert conn.execute("SELECT COUNT(*) FROM signatures").fetchone()[0] == 20
    assert signature.ensure_signatures(conn, 1) is True  # exists -> untouched
    assert signature.ensure_signatures(conn, 500) is False  # nothing to compute from
    assert signature.ensure_signatures(conn, 501, path="nope.png", loader=lambda p: None) is False


# ------------------------------------------------------------------ dup.scanner drop-in


def _oracle_join(hashes, cfg, allow):
    """Injected in place of the GPU join: all pairs + band predicate + allow mask, from the oracle."""
    i, j, d = oracle.hamming_join(hashes, cfg.hamming_threshold, require_band=False, threads=4)
    keep = np.zeros(len(i), bool)
    mask = (1 << cfg.band_bits) - 1
    for k, (a, b) in enumerate(zip(i.tolist(), j.tolist())):
        x = int(hashes[a]) ^ int(hashes[b])
        for band in range(cfg.band_count):
            if ((x >> (band * cfg.band_bits)) & mask) == 0 and (allow is None or ((int(allow[a]) >> band) & 1)):
                keep[k] = True
                break
    return i[keep], j[keep], d[keep]


def _golden_files(case):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.make_files(case["n"], case["seed"], case.get("ids_dupe", False))


def test_scanner_dropin_matches_reference_golden_clusters(golden_scanner, monkeypatch):
    for case in golden_scanner["cases"]:
        if "pair_cap" in case:
            monkeypatch.setenv("KE_DUP_BUCKET_PAIR_CAP", str(case["pair_cap"]))
        else:
            monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)
        files = [kscanner.DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                                        height=f["height"], phash=f["phash"]) for f in _golden_files(case)]
        scanner = kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(**case["cfg"]), join=_oracle_join)
        got = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]}
               for c in scanner.build_clusters(files)]
        assert normalise_clusters(got) == normalise_clusters(case["clusters"]), case["name"]


def _oracle_scan_table(hashes, ids, sizes, **kw):
    """Injected in place of ops.scan_table (ke_scan_table_host): the oracle's restatement of it."""
    return ref_py.scan_table(hashes, ids, sizes, **kw)


def test_scanner_table_path_matches_reference_golden_clusters(golden_scanner, monkeypatch):
    """The table path (N3: columns -> members grouped by component -> cluster objects for the members only) against the
    live reference's golden clusters; cases with repeated file ids must route themselves to the legacy path."""
    table_cases = 0
    for case in golden_scanner["cases"]:
        if "pair_cap" in case:
            monkeypatch.setenv("KE_DUP_BUCKET_PAIR_CAP", str(case["pair_cap"]))
        else:
            monkeypatch.delenv("KE_DUP_BUCKET_PAIR_CAP", raising=False)
        recs = _golden_files(case)
        files = [kscanner.DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                                        height=f["height"], phash=f["phash"]) for f in recs]
        calls = []

        def scan(*a, **kw):
            calls.append(1)
            return _oracle_scan_table(*a, **kw)

        distinct = len({f.file_id for f in files}) == len(files)
        scanner = kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(**case["cfg"]), scan_table=scan,
                                            join=None if distinct else _oracle_join)
        got = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]}
               for c in scanner.build_clusters(files)]
        assert normalise_clusters(got) == normalise_clusters(case["clusters"]), case["name"]
        assert bool(calls) == distinct, case["name"]
        table_cases += distinct
        if distinct:  # the same through the column entry point: no DuplicateFile exists before a row is a member
            made = []

            def make_file(row):
                made.append(row)
                return files[row]

            ids = np.array([f.file_id for f in files], np.int64)
            ph = np.array([f.phash for f in files], np.uint64).view(np.int64)  # what SQLite stores
            sz = np.array([f.size or 0 for f in files], np.int64)
            cols = scanner.build_clusters_from_columns(ids, ph, sz, make_file=make_file)
            got2 = [{"keeper": c.keeper_id, "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in cols]
            assert normalise_clusters(got2) == normalise_clusters(case["clusters"]), case["name"]
            assert len(made) == sum(len(c["members"]) for c in case["clusters"])
    assert table_cases >= 5
    with pytest.raises(ValueError):
        kscanner.DuplicateScanner(kscanner.DuplicateScanConfig(), scan_table=_oracle_scan_table).build_clusters_from_columns(
            np.array([1, 1]), np.array([5, 5]), None, make_file=lambda r: None)


def test_scanner_reference_unit_cases():
    """The reference's own scanner tests (tests/dup/test_scanner.py:31-163) against the drop-in."""
    DF, Cfg = kscanner.DuplicateFile, kscanner.DuplicateScanConfig

    def mk(fid, path, size, w, h, ph, emb=None):
        return DF(file_id=fid, path=Path(path), size=size, width=w, height=h, phash=ph, embedding=emb)

    base = 0xFFFF_FFFF_0000_0000
    files = [mk(1, "a.jpg", 1000, 640, 480, base), mk(2, "b.png", 2000, 640, 480, base ^ 1),
             mk(3, "c.jpg", 1500, 800, 600, base ^ 2)]
    clusters = kscanner.DuplicateScanner(Cfg(hamming_threshold=4), join=_oracle_join).build_clusters(files)
    assert len(clusters) == 1 and clusters[0].keeper_id == 2
    assert {e.file.file_id for e in clusters[0].files} == {1, 2, 3}
    assert all(e.best_hamming is not None for e in clusters[0].files)

    base = 0xAAAA_AAAA_AAAA_AAAA
    files = [mk(1, "small.jpg", 100, 100, 100, base), mk(2, "large.jpg", 1000, 100, 100, base, (0.9, 0.1)),
             mk(3, "cosine_a.jpg", 800, 200, 200, base ^ 1, (0.89, 0.11)),
             mk(4, "cosine_b.jpg", 820, 200, 200, base ^ 2, (0.88, 0.12)),
             mk(5, "cosine_bad.jpg", 830, 200, 200, base ^ 3, (-0.5, 0.3))]
    clusters = kscanner.DuplicateScanner(Cfg(hamming_threshold=4, size_ratio=0.5, cosine_threshold=0.9),
                                         join=_oracle_join).build_clusters(files)
    assert len(clusters) == 1 and {e.file.file_id for e in clusters[0].files} == {2, 3, 4}

    base = 0x1234_5678_0000_0000
    files = [mk(1, "a.jpg", 100, 10, 10, base, (1.0,)), mk(2, "b.jpg", 100, 10, 10, base, (1.0, 0.0))]
    clusters = kscanner.DuplicateScanner(Cfg(hamming_threshold=0, cosine_threshold=0.99), join=_oracle_join).build_clusters(files)
    assert len(clusters) == 1 and {e.file.file_id for e in clusters[0].files} == {1, 2}
    assert kscanner.DuplicateScanner(Cfg(), join=_oracle_join).build_clusters([]) == []
    assert kscanner.DuplicateScanner(Cfg(), join=_oracle_join).build_clusters([mk(1, "a.jpg", 1, 1, 1, 5)]) == []


def test_from_row_and_config_validation():
    DF = kscanner.DuplicateFile
    blob = DF.from_row({"file_id": 10, "path": "blob.png", "size": 12, "width": 3, "height": 4,
                        "phash_bytes": (123).to_bytes(8, "big")})
    hexed = DF.from_row({"id": 11, "file_path": "hex.png", "size": 12, "width": 3, "height": 4, "phash_hex": "ff"})
    assert (blob.file_id, blob.phash, hexed.file_id, hexed.phash) == (10, 123, 11, 255)
    assert DF.from_row({"file_id": 1, "path": "p", "phash_u64": -1}).phash == U64
    assert DF.from_row({"file_id": 1, "path": "p", "phash": "0x10"}).phash == 16
    assert DF.from_row({"file_id": 1, "path": "p", "phash": np.int64(-2)}).phash == U64 - 1
    with pytest.raises(ValueError, match="missing perceptual hash"):
        DF.from_row({"file_id": 1, "path": "broken.jpg", "phash_hex": "not-a-hex-value"})
    with pytest.raises(ValueError, match="missing perceptual hash"):
        DF.from_row({"file_id": 1, "path": "x"})
    Cfg = kscanner.DuplicateScanConfig
    for bad in ({"band_bits": 0}, {"band_count": 0}, {"hamming_threshold": -1}, {"hamming_threshold": 65},
                {"cosine_threshold": 1.5}):
        with pytest.raises(ValueError):
            Cfg(**bad)
    with pytest.raises(AssertionError):
        kscanner.DuplicateScanner(Cfg(band_bits=16, band_count=5))


# ------------------------------------------------------------------ dup.refine / dup.cluster drop-ins


def test_refine_pair_decisions_and_reasons(tmp_path, monkeypatch):
    """tests/dup/test_refine.py:24-95 of the reference with SSIM stubbed by the oracle."""
    from PIL import Image, ImageEnhance

    monkeypatch.setattr(krefine, "_compute_ssim", lambda a, b: ref_py.compute_ssim(a, b))
    pa, pb = tmp_path / "a.png", tmp_path / "b.png"
    Image.new("RGB", (64, 64), color=(200, 10, 10)).save(pa)
    ImageEnhance.Brightness(Image.open(pa).convert("RGB")).enhance(1.02).save(pb)
    r = krefine.refine_pair(1, 2, pa, pb)
    assert r.is_duplicate and r.ssim > 0.95
    g, b = tmp_path / "g.png", tmp_path / "bl.png"
    Image.new("RGB", (64, 64), (0, 255, 0)).save(g)
    Image.new("RGB", (64, 64), (0, 0, 255)).save(b)
    r = krefine.refine_pair(1, 3, g, b, thresholds=krefine.RefinementThresholds(ssim=0.95, orb=0.5))
    assert not r.is_duplicate and r.reason == "below thresholds"
    broken = tmp_path / "broken.png"
    broken.write_bytes(b"not an image")
    assert krefine.refine_pair(1, 2, pa, broken) is None

    def boom(*_):
        raise RuntimeError("x")

    monkeypatch.setattr(krefine, "_compute_orb_ratio", boom)
    r = krefine.refine_pair(1, 2, pa, pb)
    assert r.is_duplicate and r.orb_ratio is None and r.reason == "ssim>=0.9"
    monkeypatch.setattr(krefine, "_compute_ssim", boom)
    r = krefine.refine_pair(1, 2, pa, pb)
    assert (r.is_duplicate, r.ssim, r.orb_ratio, r.reason) == (False, None, None, "ssim unavailable, orb unavailable")


def test_cluster_builder():
    M = krefine.RefinedMatch
    ms = [M(1, 2, 0.95, 0.2, True, "ssim"), M(2, 3, 0.93, 0.15, True, "ssim"), M(4, 5, 0.91, 0.16, True, "ssim"),
          M(3, 5, 0.5, 0.05, False, "below")]
    clusters = kcluster.ClusterBuilder().build(ms)
    assert [c.members for c in clusters] == [[1, 2, 3], [4, 5]]
    assert [c.representative for c in clusters] == [1, 4]
    assert len(clusters[0].matches) == 2 and kcluster.ClusterBuilder().build([]) == []
    assert [(r, m) for r, m in ref_py.cluster_matches([(m.file_id_a, m.file_id_b, m.is_duplicate) for m in ms])] == \
        [(c.representative, c.members) for c in clusters]


def test_cluster_pairs_host_union_find_matches_the_oracle():
    """ke_cluster_pairs_host (host code of the library, no GPU): components with the reference's "smaller root
    wins" representative (src/dup/cluster.py:22-70), for compact and for sparse id ranges."""
    from kobato_b200 import ops

    rng = np.random.default_rng(11)
    for lo, hi, n in ((0, 3000, 2000), (0, 70000, 1200), (-5, 5, 40), (10 ** 12, 10 ** 12 + 10 ** 9, 500), (0, 2, 1), (7, 8, 3)):
        a, b = rng.integers(lo, hi, n), rng.integers(lo, hi, n)
        want = ref_py.cluster_matches((int(x), int(y), True) for x, y in zip(a, b))
        assert ops.cluster_pairs(a, b) == want, (lo, hi, n)
    assert ops.cluster_pairs([], []) == []
    assert ops.cluster_pairs([5], [5]) == [(5, [5])]
    with pytest.raises(ValueError):
        ops.cluster_pairs([1, 2], [3])


def test_cluster_set_is_the_csr_view_of_the_same_components():
    from kobato_b200 import ops, pipeline

    rng = np.random.default_rng(4)
    a, b = rng.integers(0, 500, 300), rng.integers(0, 500, 300)
    keep = rng.random(300) < 0.7
    cs = pipeline._components(a.astype(np.int64), b.astype(np.int64), keep)
    want = ref_py.cluster_matches((int(x), int(y), bool(k)) for x, y, k in zip(a, b, keep))
    assert len(cs) == len(want) and cs.as_list() == want and cs[0] == want[0] and cs[-1] == want[-1]
    members, offsets = ops.cluster_pairs_csr(a[keep], b[keep])
    assert offsets[0] == 0 and offsets[-1] == len(members) and np.all(np.diff(offsets) >= 2)
    with pytest.raises(IndexError):
        cs[len(want)]
    empty = pipeline._components(a.astype(np.int64), b.astype(np.int64), np.zeros(300, bool))
    assert len(empty) == 0 and empty.as_list() == []
