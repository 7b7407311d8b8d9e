"""N1: tile-aHash / pixel-MAE refinement, reference src/ui/dup_refine_parallel.py.

CPU part: the oracle restatement is pinned against the live reference (when mounted) and the golden vectors;
the drop-in's host logic (progress, cancel, failure summaries, thresholds) is exercised with the four GPU ops
replaced by the oracle.  GPU part (-m gpu): the CUDA kernels against Pillow / the oracle bit for bit, and the
scenarios of the reference's own tests (tests/dup/test_dup_refine_parallel.py) through the drop-in.
"""
from __future__ import annotations

import hashlib
import json
import logging
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import pytest
from conftest import GOLDEN, REFERENCE_SRC
from PIL import Image

from kobato_b200 import ops, synth
from kobato_b200.ui import dup_refine_parallel as krp
from oracle import ref_py

MODES = {1: "L", 3: "RGB", 4: "RGBA"}


@pytest.fixture(scope="session")
def golden_n1():
    return json.loads((GOLDEN / "n1_golden.json").read_text())


def _pil(case) -> Image.Image:
    arr = synth.synth_image(case["index"], case["h"], case["w"], case["c"], n_set=1 << 30)
    return Image.fromarray(arr[..., 0] if arr.ndim == 3 and case["c"] == 1 else arr, MODES[case["c"]])


def _bits_key(value: int, grid: int, tile: int) -> str:
    nbits = (grid * tile) ** 2
    return hex(value) if nbits <= 1024 else "sha256:" + hashlib.sha256(value.to_bytes((nbits + 7) // 8, "little")).hexdigest()


# ------------------------------------------------------------------ file builders for the scenario tests


@dataclass
class FileStub:
    file_id: int
    path: Path


@dataclass
class EntryStub:
    file: FileStub
    best_hamming: int | None = None


@dataclass
class ClusterStub:
    files: list
    keeper_id: int


def _entry(path, fid):
    return EntryStub(FileStub(fid, path))


def _half_bright(path: Path, size: int = 32, light: int = 230, dark: int = 20):
    """Top half bright, bottom half dark, plus a 2-pixel bright column at the left (not rotation symmetric)."""
    a = np.full((size, size), dark, np.uint8)
    a[: size // 2, :] = light
    a[:, :2] = light
    Image.fromarray(a, "L").save(path, format="PNG")


def _with_patch(path: Path, source: Path, value: int = 255):
    """Copy of `source` with the top-left quarter of its bottom-right 8x8 tile set to `value`."""
    with Image.open(source) as im:
        a = np.array(im)
    h, w = a.shape
    t = w // 4
    a[h - t: h - t // 2, w - t: w - t // 2] = value
    Image.fromarray(a, "L").save(path, format="PNG")


def _rot90(path: Path, source: Path):
    with Image.open(source) as im:
        im.transpose(Image.Transpose.ROTATE_90).save(path, format="PNG")


def _clone(path: Path, source: Path):
    with Image.open(source) as im:
        im.copy().save(path, format="PNG")


# ------------------------------------------------------------------ CPU: oracle pinned


def test_oracle_n1_matches_golden(golden_n1):
    planes = {}
    for k, case in enumerate(golden_n1["cases"]):
        img = _pil(case)
        for key, want in case["tile_bits"].items():
            grid, tile = map(int, key.split("x"))
            assert _bits_key(ref_py.tile_ahash_bits_image(img, grid, tile), grid, tile) == want, (case, key)
        for size, want in case["small_gray_sha256"].items():
            plane = ref_py.small_gray(img, int(size))
            planes[(k, int(size))] = plane
            assert hashlib.sha256(plane.tobytes()).hexdigest() == want, (case, size)
    for m in golden_n1["mae"]:
        assert ref_py.mae01(planes[(m["a"], m["size"])], planes[(m["b"], m["size"])]) == m["mae"]


@pytest.mark.reference
def test_oracle_n1_matches_live_reference(tmp_path):
    import sys

    sys.path.append(str(REFERENCE_SRC))
    from ui import dup_refine_parallel as ref

    rng = np.random.default_rng(5)
    for k, (h, w, c) in enumerate(((64, 64, 3), (37, 91, 1), (200, 120, 4), (16, 16, 3))):
        arr = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        img = Image.fromarray(arr[..., 0] if c == 1 else arr, MODES[c])
        path = tmp_path / f"r{k}.png"
        img.save(path, format="PNG")
        for grid, tile in ((4, 8), (8, 8), (2, 3)):
            assert ref.tile_ahash_bits(path, grid=grid, tile=tile) == ref_py.tile_ahash_bits_image(img, grid, tile)
        assert np.array_equal(ref._load_small_gray(path, 48), ref_py.small_gray(img, 48))
    a, b = rng.integers(0, 256, (2, 32, 32), dtype=np.uint8)
    assert ref._mae01(a, b) == ref_py.mae01(a, b)
    assert ref.tile_hamming(0b1011, 0b0110) == ref_py.tile_hamming(0b1011, 0b0110) == 3


# ------------------------------------------------------------------ CPU: host logic with the GPU ops stubbed by the oracle


@pytest.fixture
def oracle_ops(monkeypatch):
    import torch

    def gray_resize_batch(images, out_w, out_h, filter="bilinear"):
        arr = np.asarray(images)
        flt = Image.Resampling.BILINEAR if filter == "bilinear" else Image.Resampling.LANCZOS
        out = []
        for a in arr:
            im = Image.fromarray(a[..., 0] if a.ndim == 3 and a.shape[2] == 1 else a)
            out.append(np.asarray(im.convert("L").resize((out_w, out_h), flt), np.uint8))
        return torch.from_numpy(np.stack(out))

    def tile_ahash_bits(planes, grid=4, tile=8):
        rows = []
        for p in planes.numpy():
            a = p.reshape(grid, tile, grid, tile).transpose(0, 2, 1, 3)
            bits = (a > a.mean(axis=(2, 3), keepdims=True)).reshape(-1).astype(np.uint8)
            packed = np.packbits(bits, bitorder="little")
            packed = np.pad(packed, (0, (-len(packed)) % 4))
            rows.append(packed.view("<i4"))
        return torch.from_numpy(np.stack(rows))

    def bits_hamming_pairs(bits, ia, ib):
        b = bits.numpy().view(np.uint32)
        return torch.from_numpy(np.array([int(np.bitwise_count(b[i] ^ b[j]).sum()) for i, j in zip(ia, ib)], np.int32))

    def plane_sad_pairs(planes, ia, ib):
        p = planes.numpy().astype(np.int64)
        return torch.from_numpy(np.array([int(np.abs(p[i] - p[j]).sum()) for i, j in zip(ia, ib)], np.int64))

    for name, fn in (("gray_resize_batch", gray_resize_batch), ("tile_ahash_bits", tile_ahash_bits),
                     ("bits_hamming_pairs", bits_hamming_pairs), ("plane_sad_pairs", plane_sad_pairs)):
        monkeypatch.setattr(ops, name, fn)


def _scenarios(tmp_path, caplog):
    """The behaviour the reference's tests/dup/test_dup_refine_parallel.py:96-330 pin, through the drop-in."""
    base, clone, rotated, variant = (tmp_path / n for n in ("base.png", "clone.png", "rotated.png", "variant.png"))
    _half_bright(base)
    _clone(clone, base)
    _rot90(rotated, base)
    _with_patch(variant, base)

    sig = {p: krp.tile_ahash_bits(p) for p in (base, clone, rotated, variant)}
    assert sig[base] == sig[clone]
    d_rot, d_var = krp.tile_hamming(sig[base], sig[rotated]), krp.tile_hamming(sig[base], sig[variant])
    assert 0 < d_var < d_rot
    for p in (base, rotated, variant):  # the drop-in's value IS the reference function's value
        with Image.open(p) as im:
            assert sig[p] == ref_py.tile_ahash_bits_image(im)

    # tile-hash refinement: thresholds, both progress phases, input untouched
    cluster = ClusterStub([_entry(base, 1), _entry(clone, 2), _entry(rotated, 3)], keeper_id=1)
    broad = krp.refine_by_tilehash_parallel([cluster], max_bits=64)
    assert len(broad) == 1 and {e.file.file_id for e in broad[0].files} == {1, 2, 3}
    ticks = []
    narrow = krp.refine_by_tilehash_parallel([cluster], max_bits=0, tick=lambda done, total, phase: ticks.append((phase, done, total)))
    assert len(narrow) == 1 and {e.file.file_id for e in narrow[0].files} == {1, 2}
    assert isinstance(narrow[0], ClusterStub) and narrow[0].keeper_id == 1
    assert {t[0] for t in ticks} == {1, 2} and ticks[-1][1:] == (1, 1)
    assert (1, 3, 3) in ticks  # phase 1 ends at done == total
    assert krp.refine_by_tilehash_parallel([cluster], is_cancelled=lambda: True) == []
    # a keeper that is not in the file list, and a cluster with a single survivor, both vanish
    assert krp.refine_by_tilehash_parallel([ClusterStub([_entry(base, 1), _entry(clone, 2)], keeper_id=99)]) == []
    assert krp.refine_by_tilehash_parallel([ClusterStub([_entry(base, 1), _entry(rotated, 3)], keeper_id=1)], max_bits=0) == []

    # failures are summarised in one warning with the exception type and a sample path
    missing = tmp_path / "missing.png"
    caplog.clear()
    caplog.set_level(logging.WARNING, logger="ui.dup_refine")
    assert krp.refine_by_tilehash_parallel([ClusterStub([_entry(missing, 1)], keeper_id=1)]) == []
    msgs = [r.getMessage() for r in caplog.records if r.levelno == logging.WARNING]
    assert msgs and "TileHash phase1" in msgs[0] and "FileNotFoundError" in msgs[0] and str(missing) in msgs[0]

    # pixel refinement: duplicates stay, a black image and a missing file go; one tick for one cluster
    black = tmp_path / "black.png"
    Image.new("L", (32, 32), 0).save(black, format="PNG")
    cl = ClusterStub([_entry(base, 1), _entry(clone, 2), _entry(black, 3), _entry(missing, 4)], keeper_id=1)
    pticks = []
    kept = krp.refine_by_pixels_parallel([cl], mae_thr=0.001, tick=lambda d, t: pticks.append((d, t)))
    assert len(kept) == 1 and {e.file.file_id for e in kept[0].files} == {1, 2} and pticks == [(1, 1)]

    # mixed aspect ratios are squashed to the thumb size, not letter-boxed
    wide, tall, dark = tmp_path / "wide.png", tmp_path / "tall.png", tmp_path / "dark.png"
    Image.new("L", (64, 32), 120).save(wide, format="PNG")
    Image.new("L", (32, 64), 120).save(tall, format="PNG")
    Image.new("L", (32, 64), 20).save(dark, format="PNG")
    kept = krp.refine_by_pixels_parallel([ClusterStub([_entry(wide, 1), _entry(tall, 2), _entry(dark, 3)], keeper_id=1)],
                                         mae_thr=0.001, thumb_size=32)
    assert len(kept) == 1 and {e.file.file_id for e in kept[0].files} == {1, 2}

    # the edges of a wide image count
    wb, we = tmp_path / "wide_base.png", tmp_path / "wide_edges.png"
    a = np.full((32, 96), 120, np.uint8)
    Image.fromarray(a, "L").save(wb, format="PNG")
    a[:, :8] = 255
    a[:, -8:] = 255
    Image.fromarray(a, "L").save(we, format="PNG")
    assert krp.refine_by_pixels_parallel([ClusterStub([_entry(wb, 1), _entry(we, 2)], keeper_id=1)], mae_thr=0.001,
                                         thumb_size=32, workers=1) == []

    # keeper / member load failures: cluster dropped / member dropped, both logged with the path
    mk, mm = tmp_path / "missing_keeper.png", tmp_path / "missing_member.png"
    caplog.clear()
    kept = krp.refine_by_pixels_parallel([ClusterStub([_entry(mk, 10), _entry(clone, 11)], keeper_id=10),
                                          ClusterStub([_entry(base, 20), _entry(clone, 21), _entry(mm, 22)], keeper_id=20)],
                                         mae_thr=0.001)
    assert len(kept) == 1 and {e.file.file_id for e in kept[0].files} == {20, 21}
    msgs = [r.getMessage() for r in caplog.records if r.levelno == logging.WARNING]
    assert any("keeper load errors" in m and str(mk) in m for m in msgs)
    assert any("image load errors" in m and str(mm) in m for m in msgs)

    # cancel: nothing returned, no tick
    calls = []
    assert krp.refine_by_pixels_parallel([cl, cl], is_cancelled=lambda: True, tick=lambda d, t: calls.append(1)) == []
    assert not calls

    # tile hash then pixels, with a corrupt member in the cluster
    broken = tmp_path / "broken.png"
    broken.write_bytes(b"broken image")
    cl = ClusterStub([_entry(base, 1), _entry(clone, 2), _entry(black, 3), _entry(broken, 4)], keeper_id=1)
    t_ref = krp.refine_by_tilehash_parallel([cl], max_bits=0, io_workers=1)
    p_ref = krp.refine_by_pixels_parallel(t_ref, mae_thr=0.001, workers=1)
    assert len(t_ref) == 1 and {e.file.file_id for e in t_ref[0].files} == {1, 2}
    assert len(p_ref) == 1 and {e.file.file_id for e in p_ref[0].files} == {1, 2}


def test_dropin_host_logic_with_oracle_ops(tmp_path, caplog, oracle_ops):
    _scenarios(tmp_path, caplog)


@pytest.mark.reference
def test_dropin_decisions_equal_the_live_reference(tmp_path, oracle_ops):
    """Random clusters of synthetic files: same surviving members as the reference for several thresholds."""
    import sys

    sys.path.append(str(REFERENCE_SRC))
    from ui import dup_refine_parallel as ref

    paths = []
    for k in range(24):
        arr = synth.synth_image(k, 96, 128, 3, n_set=24, planted=0.5)
        p = tmp_path / f"s{k}.png"
        Image.fromarray(arr, "RGB").save(p, format="PNG")
        paths.append(p)
    rng = np.random.default_rng(2)
    clusters = []
    for c in range(8):
        members = rng.choice(24, size=5, replace=False).tolist()
        clusters.append(ClusterStub([_entry(paths[m], 100 * c + m) for m in members], keeper_id=100 * c + members[0]))

    def ids(cls):
        return sorted(sorted(e.file.file_id for e in cl.files) for cl in cls)

    for grid, tile, max_bits in ((4, 8, 32), (8, 8, 200), (4, 8, 0), (8, 8, 4096)):
        assert ids(krp.refine_by_tilehash_parallel(clusters, grid, tile, max_bits)) == \
            ids(ref.refine_by_tilehash_parallel(clusters, grid, tile, max_bits))
    for thr, size in ((0.006, 128), (0.02, 32), (0.0, 64), (1.0, 16)):
        assert ids(krp.refine_by_pixels_parallel(clusters, mae_thr=thr, thumb_size=size)) == \
            ids(ref.refine_by_pixels_parallel(clusters, mae_thr=thr, thumb_size=size))


# ------------------------------------------------------------------ GPU: kernels against Pillow / the oracle


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(512, 512, 3), (300, 200, 3), (33, 47, 3), (100, 33, 1), (480, 640, 4), (32, 32, 1), (31, 29, 3),
                                   (128, 128, 1), (700, 45, 3), (5, 3, 3), (256, 512, 4), (512, 256, 3), (96, 160, 3), (300, 256, 3),
                                   (130, 144, 1), (48, 80, 4), (512, 512, 1)])
def test_gray_resize_is_byte_identical_to_pillow(shape):
    h, w, c = shape
    imgs = synth.synth_images(0, 5, h, w, c, n_set=5)
    for (ow, oh), flt in (((128, 128), "bilinear"), ((32, 32), "bilinear"), ((64, 64), "bilinear"), ((w, 17), "bilinear"),
                          ((64, 32), "bilinear"), ((24, 48), "bilinear"), ((64, 64), "lanczos"), ((128, 16), "lanczos"),
                          ((9, h), "bilinear"), ((32, 32), "lanczos"), ((9, 8), "lanczos"), ((w, h), "bilinear"),
                          ((2 * w + 1, 3 * h), "bilinear")):
        got = ops.gray_resize_batch(imgs, ow, oh, flt).cpu().numpy()
        res = Image.Resampling.BILINEAR if flt == "bilinear" else Image.Resampling.LANCZOS
        for k in range(imgs.shape[0]):
            im = Image.fromarray(imgs[k][..., 0] if c == 1 and imgs[k].ndim == 3 else imgs[k], MODES[c])
            want = np.asarray(im.convert("L").resize((ow, oh), res), np.uint8)
            assert np.array_equal(got[k], want), (shape, ow, oh, flt, k)


@pytest.mark.gpu
def test_tile_bits_hamming_and_sad_match_oracle(golden_n1):
    import torch

    imgs = [_pil(case) for case in golden_n1["cases"]]
    arrays = [np.asarray(im, np.uint8) for im in imgs]
    for grid, tile in ((4, 8), (8, 8), (16, 16), (3, 5), (1, 7)):
        bits = krp.tile_ahash_bits_many(arrays, grid, tile)
        got = ops.bits_to_ints(bits)
        for case, value in zip(golden_n1["cases"], got):
            assert _bits_key(value, grid, tile) == case["tile_bits"][f"{grid}x{tile}"], (case["index"], grid, tile)
        n = len(got)
        ia, ib = np.meshgrid(np.arange(n), np.arange(n))
        d = ops.bits_hamming_pairs(bits, ia.ravel(), ib.ravel()).cpu().numpy()
        want = [ref_py.tile_hamming(got[i], got[j]) for i, j in zip(ia.ravel(), ib.ravel())]
        assert d.tolist() == want
    for size in (128, 32, 7):
        planes = krp._resize_groups(arrays, size)
        host = planes.cpu().numpy()
        for k, case in enumerate(golden_n1["cases"]):
            assert hashlib.sha256(host[k].tobytes()).hexdigest() == case["small_gray_sha256"][str(size)]
        pairs = [(m["a"], m["b"]) for m in golden_n1["mae"] if m["size"] == size]
        if pairs:
            sad = ops.plane_sad_pairs(planes, [a for a, _ in pairs], [b for _, b in pairs]).cpu().numpy()
            for (a, b), s, m in zip(pairs, sad, [m for m in golden_n1["mae"] if m["size"] == size]):
                assert (int(s) / float(size * size)) / 255.0 == m["mae"] == ref_py.mae01(host[a], host[b])
    with pytest.raises(ValueError):
        ops.bits_hamming_pairs(bits, [0], [len(imgs)])
    with pytest.raises(ValueError):
        ops.tile_ahash_bits(torch.zeros((2, 31, 32), dtype=torch.uint8, device="cuda"), 4, 8)
    assert ops.bits_hamming_pairs(bits, [], []).numel() == 0 and ops.gray_resize_batch(np.zeros((0, 8, 8, 3), np.uint8), 4, 4).shape == (0, 4, 4)


@pytest.mark.gpu
def test_reference_scenarios_on_gpu(tmp_path, caplog):
    _scenarios(tmp_path, caplog)


@pytest.mark.gpu
def test_n1_batch_properties_at_scale():
    """UI-scale batch (8x8 tiles -> 4096 bits): identical planes -> distance 0 / SAD 0, symmetry, and a sampled oracle check."""
    n = 2048
    bank = ops.synth_images_device(0, n, 256, 256, 3, n_set=n, planted=0.3)
    planes = ops.gray_resize_batch(bank, 64, 64, "bilinear")
    bits = ops.tile_ahash_bits(planes, 8, 8)
    rng = np.random.default_rng(3)
    ia, ib = rng.integers(0, n, 20000), rng.integers(0, n, 20000)
    d_ab, d_ba = ops.bits_hamming_pairs(bits, ia, ib), ops.bits_hamming_pairs(bits, ib, ia)
    s_ab, s_ba = ops.plane_sad_pairs(planes, ia, ib), ops.plane_sad_pairs(planes, ib, ia)
    assert bool((d_ab == d_ba).all()) and bool((s_ab == s_ba).all())
    same = ops.bits_hamming_pairs(bits, ia, ia)
    assert int(same.abs().sum()) == 0 and int(ops.plane_sad_pairs(planes, ib, ib).abs().sum()) == 0
    host_img, host_planes = bank[:64].cpu().numpy(), planes.cpu().numpy()
    ints = ops.bits_to_ints(bits[:64])
    for k in range(0, 64, 7):
        im = Image.fromarray(host_img[k], "RGB")
        assert ints[k] == ref_py.tile_ahash_bits_image(im, 8, 8)
        assert np.array_equal(host_planes[k], ref_py.small_gray(im, 64))
    for q in range(0, 20000, 1999):
        assert int(s_ab[q]) == int(np.abs(host_planes[ia[q]].astype(np.int64) - host_planes[ib[q]].astype(np.int64)).sum())
