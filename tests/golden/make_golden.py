"""Generate the golden vectors in this directory by running the LIVE reference.

Run in the build container only (needs /root/reference and its third-party deps):

    python tests/golden/make_golden.py

Writes
  phash_golden.json    sig.phash.phash / dhash (reference, PIL+cv2) on seeded synthetic images
  planes_golden.npz    PIL convert('L').resize(LANCZOS) 32x32 and 9x8 planes for a subset
  scanner_golden.json  dup.scanner.DuplicateScanner.build_clusters on seeded hash sets
  ssim_golden.json     oracle.ref_py SSIM restatement values (scikit-image is NOT installable
                       here, so these pin the restatement against drift, not the reference)
  n1_golden.json       ui.dup_refine_parallel.tile_ahash_bits / _load_small_gray / _mae01 (reference, PIL) on
                       seeded synthetic PNG files
plus the library versions they were produced with.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
sys.path.insert(0, "/root/reference/src")

from kobato_b200 import synth  # noqa: E402

U64 = (1 << 64) - 1

# (count, h, w, c) — c: 1 'L', 3 'RGB', 4 'RGBA'
IMAGE_CASES = [
    (32, 512, 512, 3),
    (4, 300, 200, 3),
    (4, 33, 47, 3),
    (4, 100, 33, 1),
    (4, 480, 640, 4),
    (2, 32, 32, 3),
    (2, 8, 9, 1),
    (2, 31, 29, 3),
    (2, 768, 1024, 3),
    (2, 16, 500, 3),
    (2, 256, 256, 1),
]
MODES = {1: "L", 3: "RGB", 4: "RGBA"}


def versions():
    import cv2
    import PIL
    import scipy

    return {"pillow": PIL.__version__, "opencv": cv2.__version__, "numpy": np.__version__, "scipy": scipy.__version__}


def gen_phash():
    from PIL import Image
    from sig.phash import dhash, phash

    cases = []
    planes32, planes98 = {}, {}
    for ci, (count, h, w, c) in enumerate(IMAGE_CASES):
        for k in range(count):
            idx = ci * 1000 + k
            arr = synth.synth_image(idx, h, w, c, n_set=1 << 30)
            im = Image.fromarray(arr, MODES[c])
            cases.append({"index": idx, "h": h, "w": w, "c": c, "phash": f"{phash(im) & U64:016x}",
                          "dhash": f"{dhash(im) & U64:016x}"})
            if k < 2:
                planes32[str(idx)] = np.asarray(im.convert("L").resize((32, 32), Image.Resampling.LANCZOS))
                planes98[str(idx)] = np.asarray(im.convert("L").resize((9, 8), Image.Resampling.LANCZOS))
    (HERE / "phash_golden.json").write_text(json.dumps({"versions": versions(), "seed": synth.SEED, "cases": cases}, indent=1))
    np.savez_compressed(HERE / "planes_golden.npz", **{f"p32_{k}": v for k, v in planes32.items()},
                        **{f"p98_{k}": v for k, v in planes98.items()})
    print("phash cases", len(cases))


def make_files(n, seed, ids_dupe=False):
    """Deterministic DuplicateFile-like metadata for n hashes."""
    h = synth.synth_hashes(n, seed=seed, planted=0.2, max_flips=12)
    rng = np.random.default_rng(seed)
    exts = ["jpg", "png", "webp", "gif", "bmp", "tiff", "jpeg"]
    files = []
    for i in range(n):
        fid = i + 1 if not (ids_dupe and i % 97 == 96) else i  # occasional repeated file_id
        files.append({"file_id": int(fid), "path": f"dir{int(rng.integers(0, 5))}/f{i:05d}.{exts[int(rng.integers(0, len(exts)))]}",
                      "size": int(rng.integers(0, 5) == 0 and rng.integers(0, 3) or rng.integers(1000, 2000000)),
                      "width": int(rng.integers(100, 4000)), "height": int(rng.integers(100, 4000)),
                      "phash": int(h[i])})
    return files


SCAN_CASES = [
    {"name": "t8_default", "n": 3000, "seed": 11, "cfg": {"hamming_threshold": 8}},
    {"name": "t4", "n": 3000, "seed": 12, "cfg": {"hamming_threshold": 4}},
    {"name": "t12", "n": 3000, "seed": 13, "cfg": {"hamming_threshold": 12}},
    {"name": "t0", "n": 2000, "seed": 14, "cfg": {"hamming_threshold": 0}},
    {"name": "t10_ratio", "n": 3000, "seed": 15, "cfg": {"hamming_threshold": 10, "size_ratio": 0.5}},
    {"name": "t8_bands8x8", "n": 2000, "seed": 16, "cfg": {"hamming_threshold": 8, "band_bits": 8, "band_count": 8}},
    {"name": "t8_bands12x5", "n": 2000, "seed": 17, "cfg": {"hamming_threshold": 8, "band_bits": 12, "band_count": 5}},
    {"name": "t8_paircap", "n": 3000, "seed": 18, "cfg": {"hamming_threshold": 8, "band_bits": 8, "band_count": 8}, "pair_cap": 60},
    {"name": "t8_dupe_ids", "n": 1500, "seed": 19, "cfg": {"hamming_threshold": 8}, "ids_dupe": True},
    {"name": "t64_small", "n": 300, "seed": 20, "cfg": {"hamming_threshold": 64}},
]


def gen_scanner():
    from dup.scanner import DuplicateFile, DuplicateScanConfig, DuplicateScanner

    out = []
    for case in SCAN_CASES:
        files = make_files(case["n"], case["seed"], case.get("ids_dupe", False))
        dfs = [DuplicateFile(file_id=f["file_id"], path=Path(f["path"]), size=f["size"], width=f["width"],
                             height=f["height"], phash=f["phash"]) for f in files]
        if "pair_cap" in case:
            os.environ["KE_DUP_BUCKET_PAIR_CAP"] = str(case["pair_cap"])
        else:
            os.environ.pop("KE_DUP_BUCKET_PAIR_CAP", None)
        clusters = DuplicateScanner(DuplicateScanConfig(**case["cfg"])).build_clusters(dfs)
        os.environ.pop("KE_DUP_BUCKET_PAIR_CAP", None)
        out.append({**case, "clusters": [{"keeper": c.keeper_id,
                                          "members": [[e.file.file_id, e.best_hamming] for e in c.files]} for c in clusters]})
        print(case["name"], "clusters", len(clusters))
    (HERE / "scanner_golden.json").write_text(json.dumps({"cases": out}, indent=None))


SSIM_CASES = [(0, 64, 64), (1, 256, 256), (2, 7, 7), (3, 8, 31), (4, 100, 37), (5, 512, 512)]


def ssim_pair(case_id, h, w):
    a = synth.synth_image(9000 + case_id, h, w, 1)
    kind = case_id % 3
    if kind == 0:
        b = np.clip(a.astype(np.int64) * 261 // 256 + 1, 0, 255).astype(np.uint8)
    elif kind == 1:
        b = synth.synth_image(9100 + case_id, h, w, 1)
    else:
        b = a.copy()
        b[::2, ::3] ^= 0x10
    return a, b


def gen_ssim():
    from oracle import ref_py

    out = []
    for cid, h, w in SSIM_CASES:
        a, b = ssim_pair(cid, h, w)
        out.append({"case": cid, "h": h, "w": w, "ssim": ref_py.ssim_of_planes(a, b)})
    # the reference's own behavioural pins (tests/dup/test_refine.py:24-46), through the restatement
    from PIL import Image, ImageEnhance

    base = Image.new("RGB", (64, 64), color=(200, 10, 10))
    var = ImageEnhance.Brightness(base).enhance(1.02)
    out.append({"case": "ref_brightness", "ssim": ref_py.compute_ssim(base, var)})
    out.append({"case": "ref_green_blue", "ssim": ref_py.compute_ssim(Image.new("RGB", (64, 64), (0, 255, 0)),
                                                                     Image.new("RGB", (64, 64), (0, 0, 255)))})
    (HERE / "ssim_golden.json").write_text(json.dumps({"versions": versions(), "note": "restatement, parity unpinned",
                                                       "cases": out}, indent=1))
    print("ssim", out)


# (h, w, c) of the files behind the N1 vectors, (grid, tile) of the tile aHash, thumb sizes of the pixel pass
N1_IMAGES = [(512, 512, 3), (300, 200, 3), (33, 47, 3), (100, 33, 1), (480, 640, 4), (64, 64, 3), (32, 32, 1), (31, 29, 3),
             (128, 128, 1), (700, 45, 3)]
N1_TILES = [(4, 8), (8, 8), (16, 16), (3, 5), (1, 7)]
N1_THUMBS = [128, 32, 7]


def gen_n1():
    import hashlib
    import tempfile

    from PIL import Image
    from ui import dup_refine_parallel as ref

    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for k, (h, w, c) in enumerate(N1_IMAGES):
            arr = synth.synth_image(1000 + k, h, w, c, n_set=1 << 30)
            path = Path(tmp) / f"n1_{k}.png"
            Image.fromarray(arr if c > 1 else arr[..., 0] if arr.ndim == 3 else arr, MODES[c]).save(path, format="PNG")
            paths.append(path)
            case = {"index": 1000 + k, "h": h, "w": w, "c": c, "tile_bits": {}, "small_gray_sha256": {}}
            for grid, tile in N1_TILES:
                v = ref.tile_ahash_bits(path, grid=grid, tile=tile)
                nbits = (grid * tile) ** 2
                # long bit strings are stored as the sha256 of their little-endian bytes
                case["tile_bits"][f"{grid}x{tile}"] = hex(v) if nbits <= 1024 else \
                    "sha256:" + hashlib.sha256(v.to_bytes((nbits + 7) // 8, "little")).hexdigest()
            for size in N1_THUMBS:
                plane = ref._load_small_gray(path, size=size)
                case["small_gray_sha256"][str(size)] = hashlib.sha256(plane.tobytes()).hexdigest()
            cases.append(case)
        mae = []
        for a, b in ((0, 1), (0, 5), (2, 3), (6, 8), (4, 9)):
            for size in (128, 32):
                mae.append({"a": a, "b": b, "size": size,
                            "mae": ref._mae01(ref._load_small_gray(paths[a], size), ref._load_small_gray(paths[b], size))})
    (HERE / "n1_golden.json").write_text(json.dumps({"versions": versions(), "cases": cases, "mae": mae}, indent=1))
    print("n1", len(cases), "files", len(mae), "mae pairs")


if __name__ == "__main__":
    gen_n1()
    gen_phash()
    gen_scanner()
    gen_ssim()
