"""CPU baseline runner — TEST/BENCH INFRASTRUCTURE ONLY (bench.py `--impl reference` and
`cpu_baseline`).

Times the reference's CPU path for one duplicate-scan step on a bounded sample:
  * pHash + dHash per image exactly as the reference's worker does after decoding
    (src/core/fastsig.py:24-37 -> src/sig/phash.py:33-57: PIL convert/resize + cv2.dct), fanned out
    over a process pool with all host cores like ``compute_signatures_mp`` (src/core/fastsig.py:65-99);
  * ``DuplicateScanner.build_clusters`` (src/dup/scanner.py:211-356) single-process, as it is
    single-threaded by design;
  * ``_compute_ssim`` (src/dup/refine.py:44-52) on the candidate pairs, process pool.
``/root/reference`` is not available on the GPU box, so these are the oracle restatements
(oracle/ref_py.py), which call the same Pillow / OpenCV / SciPy routines (kind "port").
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "kobato-eyes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

_POOL_IMAGES: np.ndarray | None = None  # inherited by forked workers (copy-on-write, no pickling)


def _gen(args):
    from kobato_b200 import synth

    i, h, w, c, n_set = args
    return synth.synth_image(i, h, w, c, n_set=n_set)


def _hash(idx: int):
    from PIL import Image

    from oracle import ref_py

    arr = _POOL_IMAGES[idx % len(_POOL_IMAGES)]
    im = Image.fromarray(arr, "RGB" if arr.ndim == 3 else "L")
    return ref_py.phash(im), ref_py.dhash(im)


def _ssim(pair):
    from PIL import Image

    from oracle import ref_py

    a = _POOL_IMAGES[pair[0] % len(_POOL_IMAGES)]
    b = _POOL_IMAGES[pair[1] % len(_POOL_IMAGES)]
    mode = "RGB" if a.ndim == 3 else "L"
    return ref_py.compute_ssim(Image.fromarray(a, mode), Image.fromarray(b, mode))


class CpuReference:
    """Holds a small pool of unique synthetic images and runs bounded dup-scan steps over it."""

    def __init__(self, h: int = 512, w: int = 512, c: int = 3, unique: int = 256, cores: int | None = None):
        global _POOL_IMAGES
        self.cores = cores or os.cpu_count() or 1
        self.h, self.w, self.c = h, w, c
        ctx = mp.get_context("fork")
        with ctx.Pool(self.cores) as pool:
            imgs = pool.map(_gen, [(i, h, w, c, unique) for i in range(unique)], chunksize=4)
        _POOL_IMAGES = np.stack(imgs)
        self.pool = ctx.Pool(self.cores)  # forked AFTER the images exist
        self.unique = unique

    def close(self):
        self.pool.close()
        self.pool.join()

    def step(self, sample: int, threshold: int = 8, ssim_threshold: float = 0.9, pairs_per_image: float = 3589 / 70000):
        """One bounded step; returns per-stage seconds and counts.

        The three stages keep the proportions of the GPU arm's step: every image of the sample is hashed, the LSH scan
        runs over a table of ``sample`` hashes (the pool's own hashes plus synthetic ones with the same 5 % of planted
        near-duplicates), and ``round(sample * pairs_per_image)`` pairs are SSIM-verified — the candidate rate the GPU
        step sees on the 70 000-image set (3 589 pairs per 70 000 images); the pool's own edges first, then further
        pairs of pool images (SSIM cost does not depend on the content)."""
        from kobato_b200 import synth
        from oracle import ref_py

        t0 = time.perf_counter()
        sigs = self.pool.map(_hash, range(sample), chunksize=16)
        t1 = time.perf_counter()
        # distinct ids; images cycle through the unique pool, so beyond the pool the table is filled with synthetic
        # hashes (exact repeats of the pool's images would otherwise make every bucket a giant cluster)
        first = min(sample, self.unique)
        table = [sigs[i][0] & ref_py.U64 for i in range(first)]
        if sample > first:
            table += [int(x) for x in synth.synth_hashes(sample - first, seed=synth.SEED + 17, planted=0.05)]
        files = [ref_py.FileRec(i + 1, f"f{i}.png", 1000 + i, self.w, self.h, table[i]) for i in range(sample)]
        edges = ref_py.scan_edges(files, hamming_threshold=threshold)
        clusters = ref_py.build_clusters(files, hamming_threshold=threshold)
        t2 = time.perf_counter()
        want = max(1, int(round(sample * pairs_per_image)))
        pairs = [(a - 1, b - 1) for (a, b) in edges if a <= first and b <= first][:want]
        k = 0
        while len(pairs) < want:  # fill up with further pool pairs
            pairs.append((k % self.unique, (k * 7 + 1) % self.unique))
            k += 1
        scores = self.pool.map(_ssim, pairs, chunksize=2) if pairs else []
        t3 = time.perf_counter()
        return {
            "hash_s": t1 - t0, "scan_s": t2 - t1, "ssim_s": t3 - t2, "total_s": t3 - t0,
            "images": sample, "scan_files": sample, "edges": len(edges), "clusters": len(clusters),
            "ssim_pairs": len(pairs), "accepted": int(sum(s >= ssim_threshold for s in scores)),
        }
