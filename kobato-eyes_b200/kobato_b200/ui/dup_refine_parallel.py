"""Drop-in for the reference's ``ui.dup_refine_parallel`` (src/ui/dup_refine_parallel.py:59-340), the
refinement the shipped UI runs on the clusters of a scan.

Same public names, arguments, progress / cancel / failure-logging conventions:

* ``tile_ahash_bits(path, grid, tile)`` / ``tile_hamming(a, b)``                       (:59-88)
* ``refine_by_tilehash_parallel(clusters, grid, tile, max_bits, io_workers, tick, is_cancelled)``  (:110-200):
  phase 1 signs every distinct file (tick every 64 and at the end, failures summarised in ONE warning on the
  ``ui.dup_refine`` logger with ``"<ExcType>: <message>"`` keys), phase 2 keeps the members within
  ``max_bits`` of the keeper and drops clusters with fewer than two survivors (tick every 16 / at the end).
* ``_load_small_gray`` / ``_mae01`` / ``refine_by_pixels_parallel(clusters, mae_thr, thumb_size, workers,
  tick, is_cancelled)`` (:203-340): keeper load failure drops the cluster, member load failure drops the
  member, both summarised in warnings; cancel returns ``[]``.

What changed is where the arithmetic runs: host threads only open + EXIF-transpose the files (Pillow
releases the GIL); ``convert("L").resize(BILINEAR)`` (``ke_gray_resize_batch``), the tile-mean bits
(``ke_tile_ahash_bits``), the keeper/member Hamming distances (``ke_bits_hamming_pairs``) and the pixel
SADs (``ke_plane_sad_pairs``) are batched CUDA launches, one per image geometry.  Output clusters come
back in input order (the reference's pixel pass returns them in thread-completion order).
"""
from __future__ import annotations

import logging
import os
from collections import Counter
from collections.abc import Callable, Sequence
from concurrent.futures import ThreadPoolExecutor, as_completed
from pathlib import Path
from typing import Any

import numpy as np
from PIL import Image, ImageOps

from .. import ops
from ..sig.phash import _decoded_array

log = logging.getLogger("ui.dup_refine")

_GPU_WINDOW = 512  # at most this many decoded images resident on the host between GPU launches ...
_GPU_WINDOW_BYTES = int(os.environ.get("KE_REFINE_WINDOW_MB", "512")) << 20  # ... and at most this many bytes of them


def _rebuild_cluster_like(cluster, files):
    """A cluster of the same runtime type with a narrowed file list (reference :52-56)."""
    return type(cluster)(files=list(files), keeper_id=cluster.keeper_id)


def _norm_path(p) -> Path:
    try:
        return Path(p).resolve(strict=False)
    except Exception:
        return Path(os.path.normcase(os.path.abspath(str(p))))


def _format_failure_summary(counts: Counter, samples: dict) -> str:
    parts = []
    for err, count in counts.items():
        sample = samples.get(err)
        parts.append(f"{count}×{err}" if sample is None else f"{count}×{err} (例: {sample})")
    return "; ".join(parts)


def _open_decoded(path) -> np.ndarray:
    """Image.open + exif_transpose (reference :64-66 / :204-205) -> the uint8 array the kernels consume
    ('L', 'RGB', 'RGBA' as they are, every other mode through the reference's own convert("L"))."""
    with Image.open(path) as opened:
        return _decoded_array(ImageOps.exif_transpose(opened))


def _resize_groups(arrays: Sequence[np.ndarray], size: int, failed: dict | None = None):
    """[decoded arrays of any geometry] -> CUDA uint8 [n, size, size] planes (input order).

    A geometry group whose launch fails is retried image by image; with ``failed`` given, the images that still fail are
    recorded there as ``{position: "<ExcType>: <message>"}`` (their planes stay zero) instead of raising — the reference
    counts a failing file and carries on (src/ui/dup_refine_parallel.py:150-163)."""
    import torch

    out = None
    groups: dict[tuple, list[int]] = {}
    for k, a in enumerate(arrays):
        groups.setdefault(a.shape, []).append(k)

    def place(members, planes):
        nonlocal out
        if out is None:
            out = torch.zeros((len(arrays), size, size), dtype=torch.uint8, device=planes.device)
        out[torch.as_tensor(members, device=planes.device)] = planes

    for members in groups.values():
        try:
            place(members, ops.gray_resize_batch(np.stack([arrays[k] for k in members]), size, size, "bilinear"))
        except Exception:
            if failed is None:
                raise
            for k in members:
                try:
                    place([k], ops.gray_resize_batch(arrays[k][None], size, size, "bilinear"))
                except Exception as exc:
                    failed[k] = f"{type(exc).__name__}: {exc}"
    return out


def _windows(paths: Sequence):
    """Slices of ``paths`` for one decode + GPU round: at most _GPU_WINDOW files each (the byte bound is applied while
    decoding, see _decode_window)."""
    for lo in range(0, len(paths), _GPU_WINDOW):
        yield paths[lo:lo + _GPU_WINDOW]


def tile_ahash_bits_many(arrays: Sequence[np.ndarray], grid: int = 4, tile: int = 8):
    """Decoded arrays -> int32 CUDA tensor [n, words] of tile-aHash bits."""
    return ops.tile_ahash_bits(_resize_groups(arrays, grid * tile), grid, tile)


def tile_ahash_bits(path: Path, grid: int = 4, tile: int = 8) -> int:
    """(grid*tile)^2-bit tile aHash of one file as a little-endian packed integer (reference :59-83)."""
    return ops.bits_to_ints(tile_ahash_bits_many([_open_decoded(path)], grid, tile))[0]


def tile_hamming(a_bits: int, b_bits: int) -> int:
    return (a_bits ^ b_bits).bit_count()


def _decode_all(paths, workers, is_cancelled, on_done, flush=None):
    """Decode `paths` with a thread pool -> ({path: array}, {path: "<ExcType>: <message>"} in completion order),
    or None when cancelled.  `on_done()` is called once per finished file (progress).

    Host memory is bounded by BYTES, not by count (multi-megapixel photographs are ~50 MB each decoded): files are
    submitted a few per worker at a time and, with ``flush`` given, ``flush(decoded)`` is called — and the arrays
    dropped — whenever more than _GPU_WINDOW_BYTES of decoded pixels are resident."""
    decoded: dict[Any, np.ndarray] = {}
    errors: dict[Any, str] = {}
    held = 0
    step = max(1, workers) * 2
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        for lo in range(0, len(paths), step):
            futs = {ex.submit(_open_decoded, p): p for p in paths[lo:lo + step]}
            for f in as_completed(futs):
                if is_cancelled and is_cancelled():
                    for ff in futs:
                        ff.cancel()
                    return None
                p = futs[f]
                try:
                    decoded[p] = f.result()
                    held += decoded[p].nbytes
                except Exception as exc:
                    errors[p] = f"{type(exc).__name__}: {exc}"
                on_done()
            if flush is not None and held > _GPU_WINDOW_BYTES:
                flush(decoded)
                decoded = {}
                held = 0
    if flush is not None:
        if decoded:
            flush(decoded)
        return {}, errors
    return decoded, errors


def refine_by_tilehash_parallel(clusters, grid: int = 4, tile: int = 8, max_bits: int = 32, io_workers=None, tick=None,
                                is_cancelled: Callable[[], bool] | None = None):
    if is_cancelled and is_cancelled():
        return []

    # --- phase 1: signatures of every distinct file
    all_paths = [_norm_path(e.file.path) for cl in clusters for e in cl.files]
    uniq_paths = sorted(set(all_paths), key=lambda p: (p.anchor, str(p.parent)))
    total1 = len(uniq_paths)
    if io_workers is None:
        io_workers = int(os.environ.get("KE_TILEHASH_THREADS", "0")) or min(8, (os.cpu_count() or 4) * 2)
    log.info("TileHash phase1: %d files, threads=%d", total1, io_workers)

    done = [0]

    def _on_done():
        done[0] += 1
        if tick and (done[0] % 64 == 0 or done[0] == total1):
            tick(done[0], total1, phase=1)

    row_of: dict[Path, int] = {}
    chunks = []
    fail_counts: Counter = Counter()
    fail_samples: dict = {}
    def _sign(decoded):  # one GPU round over the decoded files of a window (path order kept)
        ok = [p for p in decoded_order if p in decoded]
        if not ok:
            return
        failed: dict[int, str] = {}
        planes = _resize_groups([decoded[p] for p in ok], grid * tile, failed)
        for k, key in failed.items():  # a GPU-stage failure is a skipped file, like a decode failure
            fail_counts[key] += 1
            fail_samples.setdefault(key, ok[k])
        good = [k for k in range(len(ok)) if k not in failed]
        if not good:
            return
        import torch

        bits = ops.tile_ahash_bits(planes[torch.as_tensor(good, device=planes.device)], grid, tile)
        base = sum(c.shape[0] for c in chunks)
        chunks.append(bits)
        for row, k in enumerate(good):
            row_of[ok[k]] = base + row

    for window in _windows(uniq_paths):
        decoded_order = window
        res = _decode_all(window, io_workers, is_cancelled, _on_done, flush=_sign)
        if res is None:
            return []
        for path, key in res[1].items():
            fail_counts[key] += 1
            fail_samples.setdefault(key, path)
    if fail_counts:
        log.warning("TileHash phase1 skipped %d file(s) due to errors: %s", sum(fail_counts.values()),
                    _format_failure_summary(fail_counts, fail_samples))

    # --- phase 2: keeper vs member distances, one launch for every cluster
    plan = []  # (cluster, [(entry, pair index)])
    ia, ib = [], []
    for cl in clusters:
        keep = next((e for e in cl.files if e.file.file_id == cl.keeper_id), None)
        base = row_of.get(_norm_path(keep.file.path)) if keep else None
        members = []
        if base is not None:
            for e in cl.files:
                row = row_of.get(_norm_path(e.file.path))
                if row is not None:
                    members.append((e, len(ia)))
                    ia.append(base)
                    ib.append(row)
        plan.append((cl, members if base is not None else None))
    dist = []
    if ia:
        import torch

        dist = ops.bits_hamming_pairs(torch.cat(chunks), ia, ib).cpu().tolist()

    out = []
    total2 = len(clusters)
    for i, (cl, members) in enumerate(plan, 1):
        if is_cancelled and is_cancelled():
            return []
        if members is not None:
            oks = [e for e, q in members if dist[q] <= max_bits]
            if len(oks) >= 2:
                out.append(_rebuild_cluster_like(cl, oks))
        if tick and (i % 16 == 0 or i == total2):
            tick(i, total2, phase=2)
    return out


def _load_small_gray(path: Path, size: int = 128) -> np.ndarray:
    """convert("L").resize((size, size), BILINEAR) of one file as a uint8 array (reference :203-207)."""
    return _resize_groups([_open_decoded(path)], size)[0].cpu().numpy()


def _mae01(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.mean(np.abs(a.astype(np.int16) - b.astype(np.int16))) / 255.0)


def refine_by_pixels_parallel(clusters, mae_thr: float = 0.006, thumb_size: int = 128, workers=None, tick=None,
                              is_cancelled: Callable[[], bool] | None = None):
    total = len(clusters)
    if is_cancelled and is_cancelled():
        return []
    worker_count = workers if workers is not None else min(8, (os.cpu_count() or 4))

    # every file any cluster needs, decoded once (the reference re-opens the keeper per cluster)
    paths = []
    seen = set()
    for cl in clusters:
        for e in cl.files:
            if e.file.path not in seen:
                seen.add(e.file.path)
                paths.append(e.file.path)
    row_of: dict[Any, int] = {}
    errors: dict[Any, str] = {}
    chunks = []
    def _thumb(decoded):  # one GPU round over the decoded files of a window (path order kept)
        ok = [p for p in decoded_order if p in decoded]
        if not ok:
            return
        failed: dict[int, str] = {}
        planes = _resize_groups([decoded[p] for p in ok], thumb_size, failed)
        for k, key in failed.items():  # a GPU-stage failure counts like a load failure of that file
            errors[ok[k]] = key
        good = [k for k in range(len(ok)) if k not in failed]
        if not good:
            return
        import torch

        base = sum(c.shape[0] for c in chunks)
        chunks.append(planes[torch.as_tensor(good, device=planes.device)])
        for row, k in enumerate(good):
            row_of[ok[k]] = base + row

    for window in _windows(paths):
        decoded_order = window
        res = _decode_all(window, worker_count, is_cancelled, lambda: None, flush=_thumb)
        if res is None:
            return []
        errors.update(res[1])

    keeper_failure_counts: Counter = Counter()
    keeper_failure_samples: dict = {}
    entry_failure_counts: Counter = Counter()
    entry_failure_samples: dict = {}
    plan = []
    ia, ib = [], []
    for cl in clusters:
        keep = next((e for e in cl.files if e.file.file_id == cl.keeper_id), None)
        if not keep:
            plan.append((cl, None))
            continue
        if keep.file.path not in row_of:
            key = errors.get(keep.file.path, "RuntimeError: decode failed")
            keeper_failure_counts[key] += 1
            keeper_failure_samples.setdefault(key, keep.file.path)
            plan.append((cl, None))
            continue
        members = []
        for e in cl.files:
            row = row_of.get(e.file.path)
            if row is None:
                key = errors.get(e.file.path, "RuntimeError: decode failed")
                entry_failure_counts[key] += 1
                entry_failure_samples.setdefault(key, e.file.path)
                continue
            members.append((e, len(ia)))
            ia.append(row)
            ib.append(row_of[keep.file.path])
        plan.append((cl, members))
    sad = []
    if ia:
        import torch

        sad = ops.plane_sad_pairs(torch.cat(chunks), ia, ib).cpu().tolist()
    n_px = float(thumb_size * thumb_size)

    out = []
    done = 0
    for cl, members in plan:
        if is_cancelled and is_cancelled():
            return []
        if members is not None:
            # np.mean of the int16 differences is sad / n in float64; then / 255.0 (reference :210-212)
            oks = [e for e, q in members if (sad[q] / n_px) / 255.0 <= mae_thr]
            if len(oks) >= 2:
                out.append(_rebuild_cluster_like(cl, oks))
        done += 1
        if tick and (done % 16 == 0 or done == total):
            tick(done, total)

    if keeper_failure_counts:
        log.warning("Pixel MAE skipped %d cluster(s) due to keeper load errors: %s", sum(keeper_failure_counts.values()),
                    _format_failure_summary(keeper_failure_counts, keeper_failure_samples))
    if entry_failure_counts:
        log.warning("Pixel MAE excluded %d file(s) due to image load errors: %s", sum(entry_failure_counts.values()),
                    _format_failure_summary(entry_failure_counts, entry_failure_samples))
    return out


__all__ = ["tile_ahash_bits", "tile_hamming", "refine_by_tilehash_parallel", "refine_by_pixels_parallel",
           "tile_ahash_bits_many"]
