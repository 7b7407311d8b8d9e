"""Drop-in for the reference's ``core.fastsig`` (src/core/fastsig.py:19-126).

Same public names and contracts:

* ``compute_signatures_mp(tasks, *, max_workers, chunksize, progress, cancel_fn)`` returns the
  ordered ``[(file_id, phash, dhash)]`` (signed 64-bit), silently dropping files that are missing
  or fail to decode (reference :36-37), calling ``progress(done, total)`` every 200 results and at
  the end (:95-99) and returning the partial list as soon as ``cancel_fn()`` is true (:86-90).
* ``bulk_upsert_signatures`` / ``fast_fill_missing_signatures`` / ``_fast_pragmas`` keep the SQLite
  side exactly as the reference has it (the ``signatures(file_id, phash_u64, dhash_u64)`` table).

What changed is where the arithmetic runs: the reference spawns a process pool and hashes one
image per call on CPU cores; here host threads only *decode* (Pillow releases the GIL) and each
window of decoded images is hashed by ONE ``ke_phash_batch`` launch per image geometry from the
parent process (no CUDA context is ever created in a worker).
"""
from __future__ import annotations

import os
import sqlite3
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Iterable, List, Optional, Tuple

import numpy as np
from PIL import Image

from ..sig.phash import _decoded_array, dhash, phash, phash_dhash_many  # noqa: F401  (phash/dhash: patchable seams)

U64MASK = (1 << 64) - 1
PROGRESS_EVERY = 200


def _to_signed64(x: int) -> int:
    v = int(x) & U64MASK
    return v - (1 << 64) if v >= (1 << 63) else v


def _decode_worker(task: Tuple[int, str]):
    """(file_id, path) -> (file_id, decoded uint8 array) or None.  Mirrors the guard rails of the
    reference's ``_compute_worker`` (:24-37): missing path, directory or any decode error -> None."""
    fid, p = task
    try:
        path = Path(p)
        if not path.exists() or not path.is_file():
            return None
        with Image.open(path) as im:
            return int(fid), _decoded_array(im)
    except Exception:
        return None


def _compute_worker(task: Tuple[int, str]) -> Tuple[int, int, int] | None:
    """Single-file path with the reference's exact shape: (file_id, path) -> (file_id, ph, dh)."""
    fid, p = task
    try:
        path = Path(p)
        if not path.exists() or not path.is_file():
            return None
        with Image.open(path) as im:
            ph = _to_signed64(phash(im))
            dh = _to_signed64(dhash(im))
        return (int(fid), ph, dh)
    except Exception:
        return None


def _fast_pragmas(conn: sqlite3.Connection) -> None:
    for pragma in ("journal_mode=WAL", "synchronous=OFF", "temp_store=MEMORY", "mmap_size=30000000000"):
        conn.execute(f"PRAGMA {pragma}")


def bulk_upsert_signatures(conn: sqlite3.Connection, rows: Iterable[Tuple[int, int, int]]) -> int:
    """One ``executemany`` upsert of (file_id, phash, dhash), values wrapped to signed 64-bit."""
    payload = [(int(fid), _to_signed64(ph), _to_signed64(dh)) for fid, ph, dh in rows]
    if not payload:
        return 0
    with conn:
        cur = conn.executemany(
            "INSERT INTO signatures (file_id, phash_u64, dhash_u64) VALUES (?, ?, ?) "
            "ON CONFLICT(file_id) DO UPDATE SET phash_u64 = excluded.phash_u64, dhash_u64 = excluded.dhash_u64",
            payload,
        )
    return cur.rowcount or 0


def hash_decoded(decoded: list) -> list[Tuple[int, int, int] | None]:
    """[(fid, array) | None] -> [(fid, ph, dh) | None], one GPU launch per distinct geometry."""
    live = [k for k, d in enumerate(decoded) if d is not None]
    out: list[Tuple[int, int, int] | None] = [None] * len(decoded)
    if not live:
        return out
    try:
        sigs = phash_dhash_many([decoded[k][1] for k in live])
    except Exception:
        # one failing geometry (bad shape, rows that do not fit shared memory, a CUDA error status ...) must not sink
        # the window: the reference's worker swallows ANY exception of a file and returns None (:36-37), so retry one by
        # one and drop only the failures
        sigs = []
        for k in live:
            try:
                sigs.append(phash_dhash_many([decoded[k][1]])[0])
            except Exception:
                sigs.append(None)
    for k, s in zip(live, sigs):
        if s is not None:
            out[k] = (decoded[k][0], _to_signed64(s[0]), _to_signed64(s[1]))
    return out


def compute_signatures_mp(
    tasks: List[Tuple[int, str]],
    *,
    max_workers: Optional[int] = None,
    chunksize: int = 64,
    progress: Optional[Callable[[int, int], None]] = None,
    cancel_fn: Optional[Callable[[], bool]] = None,
) -> List[Tuple[int, int, int]]:
    """(file_id, path) list -> ordered (file_id, ph, dh) list; see the module docstring."""
    if not tasks:
        return []
    total = len(tasks)
    done = 0
    results: List[Tuple[int, int, int]] = []
    workers = max_workers or max(1, (os.cpu_count() or 4) - 1)
    # Decoded images in flight between GPU launches: at most chunksize x workers files AND at most KE_SIG_WINDOW_MB of
    # pixels (default 512 MB) — the reference holds one image per worker, and multi-megapixel photographs are ~50 MB
    # each once decoded, so the window is bounded by bytes, not only by count.
    window = max(1, int(chunksize)) * workers
    window_bytes = int(os.environ.get("KE_SIG_WINDOW_MB", "512")) << 20
    step = max(1, workers) * 2
    with ThreadPoolExecutor(max_workers=workers) as pool:
        start = 0
        while start < total:
            if cancel_fn and cancel_fn():
                pool.shutdown(wait=False, cancel_futures=True)
                return results
            decoded: list = []
            held = 0
            while start < total and len(decoded) < window and held < window_bytes:
                part = list(pool.map(_decode_worker, tasks[start:start + min(step, window - len(decoded))]))
                held += sum(d[1].nbytes for d in part if d is not None)
                decoded += part
                start += len(part)
            for out in hash_decoded(decoded):
                if cancel_fn and cancel_fn():
                    pool.shutdown(wait=False, cancel_futures=True)
                    return results
                done += 1
                if out is not None:
                    results.append(out)
                if progress and (done % PROGRESS_EVERY == 0 or done == total):
                    try:
                        progress(done, total)
                    except Exception:
                        pass
    return results


def fast_fill_missing_signatures(
    db_path: str,
    items: List[Tuple[int, str]],
    *,
    max_workers: Optional[int] = None,
    chunksize: int = 64,
    progress: Optional[Callable[[int, int], None]] = None,
    apply_to_db: bool = True,
    unsafe_fast: bool = True,
    cancel_fn: Optional[Callable[[], bool]] = None,
) -> List[Tuple[int, int, int]]:
    """Compute the missing signatures and (optionally) upsert them in one transaction."""
    computed = compute_signatures_mp(items, max_workers=max_workers, chunksize=chunksize, progress=progress,
                                     cancel_fn=cancel_fn)
    if apply_to_db and computed:
        with sqlite3.connect(db_path) as conn:
            if unsafe_fast:
                _fast_pragmas(conn)
            bulk_upsert_signatures(conn, computed)
    return computed


def compute_signatures_arrays(images: np.ndarray):
    """Batched form for callers that already hold decoded images: uint8 [n,h,w(,c)] (numpy or
    CUDA tensor) -> (phash int64[n], dhash int64[n])."""
    from .. import ops

    return ops.phash_dhash_batch(images)
