"""Drop-in for the reference's ``core.signature`` (src/core/signature.py:17-62)."""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Optional

from ..sig.phash import phash_dhash_many

log = logging.getLogger(__name__)


def _to_signed64(x: int) -> int:
    v = int(x) & ((1 << 64) - 1)
    return v - (1 << 64) if v >= (1 << 63) else v


def compute_signatures_from_image(im) -> tuple[int, int]:
    """(phash, dhash) of one PIL image, signed; exceptions propagate to the caller (reference :24-28).
    Both hashes come out of a single kernel launch."""
    p, d = phash_dhash_many([im])[0]
    return (_to_signed64(p), _to_signed64(d))


def ensure_signatures(conn, file_id: int, *, image=None, path: Optional[str | Path] = None, force: bool = False,
                      loader=None, upsert=None) -> bool:
    """Compute and upsert the signature row unless it exists (reference :31-62).

    ``loader`` / ``upsert`` default to the reference's own host code
    (``utils.image_io.safe_load_image`` and ``db.repository.upsert_signatures``), imported lazily so
    this module works stand-alone in tests.  Returns True when a row exists afterwards, False on any
    failure (logged, never raised)."""
    try:
        if not force:
            if conn.execute("SELECT 1 FROM signatures WHERE file_id=? LIMIT 1", (file_id,)).fetchone() is not None:
                return True
        if image is None:
            if path is None:
                return False
            if loader is None:
                from utils.image_io import safe_load_image as loader  # reference host code
            image = loader(Path(path))
            if image is None:
                return False
        p, d = compute_signatures_from_image(image)
        if upsert is None:
            try:
                from db.repository import upsert_signatures as upsert  # reference host code
            except ModuleNotFoundError:
                upsert = _upsert_signatures
        upsert(conn, file_id=file_id, phash_u64=p, dhash_u64=d)
        return True
    except Exception as exc:
        log.warning("ensure_signatures failed for %s: %s", path or f"file_id={file_id}", exc)
        return False


def _upsert_signatures(conn, *, file_id: int, phash_u64: int, dhash_u64: int) -> None:
    """Stand-alone equivalent of db.repository.upsert_signatures (src/db/repository.py:257-267)."""
    conn.execute(
        "INSERT INTO signatures (file_id, phash_u64, dhash_u64) VALUES (?, ?, ?) "
        "ON CONFLICT(file_id) DO UPDATE SET phash_u64 = excluded.phash_u64, dhash_u64 = excluded.dhash_u64",
        (file_id, phash_u64, dhash_u64),
    )
