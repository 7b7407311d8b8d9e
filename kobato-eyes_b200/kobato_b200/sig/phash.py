"""Drop-in for the reference's ``sig.phash`` (src/sig/phash.py:21-66) backed by the K1 kernel.

Same names, argument meaning, return convention (signed 64-bit ints) and error behaviour:
``RuntimeError`` when NumPy/Pillow are missing (reference :23, :36).  The arithmetic — Pillow's
``convert("L")`` + ``resize(LANCZOS)``, the DCT, the mean threshold — runs in
``ke_phash_batch``; image decode stays on the host in Pillow.  ``phash_dhash_many`` is the batched
sibling the per-image functions are built on.
"""
from __future__ import annotations

from typing import Iterable, Sequence

try:
    import numpy as np
except ModuleNotFoundError:  # pragma: no cover
    np = None  # type: ignore[assignment]

try:
    from PIL import Image
except ModuleNotFoundError:  # pragma: no cover
    Image = None  # type: ignore[assignment]

_DIRECT_MODES = {"L": 1, "RGB": 3, "RGBA": 4, "RGBX": 4}


def _to_signed(value: int) -> int:
    return value - (1 << 64) if value >= (1 << 63) else value


def _decoded_array(image):
    """Decoded uint8 array the kernel consumes.  'L', 'RGB' and 'RGBA' go to the GPU as they are
    (the kernel applies Pillow's own luma); every other mode takes the reference's own
    ``convert("L")`` on the host first, which is exactly what the reference does (:25)."""
    if Image is None or np is None:
        raise RuntimeError("NumPy and Pillow are required to compute perceptual hashes")
    if isinstance(image, np.ndarray):
        arr = np.ascontiguousarray(image, dtype=np.uint8)
        if arr.ndim == 2 or (arr.ndim == 3 and arr.shape[2] in (1, 3, 4)):
            return arr
        raise ValueError("expected an HxW or HxWx{1,3,4} uint8 array")
    if image.mode not in _DIRECT_MODES:
        image = image.convert("L")
    return np.asarray(image, dtype=np.uint8)


def phash_dhash_many(images: Iterable) -> list[tuple[int, int]]:
    """[(phash, dhash)] (signed) for PIL images / decoded arrays of ANY sizes: images are grouped
    by geometry and each group is hashed in one ``ke_phash_batch`` launch."""
    from .. import ops

    arrays = [_decoded_array(im) for im in images]
    out: list[tuple[int, int] | None] = [None] * len(arrays)
    groups: dict[tuple, list[int]] = {}
    for k, a in enumerate(arrays):
        groups.setdefault(a.shape, []).append(k)
    for shape, members in groups.items():
        batch = np.stack([arrays[k] for k in members])
        ph, dh = ops.phash_dhash_batch(batch)
        for k, p, d in zip(members, ph.tolist(), dh.tolist()):
            out[k] = (int(p), int(d))
    return out  # type: ignore[return-value]


def phash(image) -> int:
    """Perceptual hash of one image (reference :33-46), signed 64-bit."""
    return phash_dhash_many([image])[0][0]


def dhash(image) -> int:
    """Difference hash of one image (reference :49-57), signed 64-bit."""
    return phash_dhash_many([image])[0][1]


def hamming64(a: int, b: int) -> int:
    """Hamming distance of two 64-bit hashes, signed or unsigned (reference :60-63).
    Scalar convenience only; the batched search is ``kobato_b200.ops.hamming_join``."""
    return ((int(a) ^ int(b)) & 0xFFFFFFFFFFFFFFFF).bit_count()


def hamming64_many(a: Sequence[int], b: Sequence[int]):
    """Element-wise distances of two equal-length hash arrays (uint8 result)."""
    x = np.asarray(a).astype(np.int64).view(np.uint64) ^ np.asarray(b).astype(np.int64).view(np.uint64)
    return np.bitwise_count(x).astype(np.uint8)


__all__ = ["phash", "dhash", "hamming64", "phash_dhash_many"]
