"""Counter-based synthetic workloads: integer-only, so the NumPy generators
here and the CUDA twins in ``csrc/ke_synth.cu`` produce identical bytes.

* ``synth_hashes``  — uint64 table for the join configs (C3/C5): random hashes, the last
  ``planted`` fraction being copies of an earlier hash with 0..``max_flips`` bit flips.
* ``synth_image``   — HxWxC uint8 image ``i`` of a set: two bilinear-upsampled random grids
  (9x9 and 33x33 control points per channel) plus per-pixel noise; the last ``planted``
  fraction of a set are near-duplicates (gain 1.02 / re-noised / shifted by one pixel) of an
  earlier image.
"""
from __future__ import annotations

import numpy as np

SEED = 20261018
_M = (1 << 64) - 1
K_ITEM = 0x9E3779B97F4A7C15
K_ELEM = 0xD1B54A32D192ED03
K_CHAN = 0x8CB92BA72F3D8DD7


def splitmix64(x):
    """SplitMix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _mix(seed: int, item, elem=0):
    with np.errstate(over="ignore"):
        a = np.uint64(seed & _M) ^ (np.asarray(item, dtype=np.uint64) * np.uint64(K_ITEM))
        return splitmix64(splitmix64(a) ^ (np.asarray(elem, dtype=np.uint64) * np.uint64(K_ELEM)))


def synth_hashes(n: int, seed: int = SEED, planted: float = 0.05, max_flips: int = 12) -> np.ndarray:
    n = int(n)
    idx = np.arange(n, dtype=np.uint64)
    h = _mix(seed, idx, 0)
    n_base = n - (n * int(round(planted * 1000))) // 1000
    if n_base < 1 or n_base == n:
        return h
    t = idx[n_base:]
    src = (_mix(seed, t, 1) % np.uint64(n_base)).astype(np.int64)
    flips = (_mix(seed, t, 2) % np.uint64(max_flips + 1)).astype(np.int64)
    out = h.copy()
    v = h[src]
    for k in range(max_flips):  # k-th flip position; repeats cancel, so distance <= flips
        bit = (_mix(seed, t, 3 + k) % np.uint64(64)).astype(np.uint64)
        on = (flips > k)
        v = np.where(on, v ^ (np.uint64(1) << bit), v)
    out[n_base:] = v
    return out


# ------------------------------------------------------------------ images


def _lerp_grid(seed: int, src: int, ch: int, g: int, h: int, w: int, salt: int) -> np.ndarray:
    """Bilinear upsample (8-bit fractions) of a g x g grid of random bytes to h x w, values 0..255."""
    gy, gx = np.meshgrid(np.arange(g, dtype=np.uint64), np.arange(g, dtype=np.uint64), indexing="ij")
    with np.errstate(over="ignore"):
        elem = (gy * np.uint64(g) + gx) + np.uint64(salt) * np.uint64(1 << 20) + np.uint64(ch) * np.uint64(K_CHAN)
    grid = (_mix(seed, src, elem) >> np.uint64(56)).astype(np.int64)  # top byte
    fy = (np.arange(h, dtype=np.int64) * (g - 1) * 65536) // h
    fx = (np.arange(w, dtype=np.int64) * (g - 1) * 65536) // w
    y0, ty = fy >> 16, (fy >> 8) & 255
    x0, tx = fx >> 16, (fx >> 8) & 255
    g00 = grid[y0[:, None], x0[None, :]]
    g01 = grid[y0[:, None], x0[None, :] + 1]
    g10 = grid[y0[:, None] + 1, x0[None, :]]
    g11 = grid[y0[:, None] + 1, x0[None, :] + 1]
    top = g00 * (256 - tx)[None, :] + g01 * tx[None, :]
    bot = g10 * (256 - tx)[None, :] + g11 * tx[None, :]
    return (top * (256 - ty)[:, None] + bot * ty[:, None]) >> 16


def image_source(i: int, n_set: int, seed: int = SEED, planted: float = 0.05):
    """(source index, variant) of image i: variant 0 = original, 1 gain, 2 re-noised, 3 shifted."""
    n_base = n_set - (n_set * int(round(planted * 1000))) // 1000
    if i < n_base or n_base < 1:
        return i, 0
    src = int(_mix(seed ^ 0x5EED, i, 1) % np.uint64(n_base))
    variant = 1 + int(_mix(seed ^ 0x5EED, i, 2) % np.uint64(3))
    return src, variant


def synth_image(i: int, h: int, w: int, c: int = 3, *, n_set: int = 1 << 30, seed: int = SEED,
                planted: float = 0.05) -> np.ndarray:
    src, variant = image_source(i, n_set, seed, planted)
    shift = 1 if variant == 3 else 0
    amp = 3 if variant == 2 else 8
    noise_item = i if variant == 2 else src
    out = np.empty((h, w, c), np.uint8)
    xs = np.minimum(np.arange(w, dtype=np.int64) + shift, w - 1)
    pix = (np.arange(h, dtype=np.uint64)[:, None] * np.uint64(w) + xs.astype(np.uint64)[None, :])
    for ch in range(c):
        coarse = _lerp_grid(seed, src, ch, 9, h, w, 1)[:, xs]
        fine = _lerp_grid(seed, src, ch, 33, h, w, 2)[:, xs]
        base = (coarse * 3 + fine) >> 2
        r = _mix(seed ^ 0xA5A5, noise_item, pix * np.uint64(4) + np.uint64(ch))
        noise = ((r >> np.uint64(40)) % np.uint64(2 * amp + 1)).astype(np.int64) - amp
        v = base + noise
        if variant == 1:
            v = (v * 261 + 128) >> 8
        out[:, :, ch] = np.clip(v, 0, 255).astype(np.uint8)
    return out if c > 1 else out[:, :, 0]


def synth_images(start: int, count: int, h: int, w: int, c: int = 3, **kw) -> np.ndarray:
    return np.stack([synth_image(start + k, h, w, c, **kw) for k in range(count)])
