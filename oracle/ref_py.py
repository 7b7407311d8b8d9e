"""Python restatement of the reference's duplicate-detection path — TEST INFRASTRUCTURE ONLY.

This restates, on the same third-party libraries the reference calls, what these reference
functions compute (paths relative to /root/reference):

* ``phash`` / ``dhash`` / ``hamming64``            src/sig/phash.py:21-63
* ``scan_edges`` / ``build_clusters``             src/dup/scanner.py:211-415 (LSH banding + DSU)
* ``structural_similarity``                        scikit-image 0.25.2
  ``skimage/metrics/_structural_similarity.py`` (third-party, absent in this image — restated
  from its published algorithm on ``scipy.ndimage.uniform_filter``; *parity unpinned*)
* ``compute_ssim`` / ``refine_decision``          src/dup/refine.py:44-52, 100-117
* ``cluster_matches``                              src/dup/cluster.py:22-70

It is written array-at-a-time (NumPy) rather than as the reference's per-item loops; equality
with the live reference is asserted by ``tests/test_oracle_pinned.py`` whenever
``/root/reference`` is mounted and by the vectors in ``tests/golden/`` otherwise.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import PurePath
from typing import Iterable, Sequence

import numpy as np

U64 = (1 << 64) - 1


# --------------------------------------------------------------------------- hashes


def _gray_plane(image, size_wh):
    """src/sig/phash.py:21-26 — convert('L') then LANCZOS resize, as float32."""
    from PIL import Image

    plane = image.convert("L").resize(size_wh, Image.Resampling.LANCZOS)
    return np.asarray(plane, dtype=np.float32)


def _pack_msb_first(bits: np.ndarray) -> int:
    by = np.packbits(np.asarray(bits, dtype=bool).ravel())  # MSB-first, like the reference's shift loop
    return int.from_bytes(by.tobytes(), "big")


def to_signed64(v: int) -> int:
    """src/sig/phash.py:29-30, src/core/signature.py:17-21."""
    v = int(v) & U64
    return v - (1 << 64) if v >> 63 else v


def phash_from_plane(plane32: np.ndarray) -> int:
    """src/sig/phash.py:37-46 on an already-resized 32x32 plane (unsigned result)."""
    import cv2

    coeff = cv2.dct(np.asarray(plane32, dtype=np.float32))
    low = coeff[:8, :8].reshape(-1)
    thr = low[1:].mean()  # float32 mean of the 63 AC terms; DC is compared but not averaged
    return _pack_msb_first(low > thr)


def phash(image) -> int:
    return to_signed64(phash_from_plane(_gray_plane(image, (32, 32))))


def dhash_from_plane(plane9x8: np.ndarray) -> int:
    p = np.asarray(plane9x8)
    return _pack_msb_first(p[:, 1:] > p[:, :-1])


def dhash(image) -> int:
    return to_signed64(dhash_from_plane(_gray_plane(image, (9, 8))))


def hamming64(a: int, b: int) -> int:
    return ((int(a) ^ int(b)) & U64).bit_count()


# --------------------------------------------------------------------------- scanner

EXT_PRIORITY = {"png": 4, "apng": 4, "webp": 3, "tiff": 2, "tif": 2, "bmp": 1, "gif": 1}


@dataclass(frozen=True)
class FileRec:
    """The fields of the reference's DuplicateFile the scanner reads."""

    file_id: int
    path: str
    size: int | None
    width: int | None
    height: int | None
    phash: int
    embedding: tuple | None = None


def _cosine(u, v):
    if u is None or v is None or len(u) == 0 or len(v) == 0 or len(u) != len(v):
        return None
    u = [float(x) for x in u]
    v = [float(x) for x in v]
    dot = sum(a * b for a, b in zip(u, v))
    nu = sum(a * a for a in u) ** 0.5
    nv = sum(b * b for b in v) ** 0.5
    if nu == 0.0 or nv == 0.0:
        return None
    return dot / (nu * nv)


def scan_edges(files: Sequence[FileRec], *, hamming_threshold: int = 8, size_ratio: float | None = None,
               band_bits: int = 16, band_count: int = 4, cosine_threshold: float | None = None,
               bucket_pair_cap: int | None = None) -> dict[tuple[int, int], int]:
    """Edge set of src/dup/scanner.py:227-290 as {(min_id, max_id): hamming}.

    A pair is an edge iff the two files share the value of at least one band (whose bucket is
    not skipped by the pair cap), have different file ids, pass the size-ratio gate, are
    within the Hamming threshold and pass the cosine gate.  When several list positions map to
    the same id pair, the reference keeps the FIRST one it visits (bucket insertion order);
    the distance can only differ if one file_id appears with two different hashes."""
    n = len(files)
    if n == 0:
        return {}
    ph = np.array([f.phash & U64 for f in files], dtype=np.uint64)
    mask = np.uint64((1 << band_bits) - 1)
    visits: list[tuple[int, int, int, int]] = []  # (first-key order, i, j) candidates per band
    for band in range(band_count):
        vals = (ph >> np.uint64(band * band_bits)) & mask
        order = np.argsort(vals, kind="stable")
        sv = vals[order]
        starts = np.flatnonzero(np.r_[True, sv[1:] != sv[:-1]])
        ends = np.r_[starts[1:], n]
        for s, e in zip(starts, ends):
            m = e - s
            if m < 2:
                continue
            if bucket_pair_cap is not None and (m * (m - 1)) // 2 > bucket_pair_cap:
                continue
            idx = order[s:e]  # ascending list positions == the reference's bucket order
            ii, jj = np.triu_indices(m, k=1)
            # reference bucket dict order: keyed by first insertion = (position of first member, band)
            first = int(idx[0])
            for a, b in zip(idx[ii], idx[jj]):
                visits.append((first, band, int(a), int(b)))
    # the reference iterates buckets in dict insertion order: bucket created when its first member
    # (lowest list position) is processed, bands in ascending order for that member
    visits.sort(key=lambda t: (t[0], t[1]))
    edges: dict[tuple[int, int], int] = {}
    for _, _, i, j in visits:
        a, b = files[i], files[j]
        if a.file_id == b.file_id:
            continue
        if size_ratio is not None and size_ratio > 0:
            sa, sb = a.size or 0, b.size or 0
            if sa > 0 and sb > 0 and (min(sa, sb) / max(sa, sb)) < size_ratio:
                continue
        d = hamming64(a.phash, b.phash)
        if d > hamming_threshold:
            continue
        if cosine_threshold is not None:
            cs = _cosine(a.embedding, b.embedding)
            if cs is not None and cs < cosine_threshold:
                continue
        key = (a.file_id, b.file_id) if a.file_id < b.file_id else (b.file_id, a.file_id)
        edges.setdefault(key, d)
    return edges


def scan_table(phash, file_id=None, size=None, *, threshold: int = 8, band_bits: int = 16, band_count: int = 4,
               size_ratio: float | None = None, pair_cap: int | None = None) -> dict:
    """What ke_scan_table_host returns, restated on ``scan_edges`` (src/dup/scanner.py:227-318) for DISTINCT file ids:
    members (rows with an edge) grouped by connected component, label = smallest row, best = min edge distance."""
    ph = np.asarray(phash).reshape(-1)
    ph = ph.view(np.uint64) if ph.dtype == np.int64 else ph.astype(np.uint64)
    n = len(ph)
    ids = np.arange(n, dtype=np.int64) if file_id is None else np.asarray(file_id, np.int64)
    sizes = np.zeros(n, np.int64) if size is None else np.asarray(size, np.int64)
    files = [FileRec(file_id=int(k), path=f"{k}.jpg", size=int(sizes[k]), width=None, height=None, phash=int(ph[k]))
             for k in range(n)]  # rows stand in for ids so that equal ids can be told apart below
    row_edges = scan_edges(files, hamming_threshold=threshold, size_ratio=size_ratio, band_bits=band_bits,
                           band_count=band_count, bucket_pair_cap=pair_cap)
    row_edges = {k: d for k, d in row_edges.items() if ids[k[0]] != ids[k[1]]}
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    best = {}
    for (i, j), d in row_edges.items():
        ri, rj = find(i), find(j)
        if ri != rj:
            parent[max(ri, rj)] = min(ri, rj)
        for r in (i, j):
            best[r] = min(best.get(r, 255), d)
    members = sorted(best, key=lambda r: (find(r), r))
    index = np.array(members, np.int64)
    label = np.array([find(r) for r in members], np.int64)
    cuts = np.flatnonzero(np.diff(label)) + 1 if len(label) else np.zeros(0, np.int64)
    offsets = np.concatenate([[0], cuts, [len(label)]]).astype(np.int64) if len(label) else np.zeros(1, np.int64)
    ei = np.array(sorted(row_edges), np.int64).reshape(-1, 2)
    return {"index": index, "label": label, "best": np.array([best[r] for r in members], np.int32), "offsets": offsets,
            "edges": (ei[:, 0].astype(np.uint32), ei[:, 1].astype(np.uint32),
                      np.array([row_edges[tuple(k)] for k in ei.tolist()], np.uint8)),
            "stats": {"edges": len(row_edges), "members": len(members), "clusters": max(0, len(offsets) - 1)}}


def _ext_priority(path: str) -> int:
    return EXT_PRIORITY.get(PurePath(path).suffix.lower().lstrip("."), 0)


def _rank_key(f: FileRec):
    p = PurePath(f.path)
    return (-(f.size or 0), -((f.width or 0) * (f.height or 0)), -_ext_priority(f.path), p.suffix.lower(),
            p.name.lower(), f.file_id)


def build_clusters(files: Iterable[FileRec], **cfg):
    """src/dup/scanner.py:304-356 on top of scan_edges.

    Returns a list of dicts {keeper, members:[(file_id, best_hamming)...]} in the reference's
    order: members keeper-first then by (-size, -resolution, -ext priority, name, id); clusters
    by (-max size, first path)."""
    files = [f for f in files if f.phash is not None]
    edges = scan_edges(files, **cfg)
    if not edges:
        return []
    by_id = {f.file_id: f for f in files}  # last occurrence wins, like the reference's dict comprehension
    parent: dict[int, int] = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    best: dict[int, int] = {}
    for (a, b), d in edges.items():
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
        for fid in (a, b):
            if fid not in best or d < best[fid]:
                best[fid] = d
    groups: dict[int, list[int]] = {}
    for fid in parent:
        groups.setdefault(find(fid), []).append(fid)
    out = []
    for members in groups.values():
        if len(members) < 2:
            continue
        recs = [by_id[m] for m in sorted(members)]
        keeper = min(recs, key=_rank_key).file_id
        recs.sort(key=lambda f: (0 if f.file_id == keeper else 1, -(f.size or 0),
                                 -((f.width or 0) * (f.height or 0)), -_ext_priority(f.path),
                                 PurePath(f.path).name.lower(), f.file_id))
        out.append({"keeper": keeper, "members": [(f.file_id, best.get(f.file_id)) for f in recs],
                    "_max_size": max(f.size or 0 for f in recs), "_first_path": PurePath(recs[0].path).as_posix().lower()})
    out.sort(key=lambda c: (-c["_max_size"], c["_first_path"]))
    return [{"keeper": c["keeper"], "members": c["members"]} for c in out]


# --------------------------------------------------------------------------- SSIM


def structural_similarity(im1, im2, *, data_range: float = 1.0, win_size: int = 7) -> float:
    """scikit-image 0.25.2 structural_similarity defaults as the reference calls it
    (src/dup/refine.py:52): uniform 7x7 window, sample covariance, K1=.01, K2=.03, the maps in
    the input float type (float32 here), border of 3 cropped, float64 mean."""
    from scipy.ndimage import uniform_filter

    im1 = np.asarray(im1)
    im2 = np.asarray(im2)
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if any(s < win_size for s in im1.shape):
        raise ValueError("win_size exceeds image extent.")
    ft = np.float32 if im1.dtype in (np.float32, np.float16) else np.float64
    a = im1.astype(ft, copy=False)
    b = im2.astype(ft, copy=False)
    npx = win_size ** a.ndim
    cov_norm = npx / (npx - 1)
    ux = uniform_filter(a, size=win_size)
    uy = uniform_filter(b, size=win_size)
    uxx = uniform_filter(a * a, size=win_size)
    uyy = uniform_filter(b * b, size=win_size)
    uxy = uniform_filter(a * b, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux**2 + uy**2 + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    core = s[tuple(slice(pad, n - pad) for n in s.shape)]
    return float(core.mean(dtype=np.float64))


def structural_similarity_gaussian(im1, im2, *, data_range: float = 1.0) -> float:
    """scikit-image 0.25.2 ``structural_similarity(..., gaussian_weights=True)``: sigma 1.5, truncate 3.5 (11 taps,
    win_size 11 -> crop 5, cov_norm 121/120), ``skimage.filters.gaussian`` = ``scipy.ndimage.gaussian_filter`` with
    mode 'reflect' on the float32 images.  NOT the reference's path (src/dup/refine.py:52 keeps the uniform window);
    restated for the optional ``gaussian`` flag of ke_ssim_batch."""
    from scipy.ndimage import gaussian_filter

    a = np.asarray(im1, dtype=np.float32)
    b = np.asarray(im2, dtype=np.float32)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    sigma, truncate = 1.5, 3.5
    win_size = 2 * int(truncate * sigma + 0.5) + 1
    if any(s < win_size for s in a.shape):
        raise ValueError("win_size exceeds image extent.")

    def filt(x):
        return gaussian_filter(x, sigma=sigma, truncate=truncate, mode="reflect")

    npx = win_size ** a.ndim
    cov_norm = npx / (npx - 1)
    ux, uy = filt(a), filt(b)
    uxx, uyy, uxy = filt(a * a), filt(b * b), filt(a * b)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux**2 + uy**2 + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    core = s[tuple(slice(pad, n - pad) for n in s.shape)]
    return float(core.mean(dtype=np.float64))


def ssim_gaussian_of_planes(pa: np.ndarray, pb: np.ndarray) -> float:
    return structural_similarity_gaussian(pa.astype(np.float32) / 255.0, pb.astype(np.float32) / 255.0)


def ssim_planes(img_a, img_b):
    """Host-side preparation of src/dup/refine.py:44-49: the two uint8 'L' planes SSIM runs on."""
    from PIL import Image, ImageOps

    size = (min(img_a.width, img_b.width), min(img_a.height, img_b.height))
    if size[0] == 0 or size[1] == 0:
        size = (max(img_a.width, img_b.width), max(img_a.height, img_b.height))
    pa = ImageOps.fit(img_a.convert("L"), size, Image.Resampling.BICUBIC)
    pb = ImageOps.fit(img_b.convert("L"), size, Image.Resampling.BICUBIC)
    return np.asarray(pa, dtype=np.uint8), np.asarray(pb, dtype=np.uint8)


def ssim_of_planes(pa: np.ndarray, pb: np.ndarray) -> float:
    return structural_similarity(pa.astype(np.float32) / 255.0, pb.astype(np.float32) / 255.0, data_range=1.0)


def compute_ssim(img_a, img_b) -> float:
    return ssim_of_planes(*ssim_planes(img_a, img_b))


def orb_cross_check(da: np.ndarray, db: np.ndarray) -> list[tuple[int, int, int]]:
    """What ``cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db)`` returns (src/dup/refine.py:64) as sorted
    ``(queryIdx, trainIdx, distance)`` triples: the MUTUAL nearest neighbours in Hamming distance, the first index winning
    a tie on either side (OpenCV modules/features2d matchers.cpp + core batch_distance.cpp, restated; pinned against the
    live cv2 by tests/test_oracle_pinned.py)."""
    da = np.ascontiguousarray(da, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    if len(da) == 0 or len(db) == 0:
        return []
    dist = np.bitwise_count(da[:, None, :] ^ db[None, :, :]).sum(-1).astype(np.int64)  # [queries, trains]
    t_of_q = dist.argmin(1)  # argmin returns the first minimum
    q_of_t = dist.argmin(0)
    return [(int(i), int(t_of_q[i]), int(dist[i, t_of_q[i]])) for i in range(len(da)) if q_of_t[t_of_q[i]] == i]


def orb_match_counts(desc_a, desc_b) -> np.ndarray:
    """``len(matches)`` per pair — the CPU stand-in for kobato_b200.ops.orb_match_pairs."""
    return np.array([0 if a is None or b is None else len(orb_cross_check(a, b)) for a, b in zip(desc_a, desc_b)], np.int32)


def refine_decision(ssim_value, orb_ratio, *, ssim_thr: float = 0.9, orb_thr: float = 0.15, errors=()):
    """src/dup/refine.py:100-117 -> (is_duplicate, reason)."""
    hits = []
    if ssim_value is not None and ssim_value >= ssim_thr:
        hits.append(f"ssim>={ssim_thr}")
    if orb_ratio is not None and orb_ratio >= orb_thr:
        hits.append(f"orb>={orb_thr}")
    if hits:
        return True, ", ".join(hits)
    if errors:
        return False, ", ".join(errors)
    return False, "below thresholds"


def cluster_matches(matches: Iterable[tuple[int, int, bool]]):
    """src/dup/cluster.py:22-70 -> [(representative, sorted members)] sorted by representative."""
    parent: dict[int, int] = {}

    def find(x):
        parent.setdefault(x, x)
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b, dup in matches:
        if not dup:
            continue
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    comps: dict[int, list[int]] = {}
    for x in parent:
        comps.setdefault(find(x), []).append(x)
    return sorted((min(m), sorted(m)) for m in comps.values())


def bucket_pair_cap_from_env() -> int | None:
    raw = os.environ.get("KE_DUP_BUCKET_PAIR_CAP")
    if raw is None or not raw.strip():
        return None
    try:
        v = int(raw)
    except ValueError:
        return None
    return v if v > 0 else None


# ---------------------------------------------------------------------------------------------------
# N1: the refinement the shipped UI runs after a scan (src/ui/dup_refine_parallel.py)


def tile_ahash_bits_image(image, grid: int = 4, tile: int = 8) -> int:
    """src/ui/dup_refine_parallel.py:59-83 after the file is open (exif_transpose is the caller's):
    convert("L").resize((side, side), BILINEAR), per-tile mean threshold, bits in (gy, gx, ty, tx) order,
    little-endian packed integer."""
    from PIL import Image

    side = grid * tile
    gray = image.convert("L").resize((side, side), Image.Resampling.BILINEAR)
    arr = np.asarray(gray, dtype=np.uint8)
    a = arr.reshape(grid, tile, grid, tile).transpose(0, 2, 1, 3)
    means = a.mean(axis=(2, 3), keepdims=True)
    bits = (a > means).reshape(-1).astype(np.uint8)
    return int.from_bytes(np.packbits(bits, bitorder="little").tobytes(), "little")


def tile_hamming(a_bits: int, b_bits: int) -> int:
    """src/ui/dup_refine_parallel.py:86-88."""
    return (int(a_bits) ^ int(b_bits)).bit_count()


def small_gray(image, size: int = 128) -> np.ndarray:
    """src/ui/dup_refine_parallel.py:203-207 after the file is open."""
    from PIL import Image

    return np.asarray(image.convert("L").resize((size, size), Image.Resampling.BILINEAR), dtype=np.uint8)


def mae01(a: np.ndarray, b: np.ndarray) -> float:
    """src/ui/dup_refine_parallel.py:210-212."""
    return float(np.mean(np.abs(a.astype(np.int16) - b.astype(np.int16))) / 255.0)
