"""CPU oracle for the duplicate-detection hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``kobato-eyes_b200/kobato_b200``) never does; it fails loudly when the CUDA library is missing.

Two layers:

* ``oracle.ref_py``  — the reference's own Python path restated on the same third-party
  libraries the reference calls (Pillow, OpenCV, NumPy, SciPy).
* ``oracle.ke_oracle.c`` (this module's ctypes wrappers) — the arithmetic *inside* those
  libraries restated in plain C so that intermediate planes can be compared byte-for-byte
  and large cases finish in seconds.

* ``oracle/skimage_shim`` — a stand-in ``skimage.metrics`` (one function, forwarding to ``oracle.ref_py``) that lets
  the LIVE reference's ``dup.refine`` / ``dup.cluster`` import unmodified in the build container, so that the
  reference's own tests for that half of the path and a field-by-field comparison with the drop-ins can run.

Parity status: pHash/dHash/Hamming/scanner and the N1 refinement (tile aHash, small gray, MAE) are pinned against the
live reference (run in the build container, vectors in ``tests/golden/``).  SSIM is **parity unpinned against
scikit-image itself**: it is not installable here, so the restatement of its algorithm is pinned by the reference's
behavioural tests (``tests/dup/test_refine.py`` and ``tests/dup/test_cluster.py`` run through the shim) and by the
equality of ``refine_pair`` / ``ClusterBuilder`` with the live reference, not by scikit-image's own output.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

LANCZOS, BILINEAR, BICUBIC, BOX = 1, 2, 3, 4


def build(force: bool = False) -> Path:
    """Compile oracle/ke_oracle.c -> oracle/libke_oracle.so (gcc, no external deps)."""
    so = _HERE / "libke_oracle.so"
    src = _HERE / "ke_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O2", "-fPIC", "-std=gnu11", "-shared", "-o", str(so), str(src), "-lm"],
            check=True,
            env={**os.environ, "CC": "gcc"},
        )
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        u8p, i32p, i64p, u32p, u64p, f64p = (
            C.POINTER(C.c_uint8),
            C.POINTER(C.c_int32),
            C.POINTER(C.c_int64),
            C.POINTER(C.c_uint32),
            C.POINTER(C.c_uint64),
            C.POINTER(C.c_double),
        )
        L.ko_rgb_to_l.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int64, u8p]
        L.ko_rgb_to_l.restype = None
        L.ko_resample_ksize.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ko_resample_table.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_int]
        L.ko_resample_u8.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.ko_dhash_from_plane.argtypes = [u8p]
        L.ko_dhash_from_plane.restype = C.c_uint64
        L.ko_phash_from_plane_f64.argtypes = [u8p, f64p]
        L.ko_phash_from_plane_f64.restype = C.c_uint64
        L.ko_signature.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int64, u64p, u64p, f64p, u8p, u8p]
        L.ko_signature_batch.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, u64p, u64p, f64p]
        L.ko_hamming_join.argtypes = [u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                      u32p, u32p, u8p, C.c_int64]
        L.ko_hamming_join.restype = C.c_int64
        L.ko_ssim_u8.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int64, C.c_int64]
        L.ko_ssim_u8.restype = C.c_double
        L.ko_ssim_u8_exact.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int64, C.c_int64]
        L.ko_ssim_u8_exact.restype = C.c_double
        L.ko_ssim_batch.argtypes = [u8p, C.c_int, C.c_int, C.c_int64, i64p, i64p, C.c_int64, f64p, C.c_int]
        _LIB = L
    return _LIB


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _split(n: int, parts: int):
    parts = max(1, min(parts, n))
    edges = np.linspace(0, n, parts + 1).astype(np.int64)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(parts) if edges[i + 1] > edges[i]]


def resample_table(in_size: int, out_size: int, filter_id: int = LANCZOS):
    """(kk[out,ksize] int32, bounds[out,2] int32) exactly as Pillow's 8bpc resampler builds them."""
    L = lib()
    ks = L.ko_resample_ksize(in_size, out_size, filter_id)
    if ks <= 0:
        raise ValueError("bad resample geometry")
    kk = np.zeros((out_size, ks), np.int32)
    bd = np.zeros((out_size, 2), np.int32)
    rc = L.ko_resample_table(in_size, out_size, filter_id, _p(kk, C.c_int32), _p(bd, C.c_int32), ks)
    if rc:
        raise RuntimeError(f"ko_resample_table rc={rc}")
    return kk, bd


def to_l(img: np.ndarray) -> np.ndarray:
    """convert('L') of an HxW (gray) or HxWx{3,4} uint8 array."""
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        return img.copy()
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    lib().ko_rgb_to_l(_p(img, C.c_uint8), h, w, c, w * c, _p(out, C.c_uint8))
    return out


def resize_l(gray: np.ndarray, size_wh: tuple[int, int], filter_id: int = LANCZOS) -> np.ndarray:
    gray = np.ascontiguousarray(gray, np.uint8)
    ow, oh = size_wh
    out = np.empty((oh, ow), np.uint8)
    rc = lib().ko_resample_u8(_p(gray, C.c_uint8), gray.shape[0], gray.shape[1], _p(out, C.c_uint8), oh, ow, filter_id)
    if rc:
        raise RuntimeError(f"ko_resample_u8 rc={rc}")
    return out


def signature(img: np.ndarray):
    """(phash_u64 [f64 DCT], dhash_u64, min_margin, plane32[32,32], plane9x8[8,9]) of one decoded image."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    ph, dh, mm = C.c_uint64(), C.c_uint64(), C.c_double()
    p32 = np.empty((32, 32), np.uint8)
    p98 = np.empty((8, 9), np.uint8)
    rc = lib().ko_signature(_p(img, C.c_uint8), h, w, c, w * c, C.byref(ph), C.byref(dh), C.byref(mm),
                            _p(p32, C.c_uint8), _p(p98, C.c_uint8))
    if rc:
        raise RuntimeError(f"ko_signature rc={rc}")
    return ph.value, dh.value, mm.value, p32, p98


def signature_batch(imgs: np.ndarray, threads: int = 1):
    """imgs: [n,h,w] or [n,h,w,c] uint8 -> (phash u64[n], dhash u64[n], margin f64[n]) (f64 DCT)."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape[:3]
    c = 1 if imgs.ndim == 3 else imgs.shape[3]
    ph = np.zeros(n, np.uint64)
    dh = np.zeros(n, np.uint64)
    mm = np.zeros(n, np.float64)
    L = lib()
    stride = h * w * c

    def run(lo, hi):
        rc = L.ko_signature_batch(_p(imgs[lo:hi], C.c_uint8), hi - lo, h, w, c, stride, w * c,
                                  _p(ph[lo:hi], C.c_uint64), _p(dh[lo:hi], C.c_uint64), _p(mm[lo:hi], C.c_double))
        if rc:
            raise RuntimeError(f"ko_signature_batch rc={rc}")

    parts = _split(n, threads * 4 if threads > 1 else 1)
    if threads <= 1:
        for lo, hi in parts:
            run(lo, hi)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda r: run(*r), parts))
    return ph, dh, mm


def phash_f64_from_plane(p32: np.ndarray):
    p32 = np.ascontiguousarray(p32, np.uint8)
    mm = C.c_double()
    v = lib().ko_phash_from_plane_f64(_p(p32, C.c_uint8), C.byref(mm))
    return int(v), mm.value


def dhash_from_plane(p98: np.ndarray) -> int:
    p98 = np.ascontiguousarray(p98, np.uint8)
    return int(lib().ko_dhash_from_plane(_p(p98, C.c_uint8)))


def hamming_join(hashes: np.ndarray, threshold: int, *, require_band: bool = False, band_bits: int = 16,
                 band_count: int = 4, threads: int = 1, row_begin: int = 0, row_end: int | None = None):
    """All pairs i<j (i in [row_begin,row_end)) within the threshold, sorted by (i, j).

    Returns (i u32[m], j u32[m], dist u8[m])."""
    hashes = np.ascontiguousarray(hashes, np.uint64)
    n = hashes.shape[0]
    row_end = n if row_end is None else min(row_end, n)
    L = lib()

    def run(lo, hi):
        cap = 1 << 16
        while True:
            oi = np.empty(cap, np.uint32)
            oj = np.empty(cap, np.uint32)
            od = np.empty(cap, np.uint8)
            cnt = L.ko_hamming_join(_p(hashes, C.c_uint64), n, threshold, int(require_band), band_bits, band_count,
                                    lo, hi, _p(oi, C.c_uint32), _p(oj, C.c_uint32), _p(od, C.c_uint8), cap)
            if cnt <= cap:
                return oi[:cnt], oj[:cnt], od[:cnt]
            cap = int(cnt)

    # rows near the top of the triangle carry more pairs: use many small stripes
    rows = row_end - row_begin
    parts = [(row_begin + a, row_begin + b) for a, b in _split(rows, max(1, threads) * 16)] if rows > 0 else []
    if threads <= 1:
        res = [run(lo, hi) for lo, hi in parts]
    else:
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(lambda r: run(*r), parts))
    if not res:
        z = np.zeros(0, np.uint32)
        return z, z.copy(), np.zeros(0, np.uint8)
    oi = np.concatenate([r[0] for r in res])
    oj = np.concatenate([r[1] for r in res])
    od = np.concatenate([r[2] for r in res])
    order = np.lexsort((oj, oi))
    return oi[order], oj[order], od[order]


def ssim_u8(a: np.ndarray, b: np.ndarray, exact: bool = False) -> float:
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    if a.shape != b.shape or a.ndim != 2:
        raise ValueError("ssim_u8 wants two equal-shape 2-D uint8 arrays")
    f = lib().ko_ssim_u8_exact if exact else lib().ko_ssim_u8
    return float(f(_p(a, C.c_uint8), _p(b, C.c_uint8), a.shape[0], a.shape[1], a.shape[1], b.shape[1]))


def ssim_batch(bank: np.ndarray, ia: np.ndarray, ib: np.ndarray, *, exact: bool = False, threads: int = 1) -> np.ndarray:
    """bank: [m,h,w] uint8; pairs (ia[p], ib[p]) -> float64[p]."""
    bank = np.ascontiguousarray(bank, np.uint8)
    ia = np.ascontiguousarray(ia, np.int64)
    ib = np.ascontiguousarray(ib, np.int64)
    m, h, w = bank.shape
    out = np.zeros(ia.shape[0], np.float64)
    L = lib()

    def run(lo, hi):
        L.ko_ssim_batch(_p(bank, C.c_uint8), h, w, h * w, _p(ia[lo:hi], C.c_int64), _p(ib[lo:hi], C.c_int64),
                        hi - lo, _p(out[lo:hi], C.c_double), int(exact))

    parts = _split(ia.shape[0], threads * 4 if threads > 1 else 1)
    if threads <= 1:
        for lo, hi in parts:
            run(lo, hi)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda r: run(*r), parts))
    return out
