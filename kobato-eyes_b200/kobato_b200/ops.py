"""Array-level entry points over the C ABI (include/kobato_b200.h).

Inputs may be CUDA ``torch`` tensors (zero-copy: only ``data_ptr()`` and the current stream are
handed to the library) or host ``numpy`` arrays (the library's ``*_host`` entry points move the
data through pinned staging).  No path computes on the CPU; if the CUDA library or device is
missing these functions raise ``KobatoNativeError``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat

U64 = (1 << 64) - 1


def _torch():
    import torch

    return torch


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


def _stream_ptr(device_index: int) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(device_index).cuda_stream)


def _np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------- K1


def phash_dhash_batch(images, *, want_margin: bool = False, want_planes: bool = False):
    """pHash and dHash of a batch of decoded images (reference src/sig/phash.py:33-57).

    images: uint8 ``[n,h,w]`` ('L') or ``[n,h,w,c]`` with c in {1,3,4}; CUDA tensor or numpy array.
    Returns ``(phash, dhash)`` as **signed** int64 (the reference's ``_to_signed`` wrap, ready for
    SQLite), plus ``margin`` (float32, min |coef-mean|) and ``(plane32, plane9x8)`` on request.
    CUDA tensors in -> CUDA tensors out; numpy in -> numpy out.
    """
    lib = nat.load()
    if _is_tensor(images):
        torch = _torch()
        if not images.is_cuda or images.dtype != torch.uint8:
            raise ValueError("phash_dhash_batch wants a CUDA uint8 tensor (or a numpy array)")
        x = images if images.dim() == 4 else images.unsqueeze(-1)
        if x.dim() != 4:
            raise ValueError("images must be [n,h,w] or [n,h,w,c]")
        n, h, w, c = x.shape
        if x.stride(3) != 1 or x.stride(2) != c:
            x = x.contiguous()
        dev = x.device.index
        ctx = nat.context(dev)
        ph = torch.empty(n, dtype=torch.int64, device=x.device)
        dh = torch.empty(n, dtype=torch.int64, device=x.device)
        mg = torch.empty(n, dtype=torch.float32, device=x.device) if want_margin else None
        p32 = torch.empty((n, 32, 32), dtype=torch.uint8, device=x.device) if want_planes else None
        p98 = torch.empty((n, 8, 9), dtype=torch.uint8, device=x.device) if want_planes else None
        if n:
            with ctx.lock:
                nat.check(lib.ke_phash_batch(ctx.handle, x.data_ptr(), n, h, w, c, x.stride(0), x.stride(1),
                                             ph.data_ptr(), dh.data_ptr(), mg.data_ptr() if want_margin else None,
                                             p32.data_ptr() if want_planes else None,
                                             p98.data_ptr() if want_planes else None, _stream_ptr(dev)),
                          "ke_phash_batch")
        out = [ph, dh]
    else:
        x = np.ascontiguousarray(images, dtype=np.uint8)
        if x.ndim == 3:
            x = x[..., None]
        if x.ndim != 4:
            raise ValueError("images must be [n,h,w] or [n,h,w,c]")
        n, h, w, c = x.shape
        if want_planes:
            raise ValueError("planes are only returned for device tensors")
        ctx = nat.group()  # every device of the process-wide context takes a contiguous range of the images
        ph = np.empty(n, np.int64)
        dh = np.empty(n, np.int64)
        mg = np.empty(n, np.float32) if want_margin else None
        if n:
            with ctx.lock:
                nat.check(lib.ke_phash_batch_host(ctx.handle, _np_ptr(x), n, h, w, c, _np_ptr(ph), _np_ptr(dh),
                                                  _np_ptr(mg) if want_margin else None), "ke_phash_batch_host")
        out = [ph, dh]
        p32 = p98 = None
    if want_margin:
        out.append(mg)
    if want_planes:
        out.append((p32, p98))
    return tuple(out)


# ----------------------------------------------------------------------------- K2


def hamming_join(hashes, threshold: int, *, require_band: bool = False, band_bits: int = 16, band_count: int = 4,
                 band_allow=None, part_index: int = 0, part_count: int = 1, capacity: int | None = None):
    """All pairs i<j with popcount(h[i]^h[j]) <= threshold (reference src/dup/scanner.py:262-290).

    hashes: int64/uint64 table, CUDA tensor or numpy array.  Returns ``(i, j, dist)`` numpy arrays
    (uint32, uint32, uint8) sorted by (i, j).  The output buffer grows and the join re-runs when
    more pairs qualify than ``capacity`` (nothing is truncated)."""
    lib = nat.load()
    flags = nat.KE_JOIN_REQUIRE_BAND if require_band else 0
    if _is_tensor(hashes):
        torch = _torch()
        if not hashes.is_cuda or hashes.dtype not in (torch.int64, torch.uint64):
            raise ValueError("hamming_join wants a CUDA int64/uint64 tensor (or a numpy array)")
        h = hashes.contiguous().view(-1)
        n = h.numel()
        dev = h.device.index
        ctx = nat.context(dev)
        allow = None
        if band_allow is not None:
            allow = band_allow.contiguous() if _is_tensor(band_allow) else \
                torch.from_numpy(np.ascontiguousarray(band_allow, np.uint64).view(np.int64)).to(h.device)
        cap = int(capacity) if capacity is not None else max(1 << 16, 4 * n)
        while True:
            oi = torch.empty(cap, dtype=torch.int32, device=h.device)
            oj = torch.empty(cap, dtype=torch.int32, device=h.device)
            od = torch.empty(cap, dtype=torch.uint8, device=h.device)
            cnt = torch.zeros(1, dtype=torch.int64, device=h.device)
            with ctx.lock:
                nat.check(lib.ke_hamming_join(ctx.handle, h.data_ptr(), n, int(threshold), flags, band_bits, band_count,
                                              allow.data_ptr() if allow is not None else None, part_index, part_count,
                                              oi.data_ptr(), oj.data_ptr(), od.data_ptr(), cap, cnt.data_ptr(),
                                              _stream_ptr(dev)), "ke_hamming_join")
            total = int(cnt.item())
            if total <= cap:
                break
            cap = total
        ii = oi[:total].cpu().numpy().view(np.uint32)
        jj = oj[:total].cpu().numpy().view(np.uint32)
        dd = od[:total].cpu().numpy()
    else:
        h = np.ascontiguousarray(hashes).reshape(-1)
        if h.dtype not in (np.int64, np.uint64):
            raise ValueError("hashes must be int64 or uint64")
        n = h.shape[0]
        ctx = nat.group()  # the tiles of the triangle are dealt over every device of the process-wide context
        allow = np.ascontiguousarray(band_allow, np.uint64) if band_allow is not None else None
        cap = int(capacity) if capacity is not None else max(1 << 16, 4 * n)
        while True:
            ii = np.empty(cap, np.uint32)
            jj = np.empty(cap, np.uint32)
            dd = np.empty(cap, np.uint8)
            total = C.c_int64(0)
            with ctx.lock:
                st = lib.ke_hamming_join_host(ctx.handle, _np_ptr(h), n, int(threshold), flags, band_bits, band_count,
                                              _np_ptr(allow) if allow is not None else None, part_index, part_count,
                                              _np_ptr(ii), _np_ptr(jj), _np_ptr(dd), cap, C.byref(total))
            if st == nat.KE_E_CAPACITY:
                cap = int(total.value)
                continue
            nat.check(st, "ke_hamming_join_host")
            break
        ii, jj, dd = ii[: total.value], jj[: total.value], dd[: total.value]
    order = np.lexsort((jj, ii))
    return ii[order], jj[order], dd[order]


def hamming_join_device(hashes, threshold: int, *, require_band: bool = False, band_bits: int = 16, band_count: int = 4,
                        band_allow=None, part_index: int = 0, part_count: int = 1, capacity: int | None = None):
    """Like ``hamming_join`` for a CUDA table, but the (unsorted) result stays on the device:
    returns int32 ``i``, int32 ``j`` (bit patterns of the uint32 indices) and uint8 ``dist`` tensors."""
    torch = _torch()
    lib = nat.load()
    if not (_is_tensor(hashes) and hashes.is_cuda and hashes.dtype in (torch.int64, torch.uint64)):
        raise ValueError("hamming_join_device wants a CUDA int64/uint64 tensor")
    h = hashes.contiguous().view(-1)
    n = h.numel()
    dev = h.device.index
    ctx = nat.context(dev)
    flags = nat.KE_JOIN_REQUIRE_BAND if require_band else 0
    allow = None
    if band_allow is not None:
        allow = band_allow.contiguous() if _is_tensor(band_allow) else \
            torch.from_numpy(np.ascontiguousarray(band_allow, np.uint64).view(np.int64)).to(h.device)
    cap = int(capacity) if capacity is not None else max(1 << 16, 4 * n)
    while True:
        oi = torch.empty(cap, dtype=torch.int32, device=h.device)
        oj = torch.empty(cap, dtype=torch.int32, device=h.device)
        od = torch.empty(cap, dtype=torch.uint8, device=h.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=h.device)
        with ctx.lock:
            nat.check(lib.ke_hamming_join(ctx.handle, h.data_ptr(), n, int(threshold), flags, band_bits, band_count,
                                          allow.data_ptr() if allow is not None else None, part_index, part_count,
                                          oi.data_ptr(), oj.data_ptr(), od.data_ptr(), cap, cnt.data_ptr(),
                                          _stream_ptr(dev)), "ke_hamming_join")
        total = int(cnt.item())
        if total <= cap:
            return oi[:total], oj[:total], od[:total]
        cap = total


# ----------------------------------------------------------------------------- K3


def ssim_batch(bank, ia, ib, *, gaussian: bool = False, check: bool = True):
    """SSIM of pairs (bank[ia[p]], bank[ib[p]]) — reference src/dup/refine.py:52 semantics.

    bank: CUDA uint8 tensor ``[m,h,w]`` ('L' planes) or ``[m,h,w,c]`` (RGB/RGBA, Pillow luma applied
    on the fly); ia/ib: index sequences.  Returns a float64 CUDA tensor ``[n_pairs]``.
    ``gaussian=True`` is skimage's ``gaussian_weights=True`` window (not the reference's path).  ``check=False`` skips the
    range check of DEVICE index tensors (two reductions and a synchronisation; for indices the caller derived itself)."""
    torch = _torch()
    lib = nat.load()
    if not (_is_tensor(bank) and bank.is_cuda and bank.dtype == torch.uint8):
        raise ValueError("ssim_batch wants a CUDA uint8 bank tensor")
    x = bank if bank.dim() == 4 else bank.unsqueeze(-1)
    m, h, w, c = x.shape
    if x.stride(3) != 1 or x.stride(2) != c:
        x = x.contiguous()
    dev = x.device.index
    ia_t, ib_t, n = _pair_index(ia, ib, m, x.device, check)
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    if n:
        ctx = nat.context(dev)
        with ctx.lock:
            st = lib.ke_ssim_batch(ctx.handle, x.data_ptr(), h, w, c, x.stride(0), x.stride(1), ia_t.data_ptr(),
                                   ib_t.data_ptr(), n, 1 if gaussian else 0, out.data_ptr(), _stream_ptr(dev))
        if st == nat.KE_E_UNSUPPORTED:
            raise ValueError(nat.last_error())  # the reference (skimage) raises ValueError here
        nat.check(st, "ke_ssim_batch")
    return out


def ssim_pairs(a, b, *, gaussian: bool = False) -> np.ndarray:
    """SSIM of host images: a, b uint8 ``[n,h,w]`` 'L' planes (or one ``[h,w]`` pair) or
    ``[n,h,w,c]`` RGB(A) (Pillow luma applied on the GPU) -> float64 ``[n]``."""
    lib = nat.load()
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    if a.ndim == 2:
        a, b = a[None], b[None]
    if a.ndim == 3:
        a, b = a[..., None], b[..., None]
    n, h, w, c = a.shape
    out = np.empty(n, np.float64)
    if n:
        ctx = nat.group()
        with ctx.lock:
            st = lib.ke_ssim_pairs_host(ctx.handle, _np_ptr(a), _np_ptr(b), n, h, w, c, 1 if gaussian else 0, _np_ptr(out))
        if st == nat.KE_E_UNSUPPORTED:
            raise ValueError(nat.last_error())
        nat.check(st, "ke_ssim_pairs_host")
    return out


# ----------------------------------------------------------------------------- helpers


# ----------------------------------------------------------------------------- N1 (tile aHash / pixel MAE refinement)

_FILTERS = {"lanczos": 1, "bilinear": 2, 1: 1, 2: 2}


def _as_cuda_u8(images):
    torch = _torch()
    if _is_tensor(images):
        if not images.is_cuda or images.dtype != torch.uint8:
            raise ValueError("expected a CUDA uint8 tensor (or a numpy array)")
        x = images
    else:
        x = torch.from_numpy(np.ascontiguousarray(images, dtype=np.uint8)).cuda()
    if x.dim() == 3:
        x = x.unsqueeze(-1)
    if x.dim() != 4 or x.shape[3] not in (1, 3, 4):
        raise ValueError("images must be [n,h,w] or [n,h,w,c] with c in {1,3,4}")
    if x.stride(3) != 1 or x.stride(2) != x.shape[3]:
        x = x.contiguous()
    return x


def gray_resize_batch(images, out_w: int, out_h: int, filter="bilinear"):
    """``convert("L").resize((out_w, out_h), filter)`` for a batch of decoded images, byte-identical to Pillow
    (reference src/ui/dup_refine_parallel.py:66-69, :203-207).  Returns a CUDA uint8 tensor ``[n,out_h,out_w]``."""
    torch = _torch()
    lib = nat.load()
    if filter not in _FILTERS:
        raise ValueError("filter must be 'bilinear' or 'lanczos'")
    x = _as_cuda_u8(images)
    n, h, w, c = x.shape
    out_w, out_h = int(out_w), int(out_h)
    if out_w <= 0 or out_h <= 0:
        raise ValueError("height and width must be > 0")  # Pillow's message for resize((0, n))
    dev = x.device.index
    mid = torch.empty((n, h, out_w), dtype=torch.uint8, device=x.device)
    out = torch.empty((n, out_h, out_w), dtype=torch.uint8, device=x.device)
    if n:
        ctx = nat.context(dev)
        with ctx.lock:
            nat.check(lib.ke_gray_resize_batch(ctx.handle, x.data_ptr(), n, h, w, c, x.stride(0), x.stride(1), out_w, out_h,
                                               _FILTERS[filter], mid.data_ptr(), out.data_ptr(), _stream_ptr(dev)),
                      "ke_gray_resize_batch")
    return out


def tile_ahash_bits(planes, grid: int = 4, tile: int = 8):
    """Tile-mean threshold bits of ``[n, grid*tile, grid*tile]`` uint8 planes (reference :71-83): returns an int32
    CUDA tensor ``[n, words]`` whose little-endian bit string is the reference's packed integer."""
    torch = _torch()
    lib = nat.load()
    x = planes if _is_tensor(planes) else torch.from_numpy(np.ascontiguousarray(planes, np.uint8)).cuda()
    side = int(grid) * int(tile)
    if x.dim() != 3 or x.shape[1] != side or x.shape[2] != side or x.dtype != torch.uint8:
        raise ValueError(f"planes must be uint8 [n,{side},{side}]")
    x = x.contiguous()
    n = x.shape[0]
    words = (side * side + 31) // 32
    bits = torch.zeros((n, words), dtype=torch.int32, device=x.device)
    if n:
        dev = x.device.index
        ctx = nat.context(dev)
        with ctx.lock:
            nat.check(lib.ke_tile_ahash_bits(ctx.handle, x.data_ptr(), n, int(grid), int(tile), bits.data_ptr(),
                                             _stream_ptr(dev)), "ke_tile_ahash_bits")
    return bits


def bits_to_ints(bits) -> list[int]:
    """``[n, words]`` int32 bit words -> Python ints (little endian), the reference's tile_ahash_bits values."""
    arr = bits.cpu().numpy() if _is_tensor(bits) else np.asarray(bits)
    raw = np.ascontiguousarray(arr.astype("<i4", copy=False)).view(np.uint8).reshape(arr.shape[0], -1)
    return [int.from_bytes(row.tobytes(), "little") for row in raw]


def _pair_index(ia, ib, m: int, device, check: bool = True):
    """Pair index sequences -> int64 device tensors, range-checked.  Host sequences (lists, numpy arrays) are checked on
    the host before they are uploaded; only device tensors cost a device reduction and a synchronisation."""
    torch = _torch()
    if not (_is_tensor(ia) and ia.is_cuda) and not (_is_tensor(ib) and ib.is_cuda):
        a = np.ascontiguousarray(ia.cpu().numpy() if _is_tensor(ia) else ia, np.int64)
        b = np.ascontiguousarray(ib.cpu().numpy() if _is_tensor(ib) else ib, np.int64)
        if a.shape != b.shape or a.ndim != 1:
            raise ValueError("ia and ib must be 1-D and of equal length")
        if a.size and (max(int(a.max()), int(b.max())) >= m or min(int(a.min()), int(b.min())) < 0):
            raise ValueError("pair index out of range")
        both = torch.from_numpy(np.stack([a, b])).to(device)  # one upload
        return both[0], both[1], int(a.size)
    ia_t = torch.as_tensor(ia, dtype=torch.int64).to(device).contiguous()
    ib_t = torch.as_tensor(ib, dtype=torch.int64).to(device).contiguous()
    if ia_t.shape != ib_t.shape or ia_t.dim() != 1:
        raise ValueError("ia and ib must be 1-D and of equal length")
    n = ia_t.numel()
    if check and n and (int(torch.max(torch.maximum(ia_t, ib_t))) >= m or int(torch.min(torch.minimum(ia_t, ib_t))) < 0):
        raise ValueError("pair index out of range")
    return ia_t, ib_t, n


def bits_hamming_pairs(bits, ia, ib):
    """popcount(bits[ia] ^ bits[ib]) per pair (reference tile_hamming :86-88) -> int32 CUDA tensor."""
    torch = _torch()
    lib = nat.load()
    if not (_is_tensor(bits) and bits.is_cuda and bits.dtype == torch.int32 and bits.dim() == 2):
        raise ValueError("bits must be the int32 CUDA tensor returned by tile_ahash_bits")
    b = bits.contiguous()
    ia_t, ib_t, n = _pair_index(ia, ib, b.shape[0], b.device)
    out = torch.empty(n, dtype=torch.int32, device=b.device)
    if n:
        dev = b.device.index
        ctx = nat.context(dev)
        with ctx.lock:
            nat.check(lib.ke_bits_hamming_pairs(ctx.handle, b.data_ptr(), b.shape[1], ia_t.data_ptr(), ib_t.data_ptr(), n,
                                                out.data_ptr(), _stream_ptr(dev)), "ke_bits_hamming_pairs")
    return out


def plane_sad_pairs(planes, ia, ib):
    """sum |planes[ia] - planes[ib]| per pair -> int64 CUDA tensor; reference ``_mae01`` (:210-212) is
    ``sad / plane.size / 255.0``."""
    torch = _torch()
    lib = nat.load()
    if not (_is_tensor(planes) and planes.is_cuda and planes.dtype == torch.uint8 and planes.dim() >= 2):
        raise ValueError("planes must be a CUDA uint8 tensor [m, ...]")
    x = planes.contiguous()
    plane_bytes = int(x[0].numel()) if x.shape[0] else 1
    ia_t, ib_t, n = _pair_index(ia, ib, x.shape[0], x.device)
    out = torch.empty(n, dtype=torch.int64, device=x.device)
    if n:
        dev = x.device.index
        ctx = nat.context(dev)
        with ctx.lock:
            nat.check(lib.ke_plane_sad_pairs(ctx.handle, x.data_ptr(), plane_bytes, ia_t.data_ptr(), ib_t.data_ptr(), n,
                                             out.data_ptr(), _stream_ptr(dev)), "ke_plane_sad_pairs")
    return out


def cluster_pairs_csr(a, b):
    """Connected components of the pairs (a[k], b[k]) in CSR form: ``(members, offsets)`` with component c =
    ``members[offsets[c]:offsets[c+1]]``, components by ascending representative (= smallest id, "smaller root wins",
    reference src/dup/cluster.py:22-70), members ascending.  Host union-find inside the library."""
    a = np.ascontiguousarray(a, np.int64)
    b = np.ascontiguousarray(b, np.int64)
    if a.shape != b.shape or a.ndim != 1:
        raise ValueError("a and b must be 1-D and of equal length")
    if a.size == 0:
        return np.empty(0, np.int64), np.zeros(1, np.int64)
    nodes = np.empty(2 * a.size, np.int64)
    reps = np.empty(2 * a.size, np.int64)
    count = C.c_int64(0)
    nat.check(nat.load().ke_cluster_pairs_host(_np_ptr(a), _np_ptr(b), a.size, _np_ptr(nodes), _np_ptr(reps), C.byref(count)),
              "ke_cluster_pairs_host")
    nodes, reps = nodes[: count.value], reps[: count.value]
    cuts = np.flatnonzero(np.diff(reps)) + 1
    offsets = np.concatenate([[0], cuts, [nodes.size]]).astype(np.int64)
    return nodes, offsets


def cluster_pairs(a, b):
    """``cluster_pairs_csr`` as ``[(representative, [members])]`` sorted by representative."""
    members, offsets = cluster_pairs_csr(a, b)
    flat, offs = members.tolist(), offsets.tolist()
    return [(flat[s], flat[s:e]) for s, e in zip(offs[:-1], offs[1:])]


def synth_images_device(start: int, count: int, h: int, w: int, c: int = 3, *, n_set: int = 1 << 30,
                        seed: int | None = None, planted: float = 0.05, device=None, out=None, stride: int = 1):
    """CUDA twin of ``synth.synth_image`` (identical bytes) -> uint8 CUDA tensor [count,h,w,c] holding the images
    ``start, start + stride, ...`` of the set."""
    from . import synth

    torch = _torch()
    lib = nat.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if out is None:
        out = torch.empty((count, h, w, c), dtype=torch.uint8, device=dev)
    ctx = nat.context(dev.index)
    with ctx.lock:
        nat.check(lib.ke_synth_images(ctx.handle, out.data_ptr(), start, int(stride), count, h, w, c, n_set,
                                      (synth.SEED if seed is None else seed) & U64, int(round(planted * 1000)),
                                      _stream_ptr(dev.index)), "ke_synth_images")
    return out if c > 1 else out[..., 0]


def orb_match_pairs(desc_a: list, desc_b: list, *, want_matches: bool = False):
    """``len(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db))`` for a batch of descriptor-set pairs
    (reference src/dup/refine.py:64-67): ``desc_a[p]`` / ``desc_b[p]`` are uint8 ``[n, 32]`` ORB descriptor arrays (or
    None / empty -> 0 matches).  Returns an int32 numpy array ``[n_pairs]``; with ``want_matches`` also the list of
    ``(queryIdx, trainIdx, distance)`` triples per pair, sorted by queryIdx."""
    torch = _torch()
    lib = nat.load()
    n_pairs = len(desc_a)
    if len(desc_b) != n_pairs:
        raise ValueError("desc_a and desc_b must have the same length")
    counts = np.zeros(n_pairs, np.int32)
    live, rows, off_a, off_b, cnt_a, cnt_b = [], [], [], [], [], []
    at = 0
    for p, (da, db) in enumerate(zip(desc_a, desc_b)):
        if da is None or db is None or len(da) == 0 or len(db) == 0:
            continue
        da = np.ascontiguousarray(da, np.uint8)
        db = np.ascontiguousarray(db, np.uint8)
        if da.ndim != 2 or db.ndim != 2 or da.shape[1] != 32 or db.shape[1] != 32:
            raise ValueError("descriptors must be uint8 [n, 32]")
        live.append(p)
        off_a.append(at), cnt_a.append(len(da))
        at += len(da)
        off_b.append(at), cnt_b.append(len(db))
        at += len(db)
        rows += [da, db]
    matches = [[] for _ in range(n_pairs)] if want_matches else None
    if not live:
        return (counts, matches) if want_matches else counts
    dev = torch.device("cuda", torch.cuda.current_device())
    d_desc = torch.from_numpy(np.concatenate(rows)).to(dev)
    t_off_a = torch.tensor(off_a, dtype=torch.int64, device=dev)
    t_off_b = torch.tensor(off_b, dtype=torch.int64, device=dev)
    t_cnt_a = torch.tensor(cnt_a, dtype=torch.int32, device=dev)
    t_cnt_b = torch.tensor(cnt_b, dtype=torch.int32, device=dev)
    max_a, max_b = max(cnt_a), max(cnt_b)
    out = torch.empty(len(live), dtype=torch.int32, device=dev)
    m_train = torch.empty((len(live), max_a), dtype=torch.int32, device=dev) if want_matches else None
    m_dist = torch.empty((len(live), max_a), dtype=torch.int32, device=dev) if want_matches else None
    ctx = nat.context(dev.index)
    with ctx.lock:
        nat.check(lib.ke_orb_match_pairs(ctx.handle, d_desc.data_ptr(), t_off_a.data_ptr(), t_cnt_a.data_ptr(),
                                         t_off_b.data_ptr(), t_cnt_b.data_ptr(), len(live), max_a, max_b, out.data_ptr(),
                                         m_train.data_ptr() if want_matches else None,
                                         m_dist.data_ptr() if want_matches else None, _stream_ptr(dev.index)),
                  "ke_orb_match_pairs")
    counts[np.asarray(live)] = out.cpu().numpy()
    if want_matches:
        tr, ds = m_train.cpu().numpy(), m_dist.cpu().numpy()
        for k, p in enumerate(live):
            q = np.flatnonzero(tr[k, :cnt_a[k]] >= 0)
            matches[p] = [(int(i), int(tr[k, i]), int(ds[k, i])) for i in q]
        return counts, matches
    return counts


def luma_planes(bank, idx, out=None, check: bool = True):
    """``convert("L")`` planes of ``bank[idx]`` (Pillow rgb2l; reference src/dup/refine.py:48-49) -> uint8 CUDA tensor
    ``[len(idx), h, w]`` (written into ``out`` — contiguous, ``len(idx) * h * w`` bytes — when given)."""
    torch = _torch()
    lib = nat.load()
    x = _as_cuda_u8(bank)
    m, h, w, c = x.shape
    idx_t = torch.as_tensor(idx, dtype=torch.int64).to(x.device).contiguous()
    n = idx_t.numel()
    if check and n and (int(idx_t.max()) >= m or int(idx_t.min()) < 0):
        raise ValueError("image index out of range")
    if out is None:
        out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    elif not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == n * h * w):
        raise ValueError("out must be a contiguous CUDA uint8 tensor of len(idx) * h * w bytes")
    if n:
        dev = x.device.index
        ctx = nat.context(dev)
        with ctx.lock:
            nat.check(lib.ke_luma_planes(ctx.handle, x.data_ptr(), h, w, c, x.stride(0), x.stride(1), idx_t.data_ptr(), n,
                                         out.data_ptr(), _stream_ptr(dev)), "ke_luma_planes")
    return out


def cluster_pairs_device(a, b, n_nodes: int):
    """Component label (= smallest member index) of every node that occurs in a pair, -1 for the others: the device
    union-find (``ke_cluster_pairs``) over CUDA int32/uint32 index tensors.  Returns an int64 CUDA tensor ``[n_nodes]``."""
    torch = _torch()
    lib = nat.load()
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    if not (a.is_cuda and b.is_cuda) or a.shape != b.shape or a.dim() != 1:
        raise ValueError("cluster_pairs_device wants two 1-D CUDA index tensors of equal length")
    a32 = a.to(torch.int32).contiguous()
    b32 = b.to(torch.int32).contiguous()
    label = torch.empty(int(n_nodes), dtype=torch.int32, device=a.device)
    dev = a.device.index
    ctx = nat.context(dev)
    with ctx.lock:
        nat.check(lib.ke_cluster_pairs(ctx.handle, a32.data_ptr(), b32.data_ptr(), a32.numel(), int(n_nodes), label.data_ptr(),
                                       _stream_ptr(dev)), "ke_cluster_pairs")
    return label.to(torch.int64)  # 0xFFFFFFFF reads back as -1


def scan_table(phash, file_id=None, size=None, *, threshold: int = 8, band_bits: int = 16, band_count: int = 4,
               size_ratio: float | None = None, pair_cap: int | None = None, want_edges: bool = False) -> dict:
    """Table-level duplicate scan (``ke_scan_table_host``): the columns of ``iter_files_for_dup`` (reference
    src/db/repository.py:416-455) -> members grouped by component.

    phash: int64 (SQLite's signed ``phash_u64``) or uint64 array; file_id / size: int64 arrays or None.
    Returns ``{"index", "label", "best", "offsets", "stats"}`` (+ ``"edges": (i, j, dist)``): component c is
    ``index[offsets[c]:offsets[c+1]]`` (table rows, ascending), its label the smallest row, ``best`` the reference's
    best_hamming per member (src/dup/scanner.py:304-313).  File ids must be distinct."""
    lib = nat.load()
    ph = np.ascontiguousarray(phash).reshape(-1)
    if ph.dtype == np.uint64:
        ph = ph.view(np.int64)
    if ph.dtype != np.int64:
        raise ValueError("phash must be int64 or uint64")
    n = ph.shape[0]
    fid = None if file_id is None else np.ascontiguousarray(file_id, np.int64).reshape(-1)
    sz = None if size is None else np.ascontiguousarray(size, np.int64).reshape(-1)
    if (fid is not None and fid.shape[0] != n) or (sz is not None and sz.shape[0] != n):
        raise ValueError("columns must have the same length")
    ctx = nat.group()
    stats = nat.ScanStats()
    mcap, ecap = max(1 << 16, n // 8), (max(1 << 16, n // 8) if want_edges else 0)
    while True:
        mi = np.empty(mcap, np.int64)
        ml = np.empty(mcap, np.int64)
        mb = np.empty(mcap, np.int32)
        ei = np.empty(ecap, np.uint32)
        ej = np.empty(ecap, np.uint32)
        ed = np.empty(ecap, np.uint8)
        with ctx.lock:
            st = lib.ke_scan_table_host(ctx.handle, _np_ptr(ph), _np_ptr(fid) if fid is not None else None,
                                        _np_ptr(sz) if sz is not None else None, n, int(threshold), int(band_bits),
                                        int(band_count), float(size_ratio) if size_ratio else 0.0,
                                        int(pair_cap) if pair_cap else 0, _np_ptr(mi), _np_ptr(ml), _np_ptr(mb), mcap,
                                        _np_ptr(ei) if ecap else None, _np_ptr(ej) if ecap else None,
                                        _np_ptr(ed) if ecap else None, ecap, C.byref(stats))
        if st == nat.KE_E_CAPACITY:
            mcap = max(mcap, int(stats.members))
            if want_edges:
                ecap = max(ecap, int(stats.edges))
            continue
        nat.check(st, "ke_scan_table_host")
        break
    m = int(stats.members)
    mi, ml, mb = mi[:m], ml[:m], mb[:m]
    cuts = np.flatnonzero(np.diff(ml)) + 1 if m else np.zeros(0, np.int64)
    offsets = np.concatenate([[0], cuts, [m]]).astype(np.int64) if m else np.zeros(1, np.int64)
    out = {"index": mi, "label": ml, "best": mb, "offsets": offsets, "stats": stats.as_dict()}
    if want_edges:
        e = int(stats.edges)
        out["edges"] = (ei[:e], ej[:e], ed[:e])
    return out


def popc_rate(iters: int = 4096, device: int | None = None):
    """(POPC thread-instructions per SM clock per SM, SM clock MHz) — K2's roofline denominator."""
    lib = nat.load()
    ctx = nat.context(device)
    rate, mhz = C.c_double(), C.c_double()
    with ctx.lock:
        nat.check(lib.ke_microbench_popc(ctx.handle, iters, C.byref(rate), C.byref(mhz)), "ke_microbench_popc")
    return rate.value, mhz.value
