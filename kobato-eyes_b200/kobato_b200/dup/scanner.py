"""Drop-in for the reference's ``dup.scanner`` (src/dup/scanner.py) with the candidate search
on the GPU.

Public surface and behaviour follow the reference: ``DuplicateFile`` (+ ``from_row`` accepting int /
BLOB / hex / decimal hashes, :44-117), ``DuplicateScanConfig`` (validation and ``ValueError``s,
:147-166), ``DuplicateScanner.build_clusters`` (:211-356), ``DuplicateCluster(Entry)``.
``DuplicateScanner`` can be handed to the reference UI through its sanctioned seam,
``DupViewModel(scanner_factory=DuplicateScanner)`` (src/ui/viewmodels/dup_view_model.py:31-72); it
accepts the reference's own ``DuplicateFile`` objects (anything with the same attributes) and
returns them inside the cluster entries.

How the work is split:
  * the reference's per-bucket double loop (:262-290) is replaced by ONE all-pairs Hamming join
    on the GPU(s) (``ke_hamming_join`` with the band-equality predicate, which makes the candidate
    set identical to the LSH buckets' — see tests/test_oracle_pinned.py);
  * the table path (default; ``ke_scan_table_host``): bucket statistics, the
    ``KE_DUP_BUCKET_PAIR_CAP`` mask (:239-266), the same-id / size-ratio gates (:271-279) and the DSU +
    ``best_hamming`` (:304-318) all run on the device over COLUMN ARRAYS; Python only builds the
    ``DuplicateCluster`` objects of the members (keeper choice and the sorts, :320-356).
    ``build_clusters_from_columns`` takes the SQLite columns directly (no per-row objects at all);
  * the legacy path (cosine gate in use, repeated file ids, or an injected ``join``): per-candidate
    filters, edge de-duplication, DSU on the host over the (few) surviving pairs.
"""
from __future__ import annotations

import logging
import math
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, Mapping, Sequence

import numpy as np

logger = logging.getLogger(__name__)

U64 = (1 << 64) - 1
PHASH_KEYS = ("phash_u64", "phash", "phash64", "phash_hex", "phash_bytes", "signature", "sig")
_EXTENSION_PRIORITY = {"png": 4, "apng": 4, "webp": 3, "tiff": 2, "tif": 2, "bmp": 1, "gif": 1,
                       "jpeg": 0, "jpg": 0, "jpe": 0, "jfif": 0}


def _row_get(row, key, default=None):
    try:
        if isinstance(row, dict):
            return row.get(key, default)
        if hasattr(row, "keys") and key in row.keys():
            return row[key]
        return getattr(row, key, default)
    except (AttributeError, KeyError, TypeError):
        return default


def _parse_phash_any(raw) -> int | None:
    """int / numpy int / Decimal / BLOB (big endian) / decimal-or-hex string -> unsigned 64-bit, or
    None for anything unparseable (the row is then skipped by ``from_row``)."""
    if raw is None:
        return None
    if isinstance(raw, (bytes, bytearray, memoryview)):
        try:
            return int.from_bytes(bytes(raw), "big", signed=False) & U64
        except (OverflowError, TypeError, ValueError):
            return None
    if isinstance(raw, str):
        text = raw.strip()
        if not text:
            return None
        for base in (0, 16):
            try:
                return int(text, base) & U64
            except ValueError:
                continue
        return None
    try:
        return int(raw) & U64
    except (OverflowError, TypeError, ValueError):
        return None


def _opt_int(v):
    return int(v) if isinstance(v, (int, float)) else None


@dataclass(frozen=True)
class DuplicateFile:
    """Metadata required for duplicate detection of a single file."""

    file_id: int
    path: Path
    size: int | None
    width: int | None
    height: int | None
    phash: int
    embedding: tuple[float, ...] | None = None

    @classmethod
    def from_row(cls, row: Mapping[str, object]) -> "DuplicateFile":
        raw = None
        for key in PHASH_KEYS:
            raw = _row_get(row, key, None)
            if raw is not None:
                break
        ph = _parse_phash_any(raw)
        if ph is None:
            raise ValueError("Row is missing perceptual hash information")
        return cls(
            file_id=int(_row_get(row, "file_id", _row_get(row, "id", -1))),
            path=Path(str(_row_get(row, "path", _row_get(row, "file_path", "")))),
            size=_opt_int(_row_get(row, "size")),
            width=_opt_int(_row_get(row, "width")),
            height=_opt_int(_row_get(row, "height")),
            phash=ph,
        )

    @property
    def resolution(self) -> int:
        return (self.width or 0) * (self.height or 0)

    @property
    def extension_priority(self) -> int:
        return _EXTENSION_PRIORITY.get(self.path.suffix.lower().lstrip("."), 0)


@dataclass(frozen=True)
class DuplicateClusterEntry:
    file: DuplicateFile
    best_hamming: int | None


@dataclass(frozen=True)
class DuplicateCluster:
    files: list[DuplicateClusterEntry]
    keeper_id: int


@dataclass(frozen=True)
class DuplicateScanConfig:
    hamming_threshold: int = 8
    size_ratio: float | None = None
    band_bits: int = 16
    band_count: int = 4
    cosine_threshold: float | None = None

    def __post_init__(self) -> None:
        if self.band_bits <= 0:
            raise ValueError("band_bits must be positive")
        if self.band_count <= 0:
            raise ValueError("band_count must be positive")
        if self.hamming_threshold < 0 or self.hamming_threshold > 64:
            raise ValueError("hamming_threshold must be in [0, 64]")
        if self.cosine_threshold is not None and not (-1.0 <= self.cosine_threshold <= 1.0):
            raise ValueError("cosine_threshold must be between -1.0 and 1.0")


@dataclass
class DuplicateEdge:
    file_id_a: int
    file_id_b: int
    hamming: int | None


def _safe_positive_int(value: str | None) -> int | None:
    if value is None or not value.strip():
        return None
    try:
        parsed = int(value)
    except ValueError:
        return None
    return parsed if parsed > 0 else None


def _resolution(f) -> int:
    return (getattr(f, "width", None) or 0) * (getattr(f, "height", None) or 0)


def _ext_priority(f) -> int:
    return _EXTENSION_PRIORITY.get(Path(f.path).suffix.lower().lstrip("."), 0)


class _Bands:
    """Band values of a hash table, bucket statistics and the pair-cap mask (NumPy, host)."""

    def __init__(self, hashes: np.ndarray, band_bits: int, band_count: int):
        self.bits, self.count = band_bits, band_count
        mask = np.uint64((1 << band_bits) - 1) if band_bits < 64 else np.uint64(U64)
        self.values = [(hashes >> np.uint64(b * band_bits)) & mask for b in range(band_count)]
        self.first = []   # per band: list position of the first member of my bucket
        self.sizes = []   # per band: size of my bucket
        for v in self.values:
            _, first_idx, inverse, counts = np.unique(v, return_index=True, return_inverse=True, return_counts=True)
            self.first.append(first_idx[inverse])
            self.sizes.append(counts[inverse])

    def stats(self):
        ge2 = sum(int(np.count_nonzero(np.unique(v, return_counts=True)[1] >= 2)) for v in self.values)
        n_buckets = sum(int(np.unique(v).size) for v in self.values)
        max_bucket = max((int(s.max()) for s in self.sizes), default=0)
        return n_buckets, ge2, max_bucket

    def allow_mask(self, pair_cap: int | None) -> np.ndarray | None:
        if pair_cap is None:
            return None
        allow = np.zeros(self.values[0].shape[0], np.uint64)
        for b, s in enumerate(self.sizes):
            s = s.astype(np.int64)
            ok = (s * (s - 1)) // 2 <= pair_cap
            allow |= ok.astype(np.uint64) << np.uint64(b)
        return allow


class DuplicateScanner:
    """Duplicate clusters from a GPU all-pairs Hamming join + host DSU (reference :203-415)."""

    def __init__(self, config: DuplicateScanConfig, *, join=None, scan_table=None) -> None:
        self._config = config
        assert config.band_bits * config.band_count <= 64, "band config too large"
        self._band_mask = (1 << config.band_bits) - 1
        self._join = join  # injectable for tests / multi-process joins (see kobato_b200.dist): selects the legacy path
        self._scan_table = scan_table  # injectable stand-in for ops.scan_table (tests)

    # -- candidate search ---------------------------------------------------------------

    def _gpu_join(self, hashes: np.ndarray, allow: np.ndarray | None):
        if self._join is not None:
            return self._join(hashes, self._config, allow)
        from .. import ops

        cfg = self._config
        return ops.hamming_join(hashes, cfg.hamming_threshold, require_band=True, band_bits=cfg.band_bits,
                                band_count=cfg.band_count, band_allow=allow)

    def find_edges(self, candidates: Sequence) -> dict[tuple[int, int], DuplicateEdge]:
        cfg = self._config
        n = len(candidates)
        hashes = np.fromiter((int(f.phash) & U64 for f in candidates), dtype=np.uint64, count=n)
        bands = _Bands(hashes, cfg.band_bits, cfg.band_count)
        n_buckets, ge2, max_bucket = bands.stats()
        max_pairs = (max_bucket * (max_bucket - 1)) // 2
        pair_cap = _safe_positive_int(os.environ.get("KE_DUP_BUCKET_PAIR_CAP"))
        logger.info("dup: buckets=%d (>=2:%d) max_bucket=%d max_bucket_pairs=%d pair_cap=%s", n_buckets, ge2,
                    max_bucket, max_pairs, pair_cap)
        if pair_cap is not None and max_pairs > pair_cap:
            logger.warning("dup: largest bucket has %d pair(s), above KE_DUP_BUCKET_PAIR_CAP=%d; large buckets "
                           "will be skipped", max_pairs, pair_cap)
        if ge2 == 0:
            logger.warning("dup: no bucket has 2+ items -> edges=0")
            return {}
        allow = bands.allow_mask(pair_cap)
        ii, jj, dd = self._gpu_join(hashes, allow)
        cand_total = len(ii)

        # The reference visits buckets in creation order (first member, band) and keeps the FIRST
        # edge per id pair; reproduce that order for the surviving candidates.
        if cand_total:
            x = hashes[ii] ^ hashes[jj]
            mask = np.uint64(self._band_mask) if cfg.band_bits < 64 else np.uint64(U64)
            visit = np.full(cand_total, np.iinfo(np.int64).max, np.int64)
            for b in range(cfg.band_count):
                eq = ((x >> np.uint64(b * cfg.band_bits)) & mask) == 0
                if allow is not None:
                    eq &= ((allow[ii] >> np.uint64(b)) & np.uint64(1)).astype(bool)
                key = bands.first[b][ii].astype(np.int64) * cfg.band_count + b
                visit = np.where(eq, np.minimum(visit, key), visit)
            order = np.lexsort((jj, ii, visit))
        else:
            order = np.zeros(0, np.int64)

        edges: dict[tuple[int, int], DuplicateEdge] = {}
        after_size = after_cos = 0
        for k in order:
            a, b = candidates[int(ii[k])], candidates[int(jj[k])]
            if a.file_id == b.file_id:
                continue
            if not self._passes_size_ratio(a, b):
                continue
            after_size += 1
            if not self._passes_cosine_similarity(a, b):
                continue
            after_cos += 1
            key = (a.file_id, b.file_id) if a.file_id < b.file_id else (b.file_id, a.file_id)
            if key not in edges:
                edges[key] = DuplicateEdge(a.file_id, b.file_id, int(dd[k]))
        logger.info("dup: gpu candidates (band & ham)=%d -> size=%d -> cosine=%d -> edges=%d", cand_total, after_size,
                    after_cos, len(edges))
        return edges

    # -- public API ---------------------------------------------------------------------

    def build_clusters(self, files: Iterable) -> list[DuplicateCluster]:
        cfg = self._config
        candidates = [f for f in files if f.phash is not None]
        logger.info("dup: candidates=%d band_bits=%d band_count=%d ham_th=%d size_ratio=%s cosine_th=%s",
                    len(candidates), cfg.band_bits, cfg.band_count, cfg.hamming_threshold, cfg.size_ratio,
                    cfg.cosine_threshold)
        if not candidates:
            return []
        if self._join is None and not self._cosine_in_use(candidates):
            n = len(candidates)
            ids = np.fromiter((int(f.file_id) for f in candidates), dtype=np.int64, count=n)
            if _all_distinct(ids):
                hashes = np.fromiter((int(f.phash) & U64 for f in candidates), dtype=np.uint64, count=n)
                sizes = np.fromiter((int(f.size or 0) for f in candidates), dtype=np.int64, count=n)
                scan = self._scan_columns(hashes, ids, sizes)
                return self._clusters_of(scan, candidates.__getitem__)
        edges = self.find_edges(candidates)
        if not edges:
            return []
        return assemble_clusters(candidates, edges)

    def build_clusters_from_columns(self, file_id, phash, size=None, *, make_file) -> list[DuplicateCluster]:
        """``build_clusters`` for callers that hold the rows of ``iter_files_for_dup`` (reference
        src/db/repository.py:416-455) as column arrays: ``file_id`` int64 (distinct), ``phash`` int64 — SQLite's signed
        ``phash_u64`` (src/db/schema.py:65-72) — or uint64, ``size`` int64 or None.  ``make_file(row_index)`` must
        return the ``DuplicateFile`` of a table row; it is called for the members of clusters only."""
        cfg = self._config
        if cfg.cosine_threshold is not None:
            raise ValueError("the cosine gate needs embeddings: use build_clusters(files)")
        ids = np.ascontiguousarray(file_id, np.int64).reshape(-1)
        if not _all_distinct(ids):
            raise ValueError("build_clusters_from_columns needs distinct file ids")
        logger.info("dup: candidates=%d band_bits=%d band_count=%d ham_th=%d size_ratio=%s cosine_th=%s", len(ids),
                    cfg.band_bits, cfg.band_count, cfg.hamming_threshold, cfg.size_ratio, cfg.cosine_threshold)
        if len(ids) == 0:
            return []
        return self._clusters_of(self._scan_columns(phash, ids, size), make_file)

    def _cosine_in_use(self, candidates: Sequence) -> bool:
        """The cosine gate (:372-400) can only drop a pair when a threshold is set AND embeddings exist."""
        if self._config.cosine_threshold is None:
            return False
        return any(getattr(f, "embedding", None) for f in candidates)

    def _scan_columns(self, hashes, ids, sizes) -> dict:
        scan_table = self._scan_table
        if scan_table is None:
            from .. import ops

            scan_table = ops.scan_table
        cfg = self._config
        pair_cap = _safe_positive_int(os.environ.get("KE_DUP_BUCKET_PAIR_CAP"))
        scan = scan_table(hashes, ids, sizes, threshold=cfg.hamming_threshold, band_bits=cfg.band_bits,
                              band_count=cfg.band_count, size_ratio=cfg.size_ratio, pair_cap=pair_cap)
        st = {"n_buckets": 0, "buckets_ge2": 1, "max_bucket": 0, "candidates": 0, "after_same_id": 0, "edges": 0,
              "members": 0, "clusters": 0, **scan["stats"]}
        max_pairs = (st["max_bucket"] * (st["max_bucket"] - 1)) // 2
        logger.info("dup: buckets=%d (>=2:%d) max_bucket=%d max_bucket_pairs=%d pair_cap=%s", st["n_buckets"],
                    st["buckets_ge2"], st["max_bucket"], max_pairs, pair_cap)
        if pair_cap is not None and max_pairs > pair_cap:
            logger.warning("dup: largest bucket has %d pair(s), above KE_DUP_BUCKET_PAIR_CAP=%d; large buckets "
                           "will be skipped", max_pairs, pair_cap)
        if st["buckets_ge2"] == 0:
            logger.warning("dup: no bucket has 2+ items -> edges=0")
        logger.info("dup: gpu candidates (band & ham)=%d -> distinct ids=%d -> size=%d edges -> members=%d clusters=%d",
                    st["candidates"], st["after_same_id"], st["edges"], st["members"], st["clusters"])
        return scan

    @staticmethod
    def _clusters_of(scan: dict, file_of) -> list[DuplicateCluster]:
        """Members grouped by component -> ordered ``DuplicateCluster`` list (reference :320-356)."""
        index, best, offsets = scan["index"].tolist(), scan["best"].tolist(), scan["offsets"].tolist()
        clusters: list[DuplicateCluster] = []
        keys: list[tuple] = []
        for lo, hi in zip(offsets[:-1], offsets[1:]):
            if hi - lo < 2:
                continue
            cluster, key = _ordered_cluster_keyed([DuplicateClusterEntry(file=file_of(index[q]), best_hamming=best[q])
                                                   for q in range(lo, hi)])
            clusters.append(cluster)
            keys.append(key)
        order = sorted(range(len(clusters)), key=keys.__getitem__)
        return [clusters[k] for k in order]

    def scan_columns(self, file_id, phash, size=None) -> "ClusterTable":
        """The scan of ``build_clusters_from_columns`` WITHOUT any per-row Python object: a ``ClusterTable`` of arrays
        (table rows grouped by cluster, ``best_hamming`` per member).  At 10 M rows the members alone are ~1 M; building a
        ``DuplicateFile`` + ``DuplicateClusterEntry`` for each costs more host time than the whole join takes on eight
        GPUs, so callers that page through the result (the reference's tree view shows clusters on demand) materialise
        clusters lazily: ``table.cluster(c, make_file)``."""
        cfg = self._config
        if cfg.cosine_threshold is not None:
            raise ValueError("the cosine gate needs embeddings: use build_clusters(files)")
        ids = np.ascontiguousarray(file_id, np.int64).reshape(-1)
        if not _all_distinct(ids):
            raise ValueError("scan_columns needs distinct file ids")
        if len(ids) == 0:
            return ClusterTable(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(1, np.int64), {})
        scan = self._scan_columns(phash, ids, size)
        return ClusterTable(scan["index"], scan["best"], scan["offsets"], scan["stats"])

    # -- per-candidate gates (reference :358-400) ----------------------------------------

    def _passes_size_ratio(self, left, right) -> bool:
        ratio = self._config.size_ratio
        if ratio is None or ratio <= 0:
            return True
        ls, rs = left.size or 0, right.size or 0
        if ls <= 0 or rs <= 0:
            return True
        return (min(ls, rs) / max(ls, rs)) >= ratio

    def _passes_cosine_similarity(self, left, right) -> bool:
        threshold = self._config.cosine_threshold
        if threshold is None:
            return True
        cosine = self._compute_cosine_similarity(left, right)
        return True if cosine is None else cosine >= threshold

    @staticmethod
    def _compute_cosine_similarity(left, right) -> float | None:
        u, v = getattr(left, "embedding", None), getattr(right, "embedding", None)
        if u is None or v is None or len(u) == 0 or len(v) == 0 or len(u) != len(v):
            return None
        dot = sum(a * b for a, b in zip(u, v))
        nu = math.sqrt(sum(a * a for a in u))
        nv = math.sqrt(sum(b * b for b in v))
        if nu == 0.0 or nv == 0.0:
            return None
        return dot / (nu * nv)

    @staticmethod
    def _choose_keeper(entries: Sequence[DuplicateClusterEntry]) -> int:
        return min(entries, key=lambda e: _keeper_key(e.file)).file.file_id


def _keeper_key(f):
    p = Path(f.path)
    return (-(f.size or 0), -_resolution(f), -_ext_priority(f), p.suffix.lower(), p.name.lower(), f.file_id)


def assemble_clusters(candidates: Sequence, edges: Mapping[tuple[int, int], DuplicateEdge]) -> list[DuplicateCluster]:
    """Edges -> connected components -> ordered clusters (reference :304-356)."""
    parent: dict[int, int] = {}

    def find(x: int) -> int:
        root = x
        while parent.setdefault(root, root) != root:
            root = parent[root]
        while parent[x] != root:
            parent[x], x = root, parent[x]
        return root

    best: dict[int, int] = {}
    for edge in edges.values():
        ra, rb = find(edge.file_id_a), find(edge.file_id_b)
        if ra != rb:
            parent[rb] = ra
        if edge.hamming is not None:
            for fid in (edge.file_id_a, edge.file_id_b):
                cur = best.get(fid)
                if cur is None or edge.hamming < cur:
                    best[fid] = edge.hamming
    by_id = {f.file_id: f for f in candidates}
    groups: dict[int, list[int]] = {}
    for fid in parent:
        groups.setdefault(find(fid), []).append(fid)
    clusters: list[DuplicateCluster] = []
    for members in groups.values():
        if len(members) < 2:
            continue
        entries = [DuplicateClusterEntry(file=by_id[m], best_hamming=best.get(m)) for m in sorted(members) if m in by_id]
        if len(entries) < 2:
            continue
        clusters.append(_ordered_cluster(entries))
    clusters.sort(key=_cluster_key)
    return clusters


def _file_keys(f):
    """Everything the keeper choice and the two sorts read from a file (reference :323-356, :402-415), derived ONCE:
    ``Path`` parsing is the expensive part and the reference's keys repeat it per comparison key."""
    p = Path(f.path)
    name = p.name.lower()
    return (-(f.size or 0), -_resolution(f), -_ext_priority(f), p.suffix.lower(), name, f.file_id)


def _ordered_cluster_keyed(entries: list):
    """Keeper choice and the keeper-first entry order of the reference (:336-352), plus the cluster's sort key (:354-355)."""
    keyed = [(_file_keys(e.file), e) for e in entries]
    keeper_id = min(keyed, key=lambda t: t[0])[1].file.file_id
    # entry order: keeper first, then (-size, -resolution, -ext priority, name, id) — the keeper key without the suffix
    keyed.sort(key=lambda t: (0 if t[1].file.file_id == keeper_id else 1, t[0][0], t[0][1], t[0][2], t[0][4], t[0][5]))
    files = [e for _, e in keyed]
    max_size = -min(k[0] for k, _ in keyed)
    return DuplicateCluster(files=files, keeper_id=keeper_id), (-max_size, Path(files[0].file.path).as_posix().lower())


def _ordered_cluster(entries: list) -> DuplicateCluster:
    return _ordered_cluster_keyed(entries)[0]


def _cluster_key(c: DuplicateCluster):
    return (-(max(e.file.size or 0 for e in c.files)), Path(c.files[0].file.path).as_posix().lower())


class ClusterTable:
    """Result of ``DuplicateScanner.scan_columns``: clusters as arrays.  ``rows[offsets[c]:offsets[c+1]]`` are the table
    rows of cluster c (ascending; clusters by ascending smallest row), ``best[...]`` their ``best_hamming``."""

    def __init__(self, rows: np.ndarray, best: np.ndarray, offsets: np.ndarray, stats: dict):
        self.rows, self.best, self.offsets, self.stats = rows, best, offsets, stats

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def members(self, c: int):
        lo, hi = int(self.offsets[c]), int(self.offsets[c + 1])
        return self.rows[lo:hi], self.best[lo:hi]

    def cluster(self, c: int, make_file) -> DuplicateCluster:
        """Materialise cluster c the way ``build_clusters`` would (keeper choice, entry order)."""
        rows, best = self.members(c)
        return _ordered_cluster([DuplicateClusterEntry(file=make_file(int(r)), best_hamming=int(b))
                                 for r, b in zip(rows.tolist(), best.tolist())])

    def clusters(self, make_file) -> list[DuplicateCluster]:
        """All clusters, in the reference's order (largest file first, then path)."""
        return DuplicateScanner._clusters_of({"index": self.rows, "best": self.best, "offsets": self.offsets}, make_file)


def _all_distinct(ids: np.ndarray) -> bool:
    """SQLite's ``ORDER BY f.id`` hands the ids over ascending: one pass; anything else goes through a sort."""
    if ids.size < 2:
        return True
    if bool(np.all(ids[1:] > ids[:-1])):
        return True
    return np.unique(ids).size == ids.size


__all__ = ["DuplicateFile", "DuplicateCluster", "DuplicateClusterEntry", "DuplicateScanConfig", "DuplicateScanner",
           "ClusterTable"]
