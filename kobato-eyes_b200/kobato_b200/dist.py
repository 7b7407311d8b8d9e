"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` for plumbing.

Partitioning:
  * pHash/dHash — images are independent: each rank hashes its own shard, no collective;
  * Hamming join — every rank needs the whole (tiny) hash table: ONE ``all_gather`` of the
    per-rank hash shards (or a ``broadcast`` from rank 0) over NCCL/NVLink, then the N x N triangle's
    tiles are dealt round-robin to ranks (``part_index=rank, part_count=world``) with no further
    exchange; per-rank candidate lists are gathered to rank 0 on the host;
  * SSIM — pairs are independent: sharded by contiguous pair index, no collective.

Everything here also runs on the ``gloo`` backend with CPU tensors (the join itself is injected),
which is how the world_size-2 CPU tests exercise the sharding logic.
"""
from __future__ import annotations

from typing import Callable

import os

import numpy as np

_LEGACY = bool(os.environ.get("KE_DIST_LIST_COLLECTIVES"))  # tuning probe: list-based all_gather instead of the flat one


def _dist():
    import torch.distributed as dist

    return dist


def world() -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, size: int) -> tuple[int, int]:
    """Contiguous [lo, hi) share of n items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n), size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _gather_counts(n_local: int, device) -> list[int]:
    """Row counts of every rank: one collective into one tensor, one device->host read."""
    import torch

    dist = _dist()
    _, size = world()
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    out = torch.empty(size, dtype=torch.int64, device=device)
    try:
        if _LEGACY:
            raise NotImplementedError
        dist.all_gather_into_tensor(out, mine)
    except (RuntimeError, NotImplementedError, AttributeError):  # backends without the flat variant
        parts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(size)]
        dist.all_gather(parts, mine)
        out = torch.cat(parts)
    return out.tolist()


def _gather_padded(padded):
    """[cap, ...] from every rank -> [size, cap, ...] with one collective."""
    import torch

    dist = _dist()
    _, size = world()
    out = torch.empty((size,) + tuple(padded.shape), dtype=padded.dtype, device=padded.device)
    try:
        if _LEGACY:
            raise NotImplementedError
        dist.all_gather_into_tensor(out.view(-1), padded.reshape(-1))
    except (RuntimeError, NotImplementedError, AttributeError):
        parts = [torch.empty_like(padded) for _ in range(size)]
        dist.all_gather(parts, padded)
        out = torch.stack(parts)
    return out


def all_gather_hashes(local, return_counts: bool = False):
    """Concatenate per-rank int64 hash shards (possibly of different lengths) on every rank: the counts, then ONE
    all_gather of the (padded) shards; equal shards (the weak-scaling case) need no padding and no re-packing.
    With ``return_counts`` also the per-rank shard lengths (rank r owns rows ``sum(counts[:r]) ...`` of the table)."""
    import torch

    rank, size = world()
    if size == 1:
        return (local, [int(local.numel())]) if return_counts else local
    local = local.contiguous().view(-1)
    counts = _gather_counts(local.numel(), local.device)
    cap = max(max(counts), 1)
    if local.numel() == cap:
        padded = local
    else:
        padded = torch.zeros(cap, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    out = _gather_padded(padded)
    if all(c == cap for c in counts):
        table = out.view(-1)
    else:
        table = torch.cat([out[r, :c] for r, c in enumerate(counts)])
    return (table, counts) if return_counts else table


def sorted_unique(x: np.ndarray) -> np.ndarray:
    """Sorted distinct values by sort + neighbour compare (np.unique costs 5 ms on 25 k int64 here, a sort 0.3 ms)."""
    if x.size == 0:
        return x
    y = np.sort(x)
    return y[np.concatenate([[True], y[1:] != y[:-1]])]


def plan_cross_pairs(ci: np.ndarray, cj: np.ndarray, offsets: np.ndarray, rank: int, size: int) -> dict:
    """Who verifies which candidate pair, and which images travel for it.  Pure arithmetic on the (global) candidate
    list every rank holds, so all ranks derive the same plan without exchanging anything.

    ``offsets[r]`` = first table row of rank r (``offsets[size]`` = table length; shards may be unequal).  A pair whose
    images live on one rank is scored there; a cross-shard pair (i on rank a, j on rank b) is scored by a when i + j is
    even and by b otherwise, so neither end of the table collects all of them.  The other image comes over as a luma plane.

    Returns: ``local`` (positions of the pairs this rank scores from its own bank), ``cross`` (positions it scores with
    one received image), ``send[s]`` (sorted distinct global rows of MINE that rank s needs), ``recv[r]`` (sorted distinct
    global rows of rank r that I need): ``send`` of r towards s equals ``recv`` of s from r by construction."""
    ci = np.asarray(ci, np.int64)
    cj = np.asarray(cj, np.int64)
    offsets = np.asarray(offsets, np.int64)
    own_i = np.searchsorted(offsets, ci, side="right") - 1
    own_j = np.searchsorted(offsets, cj, side="right") - 1
    is_cross = own_i != own_j
    scorer = np.where(is_cross & (((ci + cj) & 1) == 1), own_j, own_i)
    cross_all = np.flatnonzero(is_cross)
    sc = scorer[cross_all]
    # the image that has to travel: the one NOT owned by the scorer
    i_scores = sc == own_i[cross_all]
    trav = np.where(i_scores, cj[cross_all], ci[cross_all])
    trav_owner = np.where(i_scores, own_j[cross_all], own_i[cross_all])
    # what I send, grouped by destination (one sort of a combined key instead of a mask per peer), and what I receive,
    # grouped by source — rows of one owner are contiguous in the table, so the sorted distinct rows ARE grouped by source
    mine_out = trav_owner == rank
    key = sorted_unique(sc[mine_out] * (int(offsets[-1]) + 1) + trav[mine_out])
    dest, rows_out = np.divmod(key, int(offsets[-1]) + 1)
    cuts = np.searchsorted(dest, np.arange(size + 1))
    send = [rows_out[cuts[p]:cuts[p + 1]] for p in range(size)]
    rows_in = sorted_unique(trav[sc == rank])
    cuts = np.searchsorted(rows_in, offsets)
    recv = [rows_in[cuts[p]:cuts[p + 1]] for p in range(size)]
    return {"local": np.flatnonzero(~is_cross & (own_i == rank)), "cross": cross_all[sc == rank],
            "send": send, "recv": recv, "own_i": own_i, "own_j": own_j, "scorer": scorer}


def plan_cross_pairs_t(ci, cj, offsets, rank: int, size: int) -> dict:
    """``plan_cross_pairs`` on torch tensors of any device (the candidate list never leaves the GPU in ``pipeline.scan``;
    the numpy version costs 2.5-3.5 ms of host time on the 8-GPU step's 28 778 candidates).  ``ci``/``cj``: int64 tensors,
    ``offsets``: int64 tensor ``[size + 1]`` on the same device.  Returns tensors on that device — ``local`` / ``cross``
    (positions this rank scores), ``send_rows`` (my rows that travel, grouped by destination, sorted inside a group),
    ``recv_rows`` (the rows I receive: sorted, hence grouped by source), ``own_i`` — and the two count lists the
    all_to_all needs on the host (``send_counts``, ``recv_counts``: ONE device->host copy)."""
    import torch

    total = int(offsets[-1])
    own_i = torch.searchsorted(offsets, ci, right=True) - 1
    own_j = torch.searchsorted(offsets, cj, right=True) - 1
    is_cross = own_i != own_j
    scorer = torch.where(is_cross & (((ci + cj) & 1) == 1), own_j, own_i)
    i_scores = scorer == own_i
    trav = torch.where(i_scores, cj, ci)              # the image NOT owned by the scorer (meaningless for local pairs)
    trav_owner = torch.where(i_scores, own_j, own_i)
    mine_cross = is_cross & (scorer == rank)
    out_mask = is_cross & (trav_owner == rank)
    key = torch.unique(scorer[out_mask] * (total + 1) + trav[out_mask])   # sorted distinct (destination, row)
    dest = torch.div(key, total + 1, rounding_mode="floor")
    send_rows = key - dest * (total + 1)
    recv_rows = torch.unique(trav[mine_cross])
    counts = torch.stack([torch.bincount(dest, minlength=size)[:size],
                          torch.bincount(torch.searchsorted(offsets, recv_rows, right=True) - 1, minlength=size)[:size]]).cpu()
    return {"local": torch.nonzero(~is_cross & (own_i == rank)).flatten(), "cross": torch.nonzero(mine_cross).flatten(),
            "send_rows": send_rows, "recv_rows": recv_rows, "own_i": own_i,
            "send_counts": counts[0].tolist(), "recv_counts": counts[1].tolist()}


def exchange_rows(rows, send_counts, recv_counts, async_op: bool = False, out=None):
    """ONE ``all_to_all_single``: ``rows`` = [sum(send_counts), k] with the rows for rank 0 first, then rank 1, ...;
    returns [sum(recv_counts), k] ordered by source rank (received straight into ``out`` when given) — or
    ``(buffer, work)`` with ``async_op`` (``work.wait()`` before the buffer is read; ``work`` is None when nothing was
    launched)."""
    import torch

    dist = _dist()
    rank, size = world()
    if out is None:
        out = torch.empty((int(sum(recv_counts)),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    work = None
    if size > 1:
        work = dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=[int(c) for c in recv_counts],
                                      input_split_sizes=[int(c) for c in send_counts], async_op=async_op)
    return (out, work) if async_op else out


def all_gather_varlen(local):
    """Concatenate 1-D tensors of any dtype and per-rank length on every rank (padded all_gather)."""
    import torch

    rank, size = world()
    if size == 1:
        return local
    counts = _gather_counts(local.numel(), local.device)
    cap = max(max(counts), 1)
    padded = torch.zeros(cap, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = _gather_padded(padded)
    return torch.cat([out[r, :c] for r, c in enumerate(counts)])


def all_gather_rows(local):
    """Concatenate ``[m_r, k]`` tensors (per-rank m_r) along dim 0 on every rank with TWO collectives: the row counts,
    then one padded all_gather of the rows."""
    import torch

    rank, size = world()
    if size == 1:
        return local
    local = local.contiguous()
    k = local.shape[1]
    counts = _gather_counts(local.shape[0], local.device)
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = _gather_padded(padded)
    return torch.cat([out[r, :c] for r, c in enumerate(counts)])


def broadcast_table(table, src: int = 0):
    """Rank `src` holds the table (int64 tensor); every rank returns its own copy."""
    import torch

    dist = _dist()
    rank, size = world()
    if size == 1:
        return table
    dev = table.device if table is not None else torch.device("cuda", torch.cuda.current_device())
    n = torch.tensor([table.numel() if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = table if rank == src else torch.empty(int(n.item()), dtype=torch.int64, device=dev)
    dist.broadcast(buf, src)
    return buf


def gather_candidates(i: np.ndarray, j: np.ndarray, d: np.ndarray, dst: int = 0):
    """Per-rank candidate lists -> one (i, j, d) sorted by (i, j) on rank `dst` (None elsewhere)."""
    dist = _dist()
    rank, size = world()
    if size == 1:
        return i, j, d
    payload = (np.ascontiguousarray(i), np.ascontiguousarray(j), np.ascontiguousarray(d))
    gathered = [None] * size if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    ii = np.concatenate([g[0] for g in gathered])
    jj = np.concatenate([g[1] for g in gathered])
    dd = np.concatenate([g[2] for g in gathered])
    order = np.lexsort((jj, ii))
    return ii[order], jj[order], dd[order]


def distributed_join(table, threshold: int, *, require_band: bool = False, band_bits: int = 16, band_count: int = 4,
                     band_allow=None, join: Callable | None = None, dst: int = 0):
    """Tile-split Hamming join of a table every rank already holds.  Returns the merged candidate
    list on rank `dst` and None on the others.  `join` defaults to the CUDA kernel."""
    rank, size = world()
    if join is None:
        from . import ops

        join = ops.hamming_join
    i, j, d = join(table, threshold, require_band=require_band, band_bits=band_bits, band_count=band_count,
                   band_allow=band_allow, part_index=rank, part_count=size)
    return gather_candidates(i, j, d, dst)


def sharded_ssim(bank, ia, ib, *, ssim: Callable | None = None):
    """Each rank scores its contiguous share of the pair list; returns (lo, hi, scores[lo:hi])."""
    rank, size = world()
    lo, hi = shard_range(len(ia), rank, size)
    if ssim is None:
        from . import ops

        ssim = ops.ssim_batch
    return lo, hi, ssim(bank, ia[lo:hi], ib[lo:hi])
