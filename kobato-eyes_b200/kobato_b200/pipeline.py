"""The whole duplicate scan on decoded images: pHash+dHash -> all-pairs Hamming join -> SSIM
verification -> clusters.  This is the call behind bench.py's step and its end-to-end number.

Per rank (one process per GPU):
    K1  hashes its own image shard (no collective)
    --  all_gather of the hash shards (NCCL; the only exchange the path needs)
    K2  joins its share of the triangle's tiles over the full table
    --  candidate lists gathered on every rank (one packed all_gather); the list stays on the device
    K3  verifies its pairs: same-shard pairs from its own bank; a cross-shard pair is scored by one of the two owners
        (by the parity of i + j), the other image arriving as a luma plane in ONE packed all_to_all per step
    --  rank 0 unions the accepted pairs into clusters (device union-find) and reads candidates, scores and clusters back
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import dist as kdist
from . import ops


@dataclass
class ScanOutput:
    phash: object = None          # int64 CUDA tensor, local shard
    dhash: object = None
    cand_i: np.ndarray | None = None   # rank 0: merged candidates (global indices), sorted
    cand_j: np.ndarray | None = None
    cand_d: np.ndarray | None = None
    ssim: np.ndarray | None = None     # rank 0: score per candidate
    accepted: np.ndarray | None = None  # rank 0: bool per candidate
    clusters: list | None = None       # rank 0: [(representative, [members])]
    stage_ms: dict = field(default_factory=dict)   # device time of this rank's kernels
    counts: dict = field(default_factory=dict)
    bytes_h2d: int = 0
    bytes_d2h: int = 0


def _torch():
    import torch

    return torch


class ClusterSet:
    """Components of the accepted pairs in CSR form (arrays, no per-cluster Python objects until asked for):
    ``len()``, indexing and iteration yield ``(representative, [members])`` sorted by representative — the reference's
    ClusterBuilder semantics (src/dup/cluster.py:22-70)."""

    def __init__(self, members: np.ndarray, offsets: np.ndarray):
        self.members, self.offsets = members, offsets

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def __getitem__(self, c: int):
        if c < 0:
            c += len(self)
        if not 0 <= c < len(self):
            raise IndexError(c)
        m = self.members[self.offsets[c]:self.offsets[c + 1]]
        return int(m[0]), m.tolist()

    def __iter__(self):
        return (self[c] for c in range(len(self)))

    def as_list(self):
        return list(self)


def _components(i, j, keep) -> ClusterSet:
    return ClusterSet(*ops.cluster_pairs_csr(i[keep], j[keep]))


class Timer:
    """CUDA-event stage timer on the current stream."""

    def __init__(self):
        self.marks = []

    def mark(self, name):
        torch = _torch()
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def result(self):
        out = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


def verify_pairs(bank, ci, cj, offsets, *, ssim_batch=None, luma_planes=None, on_local_done=None):
    """SSIM of the candidate pairs (global table rows ``ci[k] < cj[k]``, the same list on every rank) over image shards:
    rank r holds the images of rows ``offsets[r] .. offsets[r+1]`` in ``bank``.  Returns (float64 scores for ALL pairs on
    every rank, counters).

    Same-shard pairs are scored from the rank's own bank.  A cross-shard pair is scored by one of its two owners
    (``dist.plan_cross_pairs``); the other image travels as a luma plane — a third of the RGB bytes, and SSIM of the
    planes equals SSIM of the RGB images because the luma is the same fixed-point conversion — in ONE packed
    ``all_to_all`` per call.  Every rank derives the whole plan from the global candidate list, so no metadata is
    exchanged; the plan is computed where the list lives (``dist.plan_cross_pairs_t``: on the GPU inside ``scan``), the
    host only learns the per-peer plane counts.  ``ci``/``cj``: numpy arrays or tensors.  ``ssim_batch`` /
    ``luma_planes`` default to the CUDA kernels (the gloo tests inject CPU stand-ins)."""
    torch = _torch()
    native = ssim_batch is None and luma_planes is None
    ssim_batch = ssim_batch or ops.ssim_batch
    luma_planes = luma_planes or ops.luma_planes
    kw = {"check": False} if native else {}  # indices derived here from a range-checked candidate list
    rank, size = kdist.world()
    dev = bank.device
    h, w = int(bank.shape[1]), int(bank.shape[2])
    ci = torch.as_tensor(ci, dtype=torch.int64).to(dev)
    cj = torch.as_tensor(cj, dtype=torch.int64).to(dev)
    offs = torch.as_tensor(np.asarray(offsets, np.int64)).to(dev)
    if native and ci.numel() and (int(torch.min(ci)) < 0 or int(torch.max(cj)) >= int(offsets[-1])):
        raise ValueError("candidate rows outside the table")
    if size == 1:  # one shard: every pair is local, nothing to plan
        scores = ssim_batch(bank, ci, cj, **kw) if ci.numel() else torch.zeros(0, dtype=torch.float64, device=dev)
        if on_local_done:
            on_local_done()
        return scores, {"ssim_pairs_local": int(ci.numel()), "ssim_pairs_cross": 0, "planes_sent": 0, "planes_received": 0,
                        "plane_bytes_sent": 0}
    plan = kdist.plan_cross_pairs_t(ci, cj, offs, rank, size)
    my_lo = int(offsets[rank])
    mine, cross = plan["local"], plan["cross"]
    scores = torch.zeros(ci.numel(), dtype=torch.float64, device=dev)
    n_sent, n_recv = int(sum(plan["send_counts"])), int(sum(plan["recv_counts"]))
    pending = tmp = None
    if size > 1:
        # One buffer of 'L' planes for the cross-shard pairs: [planes received | planes of my own images].  The travelling
        # planes leave FIRST (asynchronous all_to_all straight into the buffer's head); my own ends are converted and the
        # same-shard pairs scored while they fly.
        n_own = 0
        if cross.numel():
            i_mine = plan["own_i"][cross] == rank
            own_rows = torch.where(i_mine, ci[cross], cj[cross])
            far_rows = torch.where(i_mine, cj[cross], ci[cross])
            own_uni, own_pos = torch.unique(own_rows, return_inverse=True)
            far_pos = torch.searchsorted(plan["recv_rows"], far_rows)
            n_own = int(own_uni.numel())
        tmp = torch.empty((n_recv + n_own, h * w), dtype=torch.uint8, device=dev)
        planes = luma_planes(bank, plan["send_rows"] - my_lo, **kw).reshape(n_sent, h * w)
        _, pending = kdist.exchange_rows(planes, plan["send_counts"], plan["recv_counts"], async_op=True, out=tmp[:n_recv])
        if n_own:
            if native:
                luma_planes(bank, own_uni - my_lo, out=tmp[n_recv:], **kw)
            else:  # injected stand-in (tests) without an `out` parameter
                tmp[n_recv:] = luma_planes(bank, own_uni - my_lo).reshape(n_own, h * w)
    if mine.numel():
        scores[mine] = ssim_batch(bank, ci[mine] - my_lo, cj[mine] - my_lo, **kw)
    if on_local_done:
        on_local_done()
    if size > 1:
        if pending is not None:
            pending.wait()
        if cross.numel():
            a_idx = torch.where(i_mine, n_recv + own_pos, far_pos)
            b_idx = torch.where(i_mine, far_pos, n_recv + own_pos)
            scores[cross] = ssim_batch(tmp.view(-1, h, w), a_idx, b_idx, **kw)
        kdist._dist().all_reduce(scores)  # every pair was scored by exactly one rank: SUM merges
    return scores, {"ssim_pairs_local": int(mine.numel()), "ssim_pairs_cross": int(cross.numel()), "planes_sent": n_sent,
                    "planes_received": n_recv, "plane_bytes_sent": n_sent * h * w}


def _clusters_device(ci, cj, keep, n_nodes: int) -> ClusterSet:
    """Components of the accepted pairs with the device union-find (``ke_cluster_pairs``): members ascending inside a
    component, components by ascending representative (= smallest member) — ``ops.cluster_pairs_csr``'s result without the
    host round trip of the pair list."""
    torch = _torch()
    label = ops.cluster_pairs_device(ci[keep], cj[keep], n_nodes)
    nodes = torch.nonzero(label >= 0).flatten()                       # ascending
    lab, order = torch.sort(label[nodes], stable=True)                 # by representative, members stay ascending
    _, counts = torch.unique_consecutive(lab, return_counts=True)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=lab.device), torch.cumsum(counts, 0)])
    return ClusterSet(nodes[order].cpu().numpy(), offsets.cpu().numpy())


def scan(bank, *, host_images=None, threshold: int = 8, ssim_threshold: float = 0.9, require_band: bool = True,
         chunk_images: int = 2048, cluster_on_device: bool | None = None) -> ScanOutput:
    """Run the duplicate scan over this rank's shard.

    bank: CUDA uint8 [n,h,w,c] holding (or receiving) the rank's decoded images.
    host_images: optional pinned CPU uint8 tensor of the same shape; when given, the images are
        copied host->device in chunks overlapped with K1 (the end-to-end path) and every result is
        read back to the host; when None the bank is taken as already resident.
    cluster_on_device: where rank 0 unions the accepted pairs (None: by the length of the candidate list).
    """
    torch = _torch()
    rank, size = kdist.world()
    out = ScanOutput()
    n, h, w, c = bank.shape
    dev = bank.device
    tm = Timer()
    ph = torch.empty(n, dtype=torch.int64, device=dev)
    dh = torch.empty(n, dtype=torch.int64, device=dev)

    # ---- K1 ---------------------------------------------------------------------------------
    tm.mark("start")
    if host_images is None:
        p, d = ops.phash_dhash_batch(bank)
        ph.copy_(p)
        dh.copy_(d)
    else:
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)
        for lo in range(0, n, chunk_images):
            hi = min(n, lo + chunk_images)
            with torch.cuda.stream(copy_stream):
                bank[lo:hi].copy_(host_images[lo:hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            main.wait_event(ready)
            p, d = ops.phash_dhash_batch(bank[lo:hi])
            ph[lo:hi].copy_(p)
            dh[lo:hi].copy_(d)
        out.bytes_h2d += n * h * w * c
    tm.mark("phash")
    out.phash, out.dhash = ph, dh

    # ---- exchange + K2 ----------------------------------------------------------------------
    table, shard_counts = kdist.all_gather_hashes(ph, return_counts=True)
    offsets = np.concatenate([[0], np.cumsum(shard_counts)]).astype(np.int64)
    tm.mark("exchange")
    total = table.numel()
    li, lj, ld = ops.hamming_join_device(table, threshold, require_band=require_band, part_index=rank,
                                         part_count=size, capacity=max(1 << 16, 2 * total))
    tm.mark("join")
    # every rank gets every candidate: one packed gather of (i << 32 | j, dist) rows, sorted by (i, j) — and the list STAYS
    # on the device: the cross-shard plan, the scores and the clusters are computed there, the host reads the results
    # once at the end
    key = (li.to(torch.int64) & 0xFFFFFFFF) << 32 | (lj.to(torch.int64) & 0xFFFFFFFF)
    rows = kdist.all_gather_rows(torch.stack([key, ld.to(torch.int64)], dim=1))
    rows = rows[torch.argsort(rows[:, 0])]
    ci_t = (rows[:, 0] >> 32) & 0xFFFFFFFF
    cj_t = rows[:, 0] & 0xFFFFFFFF
    out.counts.update(images_local=int(n), images_total=int(total), candidates=int(rows.shape[0]))

    # ---- K3 ---------------------------------------------------------------------------------
    tm.mark("pre_ssim")
    scores_dev, k3 = verify_pairs(bank, ci_t, cj_t, offsets, on_local_done=lambda: tm.mark("ssim"))
    out.counts.update(k3)

    # ---- clusters + results to the host (rank 0) -----------------------------------------------
    if rank == 0:
        # long lists are unioned on the device (a handful of launches, ~0.5 ms whatever the length); short ones by the
        # library's host union-find after the copy (0.2 ms at 3 600 pairs, 1.5 ms at 29 000)
        on_device = rows.shape[0] > 12000 if cluster_on_device is None else bool(cluster_on_device)
        if on_device:
            out.clusters = _clusters_device(ci_t, cj_t, scores_dev >= ssim_threshold, int(total))
        host = torch.cat([rows, scores_dev.view(torch.int64).unsqueeze(1)], dim=1).cpu().numpy()  # one copy: key, dist, score bits
        tm.mark("post")
        t0 = time.perf_counter()
        out.cand_i = (host[:, 0] >> 32) & 0xFFFFFFFF
        out.cand_j = host[:, 0] & 0xFFFFFFFF
        out.cand_d = host[:, 1].astype(np.uint8)
        out.ssim = host[:, 2].copy().view(np.float64)
        out.accepted = out.ssim >= ssim_threshold
        if not on_device:
            out.clusters = _components(out.cand_i, out.cand_j, out.accepted)
        out.bytes_d2h += 17 * len(out.cand_i) + (8 * (len(out.clusters.members) + len(out.clusters.offsets)) if on_device else 0)
        out.counts.update(accepted=int(np.count_nonzero(out.accepted)), clusters=len(out.clusters))
        out.stage_ms["host_assembly"] = (time.perf_counter() - t0) * 1e3
    else:
        tm.mark("post")
    if host_images is not None:
        # the hashes go back to the host too (they are what the reference stores in SQLite)
        out.phash_host = ph.cpu().numpy()
        out.dhash_host = dh.cpu().numpy()
        out.bytes_d2h += 16 * n
    torch.cuda.synchronize(dev)
    out.stage_ms.update(tm.result())
    return out
