"""The whole duplicate scan on decoded images: pHash+dHash -> all-pairs Hamming join -> SSIM
verification -> clusters.  This is the call behind bench.py's step and its end-to-end number.

Per rank (one process per GPU):
    K1  hashes its own image shard (no collective)
    --  all_gather of the hash shards (NCCL; the only exchange the path needs)
    K2  joins its share of the triangle's tiles over the full table
    --  candidate lists gathered to rank 0 on the host, pair list broadcast back
    K3  verifies the pairs whose first image it owns (images of the rare cross-shard pairs are
        sent peer to peer)
    --  rank 0 unions the accepted pairs into clusters (host, a few thousand pairs)
"""
from __future__ import annotations

import os
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


def scan(bank, *, host_images=None, threshold: int = 8, ssim_threshold: float = 0.9, require_band: bool = True,
         chunk_images: int = 2048, n_local: int | None = None) -> ScanOutput:
    """Run the duplicate scan over this rank's shard.

    bank: CUDA uint8 [n,h,w,c] holding (or receiving) the rank's decoded images.
    host_images: optional pinned CPU uint8 tensor of the same shape; when given, the images are
        copied host->device in chunks overlapped with K1 (the end-to-end path) and every result is
        read back to the host; when None the bank is taken as already resident.
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
    table = kdist.all_gather_hashes(ph)
    tm.mark("exchange")
    total = table.numel()
    li, lj, ld = ops.hamming_join_device(table, threshold, require_band=require_band, part_index=rank,
                                         part_count=size, capacity=max(1 << 16, 2 * total))
    tm.mark("join")
    # every rank gets every candidate: one packed gather of (i << 32 | j, dist) rows, sorted by (i, j) on the device,
    # one device->host copy
    key = (li.to(torch.int64) & 0xFFFFFFFF) << 32 | (lj.to(torch.int64) & 0xFFFFFFFF)
    rows = kdist.all_gather_rows(torch.stack([key, ld.to(torch.int64)], dim=1))
    rows = rows[torch.argsort(rows[:, 0])].cpu().numpy()
    ci = (rows[:, 0] >> 32) & 0xFFFFFFFF
    cj = rows[:, 0] & 0xFFFFFFFF
    cd = rows[:, 1].astype(np.uint8)
    out.bytes_d2h += 9 * len(ci)
    out.counts.update(images_local=int(n), images_total=int(total), candidates=int(len(ci)))

    # ---- K3 ---------------------------------------------------------------------------------
    n_loc = int(n_local if n_local is not None else n)
    own_i = ci // n_loc
    own_j = cj // n_loc
    mine = np.flatnonzero((own_i == rank) & (own_j == rank))
    scores_dev = torch.zeros(len(ci), dtype=torch.float64, device=dev)
    tm.mark("pre_ssim")
    if len(mine):
        s = ops.ssim_batch(bank, ci[mine] - rank * n_loc, cj[mine] - rank * n_loc)
        scores_dev[torch.from_numpy(mine).to(dev)] = s
    tm.mark("ssim")
    cross = np.flatnonzero(own_i != own_j)
    if size > 1 and len(cross):
        # the rare cross-shard pairs: rank a (owner of i) scores the pair, rank b ships image j; all transfers are posted
        # before the first wait
        dist = kdist._dist()
        tmp = torch.empty((2 * len(cross), h, w, c), dtype=torch.uint8, device=dev)
        p2p, sel = [], []
        for k, q in enumerate(cross.tolist()):
            a, b = int(own_i[q]), int(own_j[q])
            if rank == a:
                tmp[2 * k].copy_(bank[int(ci[q]) - a * n_loc])
                p2p.append(dist.P2POp(dist.irecv, tmp[2 * k + 1], b))
                sel.append(k)
            elif rank == b:
                p2p.append(dist.P2POp(dist.isend, bank[int(cj[q]) - b * n_loc].contiguous(), a))
        # plain isend/irecv: batch_isend_irecv measured ~10 ms slower per step here (2 x B200, NCCL 2.28: the grouped
        # point-to-point launch stalls the step's next collective)
        if p2p and os.environ.get("KE_P2P_BATCHED"):  # tuning probe
            for req in dist.batch_isend_irecv(p2p):
                req.wait()
        else:
            for req in [op.op(op.tensor, op.peer) for op in p2p]:
                req.wait()
        if sel:
            s = ops.ssim_batch(tmp, [2 * k for k in sel], [2 * k + 1 for k in sel])
            scores_dev[torch.from_numpy(cross[sel]).to(dev)] = s
    if size > 1:
        kdist._dist().all_reduce(scores_dev)  # every pair was scored by exactly one rank: SUM merges
    scores = scores_dev.cpu().numpy()
    out.bytes_d2h += 8 * len(ci)
    tm.mark("post")

    # ---- host assembly (rank 0) ---------------------------------------------------------------
    if rank == 0:
        t0 = time.perf_counter()
        keep = scores >= ssim_threshold
        out.cand_i, out.cand_j, out.cand_d, out.ssim, out.accepted = ci, cj, cd, scores, keep
        out.clusters = _components(ci.astype(np.int64), cj.astype(np.int64), keep)
        out.counts.update(accepted=int(np.count_nonzero(keep)), clusters=len(out.clusters))
        out.stage_ms["host_assembly"] = (time.perf_counter() - t0) * 1e3
    if host_images is not None:
        # the hashes go back to the host too (they are what the reference stores in SQLite)
        out.phash_host = ph.cpu().numpy()
        out.dhash_host = dh.cpu().numpy()
        out.bytes_d2h += 16 * n
    torch.cuda.synchronize(dev)
    out.stage_ms.update(tm.result())
    out.counts["ssim_pairs_local"] = int(len(mine))
    return out
