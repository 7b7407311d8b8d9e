"""world_size-2 `gloo` tests (CPU) of the multi-GPU sharding logic in kobato_b200.dist: hash
all-gather, tile-split join with per-rank candidate lists gathered to rank 0, SSIM pair shards.
The CUDA kernels are replaced by the oracle through the functions' injection seams."""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, size: int, init_file: str, out_dir: str):
    for p in (str(ROOT), str(ROOT / "kobato-eyes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import oracle
    from kobato_b200 import dist as kdist
    from kobato_b200 import synth

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=size)
    try:
        assert kdist.world() == (rank, size)
        # ---- hash shards of unequal length -> the same full table everywhere
        full = synth.synth_hashes(3001, seed=5, planted=0.2)
        lo, hi = kdist.shard_range(len(full), rank, size)
        local = torch.from_numpy(full[lo:hi].view(np.int64))
        table = kdist.all_gather_hashes(local)
        assert np.array_equal(table.numpy().view(np.uint64), full)
        # ---- per-rank row blocks of unequal length (the packed candidate exchange of pipeline.scan)
        mine = torch.arange(rank * 100, rank * 100 + 5 + 3 * rank, dtype=torch.int64)
        rows = kdist.all_gather_rows(torch.stack([mine, mine * 7], dim=1))
        want = np.concatenate([np.arange(r * 100, r * 100 + 5 + 3 * r) for r in range(size)])
        assert np.array_equal(rows[:, 0].numpy(), want) and np.array_equal(rows[:, 1].numpy(), want * 7)
        assert kdist.all_gather_rows(torch.empty((0, 2), dtype=torch.int64)).shape == (0, 2) if size == 1 else True
        # ---- broadcast from rank 0
        t0 = torch.from_numpy(full.view(np.int64).copy()) if rank == 0 else torch.empty(0, dtype=torch.int64)
        got = kdist.broadcast_table(t0, src=0)
        assert np.array_equal(got.numpy().view(np.uint64), full)

        # ---- join split by part_index/part_count, gathered on rank 0
        def cpu_join(tbl, threshold, *, require_band, band_bits, band_count, band_allow, part_index, part_count):
            h = tbl.numpy().view(np.uint64)
            i, j, d = oracle.hamming_join(h, threshold, require_band=require_band, band_bits=band_bits,
                                          band_count=band_count)
            mine = (i.astype(np.int64) // 256 + j.astype(np.int64) // 256) % part_count == part_index  # "tiles"
            return i[mine], j[mine], d[mine]

        merged = kdist.distributed_join(table, 8, require_band=True, join=cpu_join)
        if rank == 0:
            wi, wj, wd = oracle.hamming_join(full, 8, require_band=True)
            assert np.array_equal(merged[0], wi) and np.array_equal(merged[1], wj) and np.array_equal(merged[2], wd)
            assert len(wi) > 50
        else:
            assert merged is None

        # ---- SSIM pairs sharded by contiguous index
        bank = synth.synth_images(0, 12, 24, 31, 1, n_set=12, planted=0.5)
        ia = np.arange(11)
        ib = np.arange(11) + 1

        def cpu_ssim(bk, a, b_):
            return oracle.ssim_batch(bk, a, b_, exact=True)

        lo, hi, s = kdist.sharded_ssim(bank, ia, ib, ssim=cpu_ssim)
        assert (lo, hi) == kdist.shard_range(11, rank, size) and len(s) == hi - lo
        np.save(os.path.join(out_dir, f"ssim_{rank}.npy"), np.asarray(s))
        np.save(os.path.join(out_dir, f"range_{rank}.npy"), np.array([lo, hi]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
    from kobato_b200 import dist as kdist

    for n in (0, 1, 7, 70000, 70001):
        for size in (1, 2, 3, 8):
            parts = [kdist.shard_range(n, r, size) for r in range(size)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    import torch.multiprocessing as mp

    import oracle
    from kobato_b200 import synth

    oracle.build()
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "rendezvous")
        mp.spawn(_worker, args=(2, init_file, tmp), nprocs=2, join=True)
        bank = synth.synth_images(0, 12, 24, 31, 1, n_set=12, planted=0.5)
        want = oracle.ssim_batch(bank, np.arange(11), np.arange(11) + 1, exact=True)
        got = np.concatenate([np.load(os.path.join(tmp, f"ssim_{r}.npy")) for r in range(2)])
        assert np.allclose(got, want, atol=0, rtol=0)


# ------------------------------------------------------------------ cross-shard SSIM verification (pipeline.verify_pairs)


def test_cross_pair_plan_is_consistent_across_ranks():
    """Every rank derives the plan from the same candidate list: what r sends to s is what s expects from r, every pair
    has exactly one scorer, cross pairs are spread over both owners, unequal shards are handled by prefix offsets."""
    sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
    from kobato_b200 import dist as kdist

    rng = np.random.default_rng(5)
    for counts in ([5, 5], [7, 3, 9], [1, 0, 4, 6], [100] * 8):
        offsets = np.concatenate([[0], np.cumsum(counts)])
        n = int(offsets[-1])
        ci = rng.integers(0, n - 1, 400)
        cj = np.array([rng.integers(a + 1, n) for a in ci])
        size = len(counts)
        plans = [kdist.plan_cross_pairs(ci, cj, offsets, r, size) for r in range(size)]
        scored = np.zeros(len(ci), int)
        for r, p in enumerate(plans):
            scored[p["local"]] += 1
            scored[p["cross"]] += 1
            for s_ in range(size):
                assert np.array_equal(p["send"][s_], plans[s_]["recv"][r])
                assert all(offsets[r] <= x < offsets[r + 1] for x in p["send"][s_])  # I only send my own rows
            assert len(p["send"][r]) == 0 and len(p["recv"][r]) == 0
            own = np.searchsorted(offsets, ci[p["local"]], side="right") - 1
            assert np.all(own == r)
        assert np.all(scored == 1)
        # the tensor version (what pipeline.scan runs, on the GPU) is the same plan
        import torch

        for r, p in enumerate(plans):
            t = kdist.plan_cross_pairs_t(torch.from_numpy(ci), torch.from_numpy(cj), torch.from_numpy(offsets.astype(np.int64)), r, size)
            assert np.array_equal(t["local"].numpy(), p["local"]) and np.array_equal(t["cross"].numpy(), p["cross"])
            assert t["send_counts"] == [len(x) for x in p["send"]] and t["recv_counts"] == [len(x) for x in p["recv"]]
            assert np.array_equal(t["send_rows"].numpy(), np.concatenate(p["send"]).astype(np.int64))
            assert np.array_equal(t["recv_rows"].numpy(), np.concatenate(p["recv"]).astype(np.int64))
            assert np.array_equal(t["own_i"].numpy(), p["own_i"])
        if size == 8:  # balance: no rank scores more than twice its fair share of the cross pairs
            cross = [len(p["cross"]) for p in plans]
            assert max(cross) <= 2 * sum(cross) / size


def _verify_worker(rank: int, size: int, init_file: str, out_dir: str, counts):
    for p in (str(ROOT), str(ROOT / "kobato-eyes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import oracle
    from kobato_b200 import pipeline, synth

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=size)
    try:
        offsets = np.concatenate([[0], np.cumsum(counts)])
        n, h, w = int(offsets[-1]), 20, 24
        images = synth.synth_images(0, n, h, w, 3, n_set=n, planted=0.4)  # the global set; this rank keeps its shard
        bank = torch.from_numpy(images[offsets[rank]:offsets[rank + 1]].copy())
        rng = np.random.default_rng(9)  # the same candidate list on every rank, pairs all over the table
        ci = rng.integers(0, n - 1, 60)
        cj = np.array([rng.integers(a + 1, n) for a in ci])
        order = np.lexsort((cj, ci))
        ci, cj = ci[order], cj[order]

        def cpu_luma(bk, idx):
            arr = bk.numpy()
            return torch.from_numpy(np.stack([oracle.to_l(arr[int(k)]) for k in np.asarray(idx)])
                                    if len(idx) else np.zeros((0, h, w), np.uint8))

        def cpu_ssim(bk, a, b):
            arr = bk.numpy()
            planes = arr if arr.ndim == 3 else np.stack([oracle.to_l(x) for x in arr])
            return torch.from_numpy(oracle.ssim_batch(planes, np.asarray(a), np.asarray(b)))

        scores, counters = pipeline.verify_pairs(bank, ci, cj, offsets, ssim_batch=cpu_ssim, luma_planes=cpu_luma)
        planes = np.stack([oracle.to_l(x) for x in images])
        want = oracle.ssim_batch(planes, ci, cj)
        assert np.array_equal(scores.numpy(), want), np.abs(scores.numpy() - want).max()
        assert counters["ssim_pairs_local"] + counters["ssim_pairs_cross"] <= len(ci)
        np.save(os.path.join(out_dir, f"cross_{rank}.npy"), np.array([counters["ssim_pairs_cross"], counters["planes_sent"]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("counts", [[9, 9], [11, 4, 7]])
def test_verify_pairs_cross_shard_gloo(counts):
    """pipeline.verify_pairs on world sizes 2 and 3 with UNEQUAL shards: every pair's score equals the oracle's on the
    global set although most pairs straddle two ranks (kernels replaced by the oracle; the exchange is the real one)."""
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_verify_worker, args=(len(counts), os.path.join(tmp, "rendezvous"), tmp, counts), nprocs=len(counts), join=True)
        cross = sum(int(np.load(os.path.join(tmp, f"cross_{r}.npy"))[0]) for r in range(len(counts)))
        assert cross >= 10  # the cross-shard branch really ran
