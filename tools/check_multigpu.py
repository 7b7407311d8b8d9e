"""Correctness of the multi-process duplicate scan on real GPUs, against the CPU oracle on the WHOLE set.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu.py \
        [--images 260] [--uneven] [--h 96] [--w 96] [--planted 0.25] [--json out.json]

Every rank holds a contiguous shard of ONE global synthetic set whose planted near-duplicates point at uniformly chosen
earlier images, so most candidate pairs straddle two ranks (with --uneven the shards also differ in length).  Each rank
runs kobato_b200.pipeline.scan (device-resident, then from pinned host memory); rank 0 regenerates the set on the CPU
and checks every pHash / dHash, the candidate list, every SSIM score (1e-5), the accept / reject decisions and the
clusters against the oracle.  Exit code 0 and a line "multigpu check ok" only when everything matches.
tests/test_gpu_round2.py runs it at world size 2; the 8-GPU output is kept under profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "kobato-eyes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=260, help="images per rank")
    ap.add_argument("--uneven", action="store_true", help="rank r holds images + 17 r images")
    ap.add_argument("--h", type=int, default=96)
    ap.add_argument("--w", type=int, default=96)
    ap.add_argument("--planted", type=float, default=0.25)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from kobato_b200 import ops, pipeline, synth

    counts = [args.images + (17 * r if args.uneven else 0) for r in range(world)]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    total = int(offsets[-1])
    lo, n = int(offsets[rank]), counts[rank]
    bank = ops.synth_images_device(lo, n, args.h, args.w, 3, n_set=total, planted=args.planted)
    t0 = time.perf_counter()
    a = pipeline.scan(bank, threshold=8, ssim_threshold=0.9)
    t_scan = time.perf_counter() - t0
    pinned = bank.cpu().pin_memory()
    b = pipeline.scan(torch.empty_like(bank), host_images=pinned, threshold=8, ssim_threshold=0.9, chunk_images=64)

    # every rank's hashes to rank 0 (the scan keeps them sharded)
    ph_all = [torch.empty(c, dtype=torch.int64, device=dev) for c in counts] if rank == 0 else None
    dh_all = [torch.empty(c, dtype=torch.int64, device=dev) for c in counts] if rank == 0 else None
    if rank == 0:
        ph_all[0].copy_(a.phash)
        dh_all[0].copy_(a.dhash)
        for r in range(1, world):
            dist.recv(ph_all[r], r)
            dist.recv(dh_all[r], r)
    else:
        dist.send(a.phash.contiguous(), 0)
        dist.send(a.dhash.contiguous(), 0)
    counters = torch.tensor([a.counts["ssim_pairs_local"], a.counts["ssim_pairs_cross"], a.counts["planes_sent"]],
                            dtype=torch.int64, device=dev)
    gathered = [torch.zeros_like(counters) for _ in range(world)]
    dist.all_gather(gathered, counters)
    ok = True
    report = {}
    if rank == 0:
        import oracle
        from oracle import ref_py

        oracle.build()
        host = synth.synth_images(0, total, args.h, args.w, 3, n_set=total, planted=args.planted)
        ph = torch.cat(ph_all).cpu().numpy().view(np.uint64)
        dh = torch.cat(dh_all).cpu().numpy().view(np.uint64)
        want_ph, want_dh = oracle.signature_batch(host, threads=os.cpu_count() or 1)[:2]
        bad_hash = int(np.count_nonzero(ph != want_ph) + np.count_nonzero(dh != want_dh))
        wi, wj, wd = oracle.hamming_join(ph, 8, require_band=True, threads=os.cpu_count() or 1)
        cand_ok = all(np.array_equal(x.cand_i, wi) and np.array_equal(x.cand_j, wj) and np.array_equal(x.cand_d, wd)
                      for x in (a, b))
        planes = np.stack([oracle.to_l(x) for x in host])
        want = np.array([ref_py.ssim_of_planes(planes[i], planes[j]) for i, j in zip(wi.tolist(), wj.tolist())])
        err = float(np.abs(a.ssim - want).max()) if cand_ok and len(want) else (0.0 if cand_ok else float("nan"))
        same_e2e = bool(cand_ok and np.array_equal(a.ssim, b.ssim))
        safe = np.abs(want - 0.9) > 1e-5
        decisions_ok = bool(cand_ok and np.array_equal(a.accepted[safe], (want >= 0.9)[safe]))
        clusters_ok = bool(cand_ok and a.clusters.as_list() ==
                           ref_py.cluster_matches(zip(wi.tolist(), wj.tolist(), a.accepted.tolist())))
        own = np.searchsorted(offsets, wi, side="right") - 1
        own_j = np.searchsorted(offsets, wj, side="right") - 1
        cross = int(np.count_nonzero(own != own_j))
        per_rank = [[int(v) for v in g.tolist()] for g in gathered]
        report = {"world": world, "shards": counts, "images": total, "hash_mismatches": bad_hash,
                  "candidates": int(len(wi)), "candidates_cross_shard": cross, "candidates_equal_oracle": bool(cand_ok),
                  "ssim_max_abs_err": err, "ssim_host_path_identical": same_e2e, "decisions_equal_oracle": decisions_ok,
                  "clusters_equal_oracle": clusters_ok, "clusters": len(a.clusters), "accepted": int(a.accepted.sum()),
                  "per_rank_[local_pairs,cross_pairs,planes_sent]": per_rank, "scan_wall_s": round(t_scan, 3)}
        ok = (bad_hash == 0 and cand_ok and err <= 1e-5 and same_e2e and decisions_ok and clusters_ok and len(wi) > 0
              and (world == 1 or cross > len(wi) // 4)
              and sum(p[0] + p[1] for p in per_rank) == len(wi))  # every pair scored exactly once
        report["ok"] = bool(ok)
        print(json.dumps(report))
        if args.json:
            Path(args.json).write_text(json.dumps(report, indent=1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and flag.item():
        print("multigpu check ok")
    return 0 if flag.item() else 1


if __name__ == "__main__":
    raise SystemExit(main())
