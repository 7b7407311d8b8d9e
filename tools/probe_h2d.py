"""Host-feed roofline: bare pinned host->device copies on N ranks at once (one process per GPU, like bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_h2d.py [--gb 8]

Prints one JSON line: per-GPU and aggregate GB/s for chunk sizes from 64 MB to the whole buffer, plus the rank's CPU
affinity and NUMA node of its GPU (nvidia-smi topo), so that the end-to-end number of bench.py (`e2e.h2d_frac`) can be
read against what the box's PCIe / host memory system delivers when N GPUs pull at the same time."""
import argparse
import json
import os
import subprocess

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=8.0)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = int(args.gb * (1 << 30))
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(rank + 1)
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
out = {}
for chunk_mb in (64, 512, 1536, nbytes >> 20):
    chunk = chunk_mb << 20

    def run():
        for lo in range(0, nbytes, chunk):
            dst[lo:lo + chunk].copy_(host[lo:lo + chunk], non_blocking=True)

    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run()
    b.record()
    b.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[f"{chunk_mb}MB"] = round(nbytes / (float(ms.item()) * 1e-3) / 1e9, 2)
aff = sorted(os.sched_getaffinity(0))
info = {"rank": rank, "cpus": f"{aff[0]}-{aff[-1]} ({len(aff)})"}
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, info)
else:
    gathered = [info]
if rank == 0:
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    print(json.dumps({"world": world, "gb_per_rank": args.gb, "h2d_gbs_per_gpu_by_chunk": out,
                      "h2d_gbs_aggregate": {k: round(v * world, 1) for k, v in out.items()}, "ranks": gathered,
                      "topo": [ln for ln in topo.splitlines() if ln.startswith("GPU")][:8]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
