"""torchrun --nproc-per-node 2 tools/probe_dist.py — time the small collectives pipeline.scan uses."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import torch, torch.distributed as dist
from kobato_b200 import dist as kdist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n * 1e3
    if rank == 0: print(f"{name}: {t:.3f} ms")

rows = torch.randint(0, 1 << 40, (3600 + rank, 2), dtype=torch.int64, device=dev)
hashes = torch.randint(0, 1 << 40, (70000,), dtype=torch.int64, device=dev)
timeit("gather_counts", lambda: kdist._gather_counts(3600 + rank, dev))
timeit("gather_padded flat [3601,2]", lambda: kdist._gather_padded(torch.zeros((3601, 2), dtype=torch.int64, device=dev)))
def listy():
    parts = [torch.empty((3601, 2), dtype=torch.int64, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, torch.zeros((3601, 2), dtype=torch.int64, device=dev)); return parts
timeit("all_gather list [3601,2]", listy)
timeit("all_gather_rows", lambda: kdist.all_gather_rows(rows))
timeit("all_gather_hashes 70k", lambda: kdist.all_gather_hashes(hashes))
timeit("rows sort+cpu", lambda: rows[torch.argsort(rows[:, 0])].cpu().numpy())
imgs = torch.zeros((12, 512, 512, 3), dtype=torch.uint8, device=dev)
def p2p():
    ops = [dist.P2POp(dist.irecv if rank == 0 else dist.isend, imgs[k], 1 - rank) for k in range(12)]
    for r in dist.batch_isend_irecv(ops): r.wait()
timeit("batch p2p 12 images", p2p)
sc = torch.zeros(7190, dtype=torch.float64, device=dev)
timeit("all_reduce scores + cpu", lambda: (dist.all_reduce(sc), sc.cpu()))
dist.barrier(); dist.destroy_process_group()
