"""K2 on a table of REAL pHashes at the 8-GPU step's size (560 k = 8 x 70 k images of ONE global synthetic set), as one
of eight ranks sees it (tiles t % 8 == 0), next to a random table of the same size; then K1 and K2 launched together
on two streams (does the join hide under the HBM-bound hash kernel?).

    python tools/probe_join_real.py [images_per_rank=70000] [ranks=8]
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
import numpy as np
import torch

from kobato_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.cuda.set_device(0)
bank = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device="cuda")
shards = []
for r in range(world):
    for lo in range(0, n, 8192):
        c = min(8192, n - lo)
        ops.synth_images_device(r + lo * world, c, 512, 512, 3, n_set=n * world, stride=world, out=bank[lo:lo + c])
    shards.append(ops.phash_dhash_batch(bank)[0].clone())
real = torch.cat(shards)
rand = torch.from_numpy(synth.synth_hashes(real.numel()).view(np.int64)).cuda()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = ev()
        fn()
        b = ev()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


out = {"hashes": int(real.numel()), "as_rank_0_of": world}
N = real.numel()
pairs_share = N * (N - 1) / 2 / world
for name, table in (("real", real), ("random", rand)):
    for band in (True, False):
        r = {}
        ms = timed(lambda: r.setdefault("x", ops.hamming_join_device(table, 8, require_band=band, part_index=0, part_count=world,
                                                                     capacity=1 << 24)))
        hits = ops.hamming_join_device(table, 8, require_band=band, part_index=0, part_count=world, capacity=1 << 24)[0].numel()
        out[f"{name}_band{int(band)}"] = {"ms": round(ms, 3), "pairs_per_s": pairs_share / (ms * 1e-3), "hits": int(hits)}
# K1 alone, K2 alone, both at once on two streams
side = torch.cuda.Stream()
k1 = timed(lambda: ops.phash_dhash_batch(bank))
k2 = timed(lambda: ops.hamming_join_device(real, 8, require_band=True, part_index=0, part_count=world, capacity=1 << 24))


def both():
    side.wait_stream(torch.cuda.current_stream())
    ops.phash_dhash_batch(bank)  # asynchronous on the main stream; the join's count read-back below only waits for `side`
    with torch.cuda.stream(side):
        ops.hamming_join_device(real, 8, require_band=True, part_index=0, part_count=world, capacity=1 << 24)
    torch.cuda.current_stream().wait_stream(side)


out["k1_ms"], out["k2_ms"], out["k1_and_k2_on_two_streams_ms"] = round(k1, 3), round(k2, 3), round(timed(both), 3)
print(json.dumps(out))
