"""Per-step device times of the resident scan loop with and without the NVML clock sampler of bench.py
(is the one-off ~40 ms stall early in the timed loop the sampler's?)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
sys.path.insert(0, str(ROOT))
import torch

import bench
from kobato_b200 import ops, pipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
torch.cuda.set_device(0)
bank = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device="cuda")
for lo in range(0, n, 8192):
    c = min(8192, n - lo)
    ops.synth_images_device(lo, c, 512, 512, 3, n_set=n, out=bank[lo:lo + c])


def loop(label, steps=25):
    for _ in range(3):
        pipeline.scan(bank)
    torch.cuda.synchronize()
    out = []
    for _ in range(steps):
        o = pipeline.scan(bank)
        out.append((round(sum(v for k, v in o.stage_ms.items() if k != "host_assembly"), 2), round(o.stage_ms["join"], 2)))
    worst = max(out)
    print(f"{label}: median {sorted(out)[len(out) // 2][0]} ms, worst {worst}, steps {[t for t, _ in out]}", flush=True)


loop("no sampler")
s = bench.ClockSampler(0)
s.__enter__()
loop("sampler started just before the loop")
loop("sampler running for a while")
s.__exit__(None, None, None)
loop("sampler stopped")
