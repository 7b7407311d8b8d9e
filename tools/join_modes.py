"""Time the K2 join in its three modes (POPC only / hybrid / bit-sliced only)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import numpy as np, torch
from kobato_b200 import _native as nat, ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = nat.context(0); lib = nat.load()
h = torch.from_numpy(synth.synth_hashes(n).view(np.int64)).cuda()
cap = 1 << 22
oi = torch.empty(cap, dtype=torch.int32, device="cuda"); oj = torch.empty_like(oi)
od = torch.empty(cap, dtype=torch.uint8, device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
def run():
    nat.check(lib.ke_hamming_join(ctx.handle, h.data_ptr(), n, 8, 0, 16, 4, None, 0, 1, oi.data_ptr(), oj.data_ptr(),
                                  od.data_ptr(), cap, cnt.data_ptr(), int(torch.cuda.current_stream().cuda_stream)), "join")
pairs = n * (n - 1) // 2
for mode, name in ((1, "popc"), (2, "hybrid"), (3, "sliced")):
    ctx.set_option(nat.KE_OPT_JOIN_MODE, mode)
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{name}: n={n} ms={min(ts):.2f} -> {pairs / (min(ts) * 1e-3):.3e} pairs/s hits={int(cnt.item())}")
