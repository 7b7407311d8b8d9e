"""K1 role probes on long rows: which role bounds the streaming kernel at 1024 and 2048 pixels per row?

Needs a library built with -DKE_TUNING_PROBES (KE_LIB_PATH points at it); KE_PHASH_DBG then switches parts of the kernel off:
    8   luma warps hand every raw slot straight back (TMA feed alone)      2   tap warps skip the horizontal MMA loop
    4   tap warps skip the vertical pass                                    6   both (luma conversion + barriers remain)
    32  wide-target warps skip their horizontal MMA loop                    64  narrow-target warps skip theirs

    KE_LIB_PATH=kobato-eyes_b200/csrc/build/libkobato_probe.so python tools/probe_phash_roles.py
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
import torch

from kobato_b200 import _native as nat
from kobato_b200 import ops

torch.cuda.set_device(0)
ctx = nat.context(0)
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6558.4


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for (h, w, c) in ((512, 512, 3), (1024, 1024, 3), (1536, 2048, 3)):
    n = max(296, int(5e9) // (h * w * c) // 296 * 296)
    bank = torch.empty((n, h, w, c), dtype=torch.uint8, device="cuda")
    for lo in range(0, n, 128):
        k = min(128, n - lo)
        ops.synth_images_device(lo, k, h, w, c, n_set=n, out=bank[lo:lo + k])
    line = {}
    for dbg in (0, 8, 2, 4, 6, 32, 64):
        os.environ["KE_PHASH_DBG"] = str(dbg)
        ms = timed(lambda: ops.phash_dhash_batch(bank))
        line[dbg] = round(n * (h * w * c + 16) / (ms * 1e-3) / 1e9 / PEAK, 3)
    os.environ["KE_PHASH_DBG"] = "0"
    print(f"{w}x{h}x{c} n={n} frac of HBM by probe: {line}", flush=True)
    del bank
