"""K1 on larger photographs: time the streaming kernel under pinned staging configurations (KE_OPT_PHASH_CFG).

    python tools/probe_phash_large.py [--auto] [HxWxC ...]   # default: 512 ... 4000-pixel rows; the automatic configuration (+ a sweep)
"""
import json
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


out = {}
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:] if "x" in a]  # e.g. 768x768x3 (h x w x c)
for (h, w, c) in (shapes or ((512, 512, 3), (1024, 1024, 3), (1536, 2048, 3), (3000, 4000, 3))):
    n = max(16, int(4e9) // (h * w * c) // 296 * 296 or 16)
    bank = torch.empty((n, h, w, c), dtype=torch.uint8, device="cuda")
    for lo in range(0, n, 128):
        k = min(128, n - lo)
        ops.synth_images_device(lo, k, h, w, c, n_set=n, out=bank[lo:lo + k])
    res = {}
    cfgs = [("auto", 0)]
    for cr16 in (() if "--auto" in sys.argv else (0, 1)):
        for place in (0, 1, 2, 3):
            for bufs in (4, 3, 2, 1):
                for sub in (16, 8, 4, 2, 1):
                    for shift in (1, 2):
                        cfgs.append((f"{'cr16 ' if cr16 else ''}p{place} b{bufs} sub{sub} s{shift}",
                                     sub | shift << 8 | bufs << 12 | place << 16 | cr16 << 20))
    for name, value in cfgs:
        ctx.set_option(nat.KE_OPT_PHASH_CFG, value)
        try:
            ms = timed(lambda: ops.phash_dhash_batch(bank))
            res[name] = round(n * (h * w * c + 16) / (ms * 1e-3) / 1e9 / PEAK, 3)
        except Exception as exc:  # configuration does not fit
            continue
        finally:
            ctx.set_option(nat.KE_OPT_PHASH_CFG, 0)
    best = sorted(res.items(), key=lambda kv: -kv[1])[:6]
    out[f"{w}x{h}x{c}"] = {"images": n, "auto_frac": res.get("auto"), "best": best}
    print(f"{w}x{h}x{c}", "auto", res.get("auto"), "best", best, flush=True)
    del bank
print(json.dumps(out))
