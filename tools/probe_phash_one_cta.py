import sys, json
sys.path.insert(0, "kobato-eyes_b200")
import torch
from kobato_b200 import _native as nat, ops
ctx = nat.context(0)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for (h, w, c) in ((512, 512, 3), (512, 512, 4), (512, 512, 1), (768, 768, 3)):
    n = 20000 if w == 512 else 9000
    bank = ops.synth_images_device(0, n, h, w, c, n_set=n)
    res = {"auto": timed(lambda: ops.phash_dhash_batch(bank))}
    for cr16 in (0, 1):
        for bufs in (2, 3, 4):
            for sub in (16, 8):
                for shift in (1, 2, 3):
                    v = sub | shift << 8 | bufs << 12 | cr16 << 20 | 1 << 21
                    ctx.set_option(nat.KE_OPT_PHASH_CFG, v)
                    try:
                        res[f"one cr{16 if cr16 else 32} b{bufs} sub{sub} s{shift}"] = timed(lambda: ops.phash_dhash_batch(bank))
                    except Exception:
                        pass
                    finally:
                        ctx.set_option(nat.KE_OPT_PHASH_CFG, 0)
    best = sorted(res.items(), key=lambda kv: kv[1])[:5]
    print(f"{w}x{h}x{c} n={n} auto {res['auto']:.3f} ms; best", [(k, round(v, 3)) for k, v in best], flush=True)
    del bank
