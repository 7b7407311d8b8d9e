#!/usr/bin/env python
"""bench.py — duplicate-scan throughput on B200 (BASELINE.json configs[1], "C2").

A *step* is one pass of the hot path over one batch of synthetic decoded images:
    pHash+dHash (K1) -> all-pairs Hamming join, T=8, reference band predicate (K2) -> SSIM
    verification of the candidate pairs at ssim>=0.9 (K3) -> cluster assembly (host, rank 0).
Workload at N=1: 70 000 synthetic 512x512 RGB images (55 GB, resident in HBM; >> L2).  With N>1
every rank holds its own 70 000-image shard (weak scaling); hashes are all-gathered (NCCL) and the
join's triangle tiles and the SSIM pairs are split across ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM, `e2e` = the same
scan from pinned HOST images (H2D copies, result read-back and host cluster assembly inside the
timed region).  See DESIGN.md §Measurement for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "kobato-eyes_b200"))
sys.path.insert(0, str(ROOT))

METRIC = ("dup-scan images/s (pHash+dHash -> all-pairs Hamming T=8 -> SSIM verify), with per-kernel "
          "pHash images/s, Hamming pairs/s, SSIM pairs/s")
H = W = 512
C = 3
# ncu figures of the K3 kernel bench.py times (see profiles/): fraction of cycles with an instruction issued per scheduler
# and warp-instructions per output pixel.  Updated whenever the kernel changes.
K3_ISSUE_FRAC = 0.789
K3_INSTR_PER_OUTPUT = 43.6
K3_ISSUE_SOURCE = "profiles/r1_ncu_k3v2_summary.txt"
IMG_BYTES = H * W * C


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        # the file holds ONE HBM figure (a burst copy, best of 10); K1 is timed inside the sustained, power-capped step
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json: burst copy)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs.

    NVML in a background thread (every 100 ms).  An `nvidia-smi -lms` child was measured to stall kernel
    launches for ~10 ms per query on these boxes (it showed up as a 12 ms bubble in front of the join), so it
    is only the fallback when the NVML binding is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.stop = index, [], None, threading.Event()
        self.source = None
        self.poll_ms_max = 0.0  # longest NVML poll (four queries) seen

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop.is_set():
            try:
                t0 = time.perf_counter()
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.poll_ms_max = max(self.poll_ms_max, (time.perf_counter() - t0) * 1e3)
                self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if rs & bits[k] else "Not Active" for k in self.NAMES])
            except Exception:
                pass
            self.stop.wait(0.1)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(mx),
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "poll_ms_max": round(self.poll_ms_max, 2),
                "source": self.source}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline


def run_reference(args) -> dict:
    """The reference's CPU path on the box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    from oracle.ref_workers import CpuReference

    cores = os.cpu_count() or 1
    ref = CpuReference(H, W, C, unique=args.ref_unique, cores=cores)
    try:
        for _ in range(args.warmup):
            ref.step(args.ref_sample)
        t0 = time.perf_counter()
        stats = [ref.step(args.ref_sample) for _ in range(args.steps)]
        total = time.perf_counter() - t0
    finally:
        ref.close()
    imgs = args.ref_sample * args.steps
    value = imgs / total
    hash_rate = imgs / sum(s["hash_s"] for s in stats)
    ssim_pairs = sum(s["ssim_pairs"] for s in stats)
    ssim_s = sum(s["ssim_s"] for s in stats)
    sample = (f"{args.ref_sample} images/step cycling {args.ref_unique} unique synthetic 512x512 RGB images "
              f"(hash: process pool x{cores}; LSH scan: single process over {stats[-1]['scan_files']} hashes; SSIM on "
              f"{stats[-1]['ssim_pairs']} pairs/step = the GPU step's 3589 candidates per 70000 images, pool x{cores}; "
              f"{ssim_pairs} SSIM pairs timed in all)")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 fixed point + f32 DCT/SSIM", "data": "synthetic",
        "config": {"workload": "C2 sample: pHash+dHash + LSH Hamming scan (T=8) + SSIM verify on CPU", "h": H, "w": W,
                   "c": C, "images_per_step": args.ref_sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                         "phash_images_per_s": hash_rate,
                         "ssim_pairs_per_s": (ssim_pairs / ssim_s) if ssim_s > 0 and ssim_pairs else None},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baseline_subprocess(args) -> dict | None:
    """Run the reference arm in a fresh process (no CUDA context is forked) on a bounded sample."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "10", "--warmup", "1",
           "--ref-sample", str(args.ref_sample), "--ref-unique", str(args.ref_unique)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as exc:  # pragma: no cover
        return {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}


# ----------------------------------------------------------------------------------------------
# CUDA arm


def run_cuda(args) -> dict:
    import numpy as np
    import torch

    from kobato_b200 import _native as nat
    from kobato_b200 import ops, pipeline, synth
    from kobato_b200 import dist as kdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    ctx = nat.context(local)
    hbm_peak, sm_max_mhz, peak_src = peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- synthetic shard, generated on the GPU ------------------------------------------------
    # ONE global set of world x n images; rank r holds the images r, r + world, r + 2 world, ...  The planted
    # near-duplicates (the last 5 % of the set) point at uniformly chosen earlier images, so (world-1)/world of the
    # candidate pairs straddle two ranks and go through the cross-shard SSIM exchange.
    n = args.images
    bank = torch.empty((n, H, W, C), dtype=torch.uint8, device=dev)
    seed = synth.SEED
    for lo in range(0, n, 8192):
        cnt = min(8192, n - lo)
        ops.synth_images_device(rank + lo * world, cnt, H, W, C, n_set=n * world, seed=seed, stride=world,
                                out=bank[lo:lo + cnt])
    torch.cuda.synchronize(dev)

    # ---- device-resident steps (`value`) ------------------------------------------------------
    # the clock sampler starts BEFORE the warm-up: loading NVML and the first query of each kind take driver locks for
    # tens of milliseconds, which showed up as a one-off stall inside the first timed step when it was started there
    clocks = ClockSampler(local)
    clocks.__enter__()
    t_wait = time.perf_counter()
    while len(clocks.rows) < 3 and time.perf_counter() - t_wait < 2.0:  # the first polls are the slow ones: let them pass
        time.sleep(0.05)
    # the warm-up keeps the previous step's result alive while the next one runs, exactly like the timed loop: with the
    # results dropped at once, the first timed step that overlaps two generations of result tensors sent the caching
    # allocator to cudaMalloc in the middle of the join stage — a one-off 45-95 ms step, always the second timed one
    for _ in range(args.warmup):
        out = pipeline.scan(bank, threshold=8, ssim_threshold=0.9)
    barrier()
    clocks.rows.clear()  # keep only the samples taken while the clock runs
    # a generational GC pass of the interpreter in the middle of a step showed up as a one-off ~25 ms launch gap:
    # collect now, keep the collector off while the clock runs
    import gc

    gc.collect()
    gc.disable()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {}
    step_ms = []  # device time of every step's stages (outliers stay visible next to the mean)
    if True:
        e0.record()
        for _ in range(args.steps):
            out = pipeline.scan(bank, threshold=8, ssim_threshold=0.9)
            for k, v in out.stage_ms.items():
                stage[k] = stage.get(k, 0.0) + v
            step_ms.append(round(sum(v for k, v in out.stage_ms.items() if k != "host_assembly"), 3))
        e1.record()
        barrier()
    launches = ctx.launches - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * n * args.steps / (ms_total * 1e-3)
    counts = out.counts

    # ---- the scan just timed, checked against the CPU oracle (outside every timed region) -----------------------
    # hashes of a few of this rank's images, the candidate list on three row stripes of the gathered table, and the SSIM of
    # a dozen candidates — cross-shard ones first — each image regenerated on the CPU from the same counter-based generator
    table_all = kdist.all_gather_hashes(out.phash).cpu().numpy().view(np.uint64)
    scan_check = None
    if rank == 0:
        import oracle
        from oracle import ref_py

        oracle.build()

        def cpu_image(row):  # table row -> (rank, local index) -> global image id of the interleaved set
            r_, k_ = divmod(int(row), n)
            return synth.synth_image(k_ * world + r_, H, W, C, n_set=n * world, seed=seed)

        bad_hash = 0
        for k_ in (0, 1, n // 2, n - 3, n - 2, n - 1):
            want = oracle.signature(cpu_image(k_))
            bad_hash += int(want[0] != int(table_all[k_])) + int(want[1] != int(out.dhash[k_].item()) & ((1 << 64) - 1))
        nt = len(table_all)
        cand_ok, in_stripes = True, 0
        for lo_ in (0, nt // 2, nt - 384):
            wi, wj, wd = oracle.hamming_join(table_all, 8, require_band=True, threads=os.cpu_count() or 1, row_begin=lo_,
                                             row_end=lo_ + 384)
            sel = (out.cand_i >= lo_) & (out.cand_i < lo_ + 384)
            cand_ok = cand_ok and np.array_equal(out.cand_i[sel], wi) and np.array_equal(out.cand_j[sel], wj) and \
                np.array_equal(out.cand_d[sel], wd)
            in_stripes += len(wi)
        cross_sel = np.flatnonzero(out.cand_i // n != out.cand_j // n)
        same_sel = np.flatnonzero(out.cand_i // n == out.cand_j // n)
        pick = np.concatenate([cross_sel[:: max(1, len(cross_sel) // 8)][:8], same_sel[:: max(1, len(same_sel) // 4)][:4]])
        worst = 0.0
        for q in pick.tolist():
            want = ref_py.ssim_of_planes(oracle.to_l(cpu_image(out.cand_i[q])), oracle.to_l(cpu_image(out.cand_j[q])))
            worst = max(worst, abs(float(out.ssim[q]) - want))
        scan_check = {"hash_mismatches_in_6_images": bad_hash, "candidate_stripes_equal_oracle": bool(cand_ok),
                      "candidates_in_stripes": int(in_stripes), "candidates_cross_shard": int(len(cross_sel)),
                      "candidates": int(len(out.cand_i)), "ssim_pairs_checked": int(len(pick)),
                      "ssim_cross_shard_pairs_checked": int(min(8, len(cross_sel))), "ssim_max_abs_err": worst,
                      "ok": bool(bad_hash == 0 and cand_ok and worst <= 1e-5)}

    # ---- per-kernel timings (CUDA events around single launches, inputs >> L2 or L2-flushed) ---
    def timed(fn, reps=3):
        fn()  # one untimed launch: the first one after a different kernel runs ~10 % long (cold instruction / L2 state)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return sum(ts) / len(ts)

    k1_ms = stage["phash"] / args.steps  # live, inside the timed region: one launch per step
    k1_bytes = n * (IMG_BYTES + 16)
    k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
    roof_k1 = {"kernel": "ke_phash_v5_kernel<3,8,1,32> (512x512x3; TMA ring + dp2a luma + mma.sync u8xs8 resample, fragments in registers)", "bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s",
               "frac": k1_gbs / hbm_peak,
               # dram read+write per launch: 787 184 B/image measured by `ncu --set full` on an 8192-image launch of
               # the same kernel (profiles/r2_ncu_k1_512_summary.txt: 6 448 615 352 B), scaled to this launch's image count
               "traffic": int(n * 787184), "traffic_source": "ncu constant 787184 B/image x images (not measured in this run)",
               "peak_source": peak_src,
               "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
               "images_per_s": n / (k1_ms * 1e-3)}

    # K1 on larger photographs (the reference feeds images up to 4096 px, src/utils/image_io.py:60-138): the same
    # kernel, one CTA per SM with every resample fragment in registers (ke_phash_v5_kernel<3,16|32,1,*>); banks of 6 GB >> L2
    roof_k1_large = {}
    for (hh, ww) in ((1024, 1024), (1536, 2048)):
        # ~6 GB banks, a whole number of images per persistent CTA (296 = 2 x 148 covers one and two CTAs per SM)
        cnt_l = max(296, int(6e9) // (hh * ww * C) // 296 * 296)
        big = torch.empty((cnt_l, hh, ww, C), dtype=torch.uint8, device=dev)
        for lo in range(0, cnt_l, 256):
            c_ = min(256, cnt_l - lo)
            ops.synth_images_device(lo, c_, hh, ww, C, n_set=cnt_l, seed=seed + 5, out=big[lo:lo + c_])
        ms_l = timed(lambda: ops.phash_dhash_batch(big))
        gbs_l = cnt_l * (hh * ww * C + 16) / (ms_l * 1e-3) / 1e9
        roof_k1_large[f"{ww}x{hh}x{C}"] = {"images": cnt_l, "ms": ms_l, "images_per_s": cnt_l / (ms_l * 1e-3), "achieved": gbs_l,
                                           "unit": "GB/s", "frac": gbs_l / hbm_peak,
                                           "kernel": f"ke_phash_v5_kernel<3,{16 if ww <= 1100 else 32},1,{32 if ww <= 1100 else 16}>"}
        del big

    # K2 at config C3: 1 M synthetic hashes, T=8, tiles split over the ranks (strong scaling)
    popc_rate, popc_mhz = ops.popc_rate(4096, local)
    hashes = torch.from_numpy(synth.synth_hashes(args.join_n).view(np.int64)).to(dev)
    lib = nat.load()
    cap = 1 << 22
    oi = torch.empty(cap, dtype=torch.int32, device=dev)
    oj = torch.empty(cap, dtype=torch.int32, device=dev)
    od = torch.empty(cap, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    def run_join(threshold=8):
        nat.check(lib.ke_hamming_join(ctx.handle, hashes.data_ptr(), args.join_n, threshold, 0, 16, 4, None, rank, world,
                                      oi.data_ptr(), oj.data_ptr(), od.data_ptr(), cap, cnt.data_ptr(),
                                      int(torch.cuda.current_stream(dev).cuda_stream)), "ke_hamming_join")

    pairs_total = args.join_n * (args.join_n - 1) // 2
    # Integer rooflines at clocks.max.sm.  (a) POPC only: 2 POPC per pair on the measured POPC issue rate.
    # (b) combined XU+ALU bound of the hybrid kernel: POPC role = 2 POPC + 3.5 ALU ops per pair, bit-sliced
    # role = 6.25 ALU ops per pair (200 LOP3 per 32 pairs), ALU pipe = 64 lanes/clk/SM:
    #   max p+q  s.t.  2p <= popc_rate,  3.5p + 6.25q <= 64   (pairs per clock per SM)
    clk = sm_max_mhz * 1e6
    popc_peak = world * ctx.sm_count * popc_rate * clk / 2.0
    p_max = popc_rate / 2.0
    q_max = max(0.0, (64.0 - 3.5 * p_max) / 6.25)
    k2_peak = world * ctx.sm_count * (p_max + q_max) * clk
    by_threshold = {}
    for T in (4, 8, 12):  # config C3 names all three
        run_join(T)
        barrier()
        t_local = timed(lambda: run_join(T))
        tt = torch.tensor([t_local], dtype=torch.float64, device=dev)
        hits = cnt.clone()
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(hits)
        rate = pairs_total / (float(tt.item()) * 1e-3)
        by_threshold[str(T)] = {"ms": float(tt.item()), "pairs_per_s": rate, "frac": rate / k2_peak,
                                "frac_of_popc_only_roofline": rate / popc_peak, "hits": int(hits.item())}
    k2_ms, k2_rate = by_threshold["8"]["ms"], by_threshold["8"]["pairs_per_s"]
    roof_k2 = {"kernel": "ke_join_fused_kernel (POPC role + bit-sliced LOP3 role)", "bound": "int-alu (XU POPC pipe + ALU LOP3 pipe)",
               "achieved": k2_rate, "peak": k2_peak, "unit": "pairs/s", "frac": k2_rate / k2_peak,
               "popc_only_peak": popc_peak, "frac_of_popc_only_roofline": k2_rate / popc_peak,
               "n_hashes": args.join_n, "threshold": 8, "hits": by_threshold["8"]["hits"], "ms": k2_ms,
               "by_threshold": by_threshold,
               "popc_per_clk_per_sm_measured": popc_rate, "sm_clock_mhz_in_microbench": popc_mhz,
               "peak_source": "LP over measured POPC issue rate (XU) and 64 ALU lanes/clk/SM at clocks.max.sm; "
                              "popc_only_peak = sm_count x POPC rate x clock / 2"}

    # Config C5: 10 M hashes (5e13 pairs), T=8, tiles dealt over the ranks, candidates compacted per rank and gathered;
    # a few row stripes of the result are checked against the CPU oracle OUTSIDE the timed launch.
    roof_c5 = None
    if args.c5 or world >= 4:
        n5 = args.c5_n
        h5_host = synth.synth_hashes(n5)
        h5 = torch.from_numpy(h5_host.view(np.int64)).to(dev)
        cap5 = 1 << 21
        o5i = torch.empty(cap5, dtype=torch.int32, device=dev)
        o5j = torch.empty(cap5, dtype=torch.int32, device=dev)
        o5d = torch.empty(cap5, dtype=torch.uint8, device=dev)

        def run_c5():
            nat.check(lib.ke_hamming_join(ctx.handle, h5.data_ptr(), n5, 8, 0, 16, 4, None, rank, world, o5i.data_ptr(),
                                          o5j.data_ptr(), o5d.data_ptr(), cap5, cnt.data_ptr(),
                                          int(torch.cuda.current_stream(dev).cuda_stream)), "ke_hamming_join")

        barrier()
        t5_local = timed(run_c5, reps=1)
        t5 = torch.tensor([t5_local], dtype=torch.float64, device=dev)
        mine5 = int(cnt.item())
        if world > 1:
            torch.distributed.all_reduce(t5, op=torch.distributed.ReduceOp.MAX)
        key5 = (o5i[:mine5].to(torch.int64) & 0xFFFFFFFF) << 32 | (o5j[:mine5].to(torch.int64) & 0xFFFFFFFF)
        rows5 = kdist.all_gather_rows(torch.stack([key5, o5d[:mine5].to(torch.int64)], dim=1))
        check = None
        if rank == 0:
            import oracle

            oracle.build()
            got = rows5[torch.argsort(rows5[:, 0])].cpu().numpy()
            gi, gj, gd = (got[:, 0] >> 32) & 0xFFFFFFFF, got[:, 0] & 0xFFFFFFFF, got[:, 1]
            stripes, ok5, checked = [(0, 192), (n5 // 3, n5 // 3 + 192), (n5 - 200_000, n5 - 200_000 + 192)], True, 0
            for lo5, hi5 in stripes:
                wi, wj, wd = oracle.hamming_join(h5_host, 8, threads=os.cpu_count() or 1, row_begin=lo5, row_end=hi5)
                sel = (gi >= lo5) & (gi < hi5)
                ok5 = ok5 and np.array_equal(gi[sel], wi) and np.array_equal(gj[sel], wj) and np.array_equal(gd[sel], wd)
                checked += len(wi)
            check = {"stripes": stripes, "pairs_in_stripes": int(checked), "equal_to_oracle": bool(ok5),
                     "sorted_unique": bool(np.all(np.diff(got[:, 0]) > 0))}
        p5 = n5 * (n5 - 1) // 2
        r5 = p5 / (float(t5.item()) * 1e-3)
        roof_c5 = {"workload": f"C5: {n5} hashes all-pairs (T=8), tiles over {world} GPU(s), candidates compacted and gathered",
                   "ms": float(t5.item()), "pairs_per_s": r5, "frac": r5 / k2_peak, "frac_of_popc_only_roofline": r5 / popc_peak,
                   "peak": k2_peak, "hits": int(rows5.shape[0]), "oracle_check": check}
        del h5, o5i, o5j, o5d, rows5

    # K3 at config C4's shape: 256x256 'L' crops, bank >> L2, pairs sharded by index
    m_bank = args.ssim_bank
    crops = torch.empty((m_bank, 256, 256, 1), dtype=torch.uint8, device=dev)
    for lo in range(0, m_bank, 16384):
        c_ = min(16384, m_bank - lo)
        ops.synth_images_device(lo, c_, 256, 256, 1, n_set=m_bank, seed=seed + 1, planted=0.5, out=crops[lo:lo + c_])
    g = torch.Generator(device="cpu").manual_seed(synth.SEED + 2 + rank)
    n_pairs = args.ssim_pairs // world  # C4: 5 M pairs of 256x256 'L' crops in all, sharded by pair index
    ia = torch.randint(0, m_bank, (n_pairs,), generator=g).to(dev)
    ib = torch.randint(0, m_bank, (n_pairs,), generator=g).to(dev)
    bank_l = crops[..., 0]
    ops.ssim_batch(bank_l, ia[:1024], ib[:1024])
    barrier()
    k3_ms_local = timed(lambda: ops.ssim_batch(bank_l, ia, ib))
    k3 = torch.tensor([k3_ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(k3, op=torch.distributed.ReduceOp.MAX)
    k3_ms = float(k3.item())
    k3_bytes = n_pairs * (2 * 256 * 256 + 8)
    k3_gbs = k3_bytes / (k3_ms_local * 1e-3) / 1e9
    roof_k3 = {"kernel": "ke_ssim4_kernel<1> (4 output columns per thread, packed FP32)", "bound": "hbm", "achieved": k3_gbs, "peak": hbm_peak, "unit": "GB/s",
               "frac": k3_gbs / hbm_peak, "pairs_per_s": world * n_pairs / (k3_ms * 1e-3), "shape": "256x256 L",
               "pairs": world * n_pairs, "bank_images": m_bank, "ms": k3_ms, "peak_source": peak_src,
               # the kernel is bound by instruction issue, not by HBM: both fractions side by side.  The issue
               # numbers are ncu's for this kernel at this shape (profiles/, `smsp__issue_active`, instructions / outputs)
               "issue_frac_ncu": K3_ISSUE_FRAC, "instr_per_output_ncu": K3_INSTR_PER_OUTPUT,
               "issue_source": K3_ISSUE_SOURCE}
    del crops, bank_l, ia, ib, hashes

    # N1 (the refinement the UI runs after a scan): tile aHash at the UI default (grid 8 x tile 8 -> 64x64
    # BILINEAR planes -> 4096 bits) over the resident bank; HBM bound like K1 (bytes/image = h*w*c + 64*64 + 512)
    def run_n1():
        planes = ops.gray_resize_batch(bank, 64, 64, "bilinear")
        return ops.tile_ahash_bits(planes, 8, 8)

    n1_ms = timed(run_n1)
    n1_bytes = n * (IMG_BYTES + 64 * 64 + 512)
    n1_gbs = n1_bytes / (n1_ms * 1e-3) / 1e9
    roof_n1 = {"kernel": "ke_resize_mma_kernel<3> + ke_tile_bits_kernel (tile aHash 8x8, 64x64 BILINEAR planes)", "bound": "hbm",
               "achieved": n1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": n1_gbs / hbm_peak,
               "images_per_s": n / (n1_ms * 1e-3), "ms": n1_ms, "peak_source": peak_src}

    # ---- end to end from pinned host memory ---------------------------------------------------
    import psutil

    avail = psutil.virtual_memory().available
    budget = int(avail * 0.30 / max(1, min(world, torch.cuda.device_count())))
    n_e2e = max(1024, min(n, budget // IMG_BYTES))
    host = torch.empty((n_e2e, H, W, C), dtype=torch.uint8, pin_memory=True)
    for lo in range(0, n_e2e, 4096):
        hi = min(n_e2e, lo + 4096)
        host[lo:hi].copy_(bank[lo:hi])
    torch.cuda.synchronize(dev)
    dev_bank = bank[:n_e2e]

    # host-feed roofline: the bare pinned host->device copy of the same bytes in the same chunks, on all ranks at once,
    # nothing else running.  e2e cannot beat it; `h2d_frac` says how close the scan's feed comes.
    def bare_h2d():
        for lo in range(0, n_e2e, 2048):
            hi = min(n_e2e, lo + 2048)
            dev_bank[lo:hi].copy_(host[lo:hi], non_blocking=True)

    bare_h2d()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(3):
        bare_h2d()
    p1.record()
    barrier()
    probe_ms = torch.tensor([p0.elapsed_time(p1) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(probe_ms, op=torch.distributed.ReduceOp.MAX)
    probe_gbs = n_e2e * IMG_BYTES / (float(probe_ms.item()) * 1e-3) / 1e9  # per GPU, all ranks copying

    e2e_steps = max(1, min(args.steps, 10))
    e2e_warm = max(1, min(args.warmup, 2))
    for _ in range(e2e_warm):
        eo = pipeline.scan(dev_bank, host_images=host, threshold=8, ssim_threshold=0.9)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        eo = pipeline.scan(dev_bank, host_images=host, threshold=8, ssim_threshold=0.9)
    t1.record()
    barrier()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_stage = {k: round(v, 3) for k, v in eo.stage_ms.items()}
    e2e_step_s = float(e2e_ms.item()) * 1e-3 / e2e_steps
    e2e_value = world * n_e2e / e2e_step_s
    clocks.__exit__(None, None, None)
    gc.enable()
    h2d_gbs = n_e2e * IMG_BYTES / e2e_step_s / 1e9
    e2e = {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(eo.bytes_h2d),
           "d2h_bytes_per_step": int(eo.bytes_d2h), "images_per_step_per_gpu": int(n_e2e), "steps": e2e_steps,
           "h2d_gbs_per_gpu": h2d_gbs, "h2d_probe_gbs_per_gpu": probe_gbs, "h2d_frac": h2d_gbs / probe_gbs,
           "stages_ms_last_step": e2e_stage,
           "h2d_probe": f"bare pinned cudaMemcpyAsync of the same {n_e2e} images in 2048-image chunks, {world} rank(s) at once",
           "api": "kobato_b200.pipeline.scan(host_images=pinned uint8 [n,512,512,3]) -> hashes, candidates, SSIM, clusters"}

    result = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32 fixed point; f64 DCT; int32+f32 SSIM", "data": "synthetic",
        "config": {"workload": "C2: 70k synthetic 512x512 RGB images per GPU: batched pHash+dHash + all-pairs Hamming "
                               "(T=8, band predicate) + SSIM verify (>=0.9)", "images_per_gpu": n, "h": H, "w": W, "c": C,
                   "hamming_threshold": 8, "ssim_threshold": 0.9, "l2": "inputs (55 GB/GPU) exceed L2; no flush needed",
                   "planting": "one global set of n_gpus x 70000 images dealt round-robin: (n_gpus-1)/n_gpus of the "
                               "candidate pairs straddle two ranks",
                   "parallelism": f"image shards x{world}; join tiles t%{world}; SSIM pairs by owner, cross-shard pairs by "
                                  "(i+j) parity with one packed all_to_all of luma planes"},
        "counts": counts, "scan_check": scan_check,
        "stages_ms_per_step": {k: v / args.steps for k, v in stage.items()}, "device_ms_of_each_step": step_ms,
        "roofline": roof_k1, "roofline_phash_large": roof_k1_large, "roofline_join": roof_k2, "roofline_join_c5": roof_c5, "roofline_ssim": roof_k3,
        "roofline_n1": roof_n1,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        del host
        result["cpu_baseline"] = cpu_baseline_subprocess(args)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return result if rank == 0 else {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--images", type=int, default=70000, help="images per GPU (C2: 70000)")
    ap.add_argument("--join-n", type=int, default=1_000_000, help="hash count for the K2 roofline run (C3)")
    ap.add_argument("--ssim-bank", type=int, default=65536, help="256x256 crops in the K3 roofline bank")
    ap.add_argument("--ssim-pairs", type=int, default=5_000_000, help="pairs in the K3 roofline run over all GPUs (C4: 5 M)")
    ap.add_argument("--c5", action="store_true", help="also time config C5 (10 M-hash join) below 4 GPUs")
    ap.add_argument("--c5-n", type=int, default=10_000_000, help="hash count of the C5 join")
    ap.add_argument("--ref-sample", type=int, default=4096, help="images per reference-arm step")
    ap.add_argument("--ref-unique", type=int, default=256, help="unique synthetic images behind the reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "cuda" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU (rank 0 of the children prints the JSON line)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", *sys.argv]
        raise SystemExit(subprocess.call(cmd))
    # Only the JSON line may reach stdout (NCCL / libraries print banners there): park the real stdout.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "cuda":
        print(f"bench.py: warmup {args.warmup} < 3 — numbers from this run are not valid bench values", file=sys.stderr)

    if args.impl == "reference":
        res = run_reference(args)
    else:
        res = run_cuda(args)
    if res:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(res) + "\n").encode())


if __name__ == "__main__":
    main()
