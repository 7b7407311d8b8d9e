#!/bin/bash
# The one-off 45-95 ms step that some bench runs showed (always the SECOND timed step, in the join stage): prints the
# cgroup's CPU-throttle counters around a 60-step bench run, the longest NVML poll and the per-step device times.
# Outcome: neither the CPU quota nor NVML — the caching allocator's cudaMalloc when two generations of result tensors
# were alive for the first time; bench.py's warm-up now keeps the previous result alive like the timed loop does.
cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc
grep -E "nr_periods|nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --join-n 262144 --ssim-pairs 100000 > gpurun_out/stall_probe.json 2> gpurun_out/stall_probe.err
grep -E "nr_periods|nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/stall_probe.json"))
x = d["device_ms_of_each_step"]
print("steps", len(x), "median", sorted(x)[len(x) // 2], "max", max(x), "at", x.index(max(x)), "clocks", d["clocks"])
PY
