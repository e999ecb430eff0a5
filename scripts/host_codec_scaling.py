"""How does the library PNG decoder scale with threads on this host? (cfg 4 is bound by it once the GPU pipeline is deep enough.)
Prints the CPU limits the process actually has (affinity, cgroup quota) and decode throughput for 1..64 threads."""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from robosat_b200 import stitch, synth  # noqa: E402

root = tempfile.mkdtemp(prefix="rsb_codec_")
synth.write_slippy_tiles(os.path.join(root, "tiles"), 18, range(1000, 1016), range(2000, 2032), size=512, seed=7, workers=32)
paths = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(root, "tiles")) for f in fs if f.endswith(".png"))
out = {"os_cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "tiles": len(paths)}
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f):
        out[f] = open(f).read().strip()
bufs = [np.empty((512, 512, 3), np.uint8) for _ in paths]
jobs = list(zip(paths, bufs))
stitch.decode_many(jobs, 512, 8)  # page cache
rates = {}
for t in (1, 4, 8, 16, 24, 32, 48, 64, 96):
    reps = 1 if t == 1 else 3
    t0 = time.perf_counter()
    for _ in range(reps):
        stitch.decode_many(jobs, 512, t)
    rates[t] = round(reps * len(jobs) / (time.perf_counter() - t0), 1)
out["decode_tiles_per_s_by_threads"] = rates
# the shape the shard loop uses: tickets of 40 tiles, one call each
for t in (24, 48):
    t0 = time.perf_counter()
    for i in range(0, len(jobs), 40):
        stitch.decode_many(jobs[i:i + 40], 512, t)
    out["tickets_of_40_with_%d_threads" % t] = round(len(jobs) / (time.perf_counter() - t0), 1)
print(json.dumps(out))
shutil.rmtree(root, ignore_errors=True)
