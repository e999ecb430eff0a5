"""Host CPU capacity this process can really use.

`os.cpu_count()` and the affinity mask report the machine's hardware threads; a container is usually also held to a CPU-time
quota (cgroup `cpu.max`, e.g. "1600000 100000" = 16 cores' worth on a 128-thread host). Sizing thread pools by the
hardware-thread count under such a quota makes the kernel throttle the whole group for the rest of every 100 ms period once
the quota is spent: the thread that launches GPU work stalls together with the codec threads (measured on the cfg-4 loop,
profiles/r2_cfg4.md). Everything in this package that sizes a pool asks `usable_cores()`.
"""

import math
import os


def _cgroup_quota():
    """CPU quota in cores (float) from cgroup v2 or v1, or None when unlimited / unreadable"""
    try:
        with open("/sys/fs/cgroup/cpu.max") as fp:  # v2: "<quota|max> <period>"
            quota, period = fp.read().split()[:2]
        if quota != "max" and int(period) > 0:
            return int(quota) / int(period)
        return None
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fp:
            quota = int(fp.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fp:
            period = int(fp.read())
        if quota > 0 and period > 0:
            return quota / period
    except (OSError, ValueError):
        pass
    return None


def usable_cores():
    """min(affinity mask, cgroup CPU quota), at least 1"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = _cgroup_quota()
    if quota is not None:
        n = min(n, max(1, int(math.floor(quota + 1e-9))))
    return max(1, n)
