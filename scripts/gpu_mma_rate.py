"""tcgen05.mma issue-rate probe: one CTA vs CTA pair (run on the GPU box)."""
import torch

from robosat_b200 import _lib

lib = _lib.load()
out = torch.zeros(148, device="cuda")
st = _lib.current_stream_ptr()
for grid in (2, 148):
    for pair in (0, 1):
        for bn in (64, 128, 256):
            for commit in (0, 1):
                out.zero_()
                _lib.check(lib.rsb_debug_mma_rate(out.data_ptr(), grid, pair, bn, 512, commit, st), "rate")
                torch.cuda.synchronize()
                v = out[:grid:2 if pair else 1]
                print("grid=%3d pair=%d N=%3d commit_each=%d  cycles/MMA min %.1f mean %.1f max %.1f" % (grid, pair, bn, commit, v.min().item(), v.mean().item(), v.max().item()), flush=True)
