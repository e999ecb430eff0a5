"""tcgen05.mma issue-rate probe: one CTA vs CTA pair, K-major vs MN-major operands (run on the GPU box)."""
import torch

from robosat_b200 import _lib

lib = _lib.load()
out = torch.zeros(148, device="cuda")
st = _lib.current_stream_ptr()
for grid in (2, 148):
    for pair in (0, 1):
        for bn in (64, 128, 256):
            for flags in (0, 1, 2):
                out.zero_()
                _lib.check(_lib.load_debug().rsb_debug_mma_rate(out.data_ptr(), grid, pair, bn, 512, flags, st), "rate")
                torch.cuda.synchronize()
                v = out[:grid:2 if pair else 1]
                print("grid=%3d pair=%d N=%3d %-12s cycles/MMA min %.1f mean %.1f max %.1f" % (
                    grid, pair, bn, ("k-major", "k-major+commit", "mn-major")[flags], v.min().item(), v.mean().item(), v.max().item()), flush=True)
