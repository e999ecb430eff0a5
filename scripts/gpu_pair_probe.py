"""A/B timing of the CTA-pair conv schedule on a few layer shapes (run on the GPU box)."""
import ctypes
import sys

import torch

sys.path.insert(0, "tests")
import conv_cases  # noqa: E402
from robosat_b200 import _lib  # noqa: E402


def time_plan(lib, plan, stream, iters=20):
    for _ in range(3):
        lib.rsb_conv_run(plan, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.rsb_conv_run(plan, stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    stream = _lib.current_stream_ptr()
    shapes = [("3x3", 32, 64, 64, 128, 128, 128), ("3x3", 32, 32, 32, 256, 256, 128), ("3x3", 32, 32, 32, 256, 256, 256),
              ("1x1", 32, 32, 32, 1024, 256, 128), ("1x1", 32, 64, 64, 128, 512, 128)]
    for kind, N, H, W, cin, cout, bn in shapes:
        case = conv_cases.conv_case(kind, N, H, W, cin, cout, dev, seed=1, block_n=bn)
        gf = 2.0 * N * H * W * cin * cout * (9 if kind == "3x3" else 1) / 1e9
        for pair in (0, 1):
            case.desc.cta_pair = pair
            plan = ctypes.c_void_p()
            _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan")
            grid, tiles, kb, smem = (ctypes.c_int32() for _ in range(4))
            lib.rsb_conv_plan_info(plan, ctypes.byref(grid), ctypes.byref(tiles), ctypes.byref(kb), ctypes.byref(smem))
            ms = time_plan(lib, plan, stream)
            print("%s N%d %dx%d %d->%d bn%d pair=%d grid=%d tiles=%d kb=%d smem=%d  %.4f ms  %.0f TF" % (
                kind, N, H, W, cin, cout, bn, pair, grid.value, tiles.value, kb.value, smem.value, ms, gf / ms), flush=True)
            lib.rsb_conv_plan_destroy(plan)


if __name__ == "__main__":
    main()
