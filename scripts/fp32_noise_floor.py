"""CPU study: how far is the reference's own fp32 path (oracle, oneDNN) from the exact (float64) result of the same network?
That distance -- logits error and argmax flips at near-ties -- is the "fp32 noise floor" the parity tests use as the yardstick
for "argmax bit-exact": two correct fp32 implementations differ by this much. Usage: python scripts/fp32_noise_floor.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle  # noqa: E402
from robosat_b200 import synth  # noqa: E402

rows = []
for classes, batch, size in [(2, 2, 64), (6, 2, 64), (2, 2, 256), (2, 1, 320), (2, 2, 512)]:
    sd = synth.make_state_dict(classes, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(batch, size, seed=1))
    with torch.no_grad():
        t0 = time.time()
        lo32 = unet_oracle.unet_forward(sd, x)
        t32 = time.time() - t0
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        t0 = time.time()
        lo64 = unet_oracle.unet_forward(sd64, x.double())
        t64 = time.time() - t0
    err = (lo32.double() - lo64).abs()
    mism = int((lo32.argmax(1) != lo64.argmax(1)).sum())
    rows.append((classes, batch, size, (err.pow(2).sum().sqrt() / lo64.pow(2).sum().sqrt()).item(), err.max().item() / lo64.abs().max().item(), mism,
                 lo64[:, 0].numel(), t32, t64))
    print(rows[-1], flush=True)

with open(os.path.join(ROOT, "profiles", "r2_fp32_noise_floor.md"), "w") as f:
    f.write("# fp32 noise floor of the reference path (CPU oracle fp32 vs the same graph in float64)\n\n")
    f.write("Produced by `scripts/fp32_noise_floor.py` in the build container (torch %s, oneDNN, %d threads).\n\n" % (torch.__version__, torch.get_num_threads()))
    f.write("| classes | batch | size | rel L2 of logits | max err / max logit | argmax flips | pixels |\n|---|---|---|---|---|---|---|\n")
    for c, b, s, l2, mx, mm, px, _, _ in rows:
        f.write("| %d | %d | %d | %.2e | %.2e | %d | %d |\n" % (c, b, s, l2, mx, mm, px))
