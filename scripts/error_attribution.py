"""Per-layer error attribution of the predict path against the fp32 oracle (CPU study).

Where does the fast precision's 2-3e-3 logits error come from, and what does the strict precision remove? The oracle's
`conv` hook re-runs the reference graph with selected roundings injected (everything else fp32, BatchNorm unfolded), and
tests/emulate.py replays the engine's real plans (folded / pre-summed weights, fp16 activation storage, hi/lo pairs):

  A  weights rounded to fp16 at every conv, activations fp32
  B  conv inputs (activations) rounded to fp16, weights fp32
  C  both (the operand rounding of the fast mode, without folding / pre-summing / fp16 storage effects)
  F  the fast plan as executed (emulated): C + BN folded into the weights before rounding + decoder taps pre-summed before
     rounding + every activation stored as fp16
  S  the strict plan as executed (emulated): hi/lo fp16 pairs for weights and activations (22+ bits), fp32 accumulate

Writes profiles/r2_error_attribution.md. Usage: python scripts/error_attribution.py [size]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import emulate  # noqa: E402
from oracle import unet_oracle  # noqa: E402
from robosat_b200 import synth  # noqa: E402
from robosat_b200.engine import UNetEngine  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = 2
sd = synth.make_state_dict(2, seed=0)
x = synth.normalize_tiles(synth.make_tiles_u8(batch, size, seed=1))
NAMES = ("stem", "enc1", "enc2", "enc3", "enc4", "center", "dec0", "dec1", "dec2", "dec3", "dec4")


def h(t):
    return t.half().float()


def variant(round_w, round_a):
    def conv(xx, w, b, stride, padding):
        return F.conv2d(h(xx) if round_a else xx, h(w) if round_w else w, b, stride=stride, padding=padding)

    with torch.no_grad():
        return unet_oracle.unet_forward(sd, x, conv=conv, return_features=True)


def rel(a, b):
    return ((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt()).item()


with torch.no_grad():
    ref, rfeat = unet_oracle.unet_forward(sd, x, return_features=True)
rows = {}
for tag, (rw, ra) in (("A weights fp16", (True, False)), ("B activations fp16", (False, True)), ("C both", (True, True))):
    lo, ft = variant(rw, ra)
    rows[tag] = [rel(ft[n], rfeat[n]) for n in NAMES] + [rel(lo, ref), int((lo.argmax(1) != ref.argmax(1)).sum())]
    print(tag, rows[tag][-2:], flush=True)
for tag, prec in (("F fast plan (emulated)", "fast"), ("S strict plan (emulated)", "strict")):
    eng = UNetEngine(sd, 2, batch, size, size, device="cpu", plan_only=True, precision=prec)
    lo = emulate.run_engine(eng, x)
    rows[tag] = [rel(eng.feature_nchw(n), rfeat[n]) for n in NAMES] + [rel(lo, ref), int((lo.argmax(1) != ref.argmax(1)).sum())]
    print(tag, rows[tag][-2:], flush=True)

with open(os.path.join(ROOT, "profiles", "r2_error_attribution.md"), "w") as f:
    f.write("# Error attribution of the predict path vs the fp32 oracle (%d x 3x%dx%d, 2 classes; relative L2 per tensor)\n\n" % (batch, size, size))
    f.write("Produced by `scripts/error_attribution.py` on the CPU (oracle `conv` hook + tests/emulate.py replay of the real plans).\n\n")
    f.write("| variant | " + " | ".join(NAMES) + " | logits | argmax flips / %d |\n" % ref[:, 0].numel())
    f.write("|---|" + "---|" * (len(NAMES) + 2) + "\n")
    for tag, v in rows.items():
        f.write("| %s | " % tag + " | ".join("%.1e" % e for e in v[:-1]) + " | %d |\n" % v[-1])
    f.write("\nReading: (1) weight rounding (A) and activation rounding (B) contribute about equally and add in quadrature (C); the error is\n"
            "injected at every one of the ~60 convolutions and grows like a random walk through the encoder, so no small subset of layers can\n"
            "be blamed -- halving it needs the split on (almost) every layer, which is why the strict mode splits all of them.\n"
            "(2) F is within ~10-20 % of C: BN folding, decoder tap pre-summing and fp16 activation storage are second-order next to the operand\n"
            "rounding itself. (3) The strict plan (S) sits at the fp32 round-off level of the oracle itself (profiles/r2_fp32_noise_floor.md);\n"
            "on the GPU the tensor core's truncating fp32 accumulator adds the rest of the measured error (see DESIGN.md, numerics).\n")
print(open(os.path.join(ROOT, "profiles", "r2_error_attribution.md")).read())
