"""Bring-up / measurement of the strict (hi/lo split) precision path on a B200: per-case errors of single convolutions against
float64 references (incl. the sign of the mean error: a truncating accumulator shows up as a bias), whole-network error vs the
fp32 oracle, and step time per precision. Usage: python scripts/gpu_strict_check.py [size ...]"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import conv_cases  # noqa: E402
from oracle import unet_oracle  # noqa: E402
from robosat_b200 import _lib, synth  # noqa: E402
from robosat_b200.engine import UNetEngine  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def run_case(case):
    plan = ctypes.c_void_p()
    _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan")
    _lib.check(lib.rsb_conv_run(plan, _lib.current_stream_ptr()), "run")
    torch.cuda.synchronize()
    got, ref = case.result().double(), case.ref().double()
    lib.rsb_conv_plan_destroy(plan)
    err = got - ref
    scale = ref.abs().max().item()
    nz = ref.abs() > 0.05 * scale
    rel_signed = (err[nz] / ref[nz]).mean().item() if nz.any() else 0.0
    print("%-44s max|err|/max %.3e  rms/rms %.3e  mean signed rel %.2e" % (
        case.name, err.abs().max().item() / scale, (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(), rel_signed), flush=True)


print("== single convolutions, strict precision vs float64")
for mk in conv_cases.split_cases(dev):
    run_case(mk())

sizes = [int(a) for a in sys.argv[1:]] or [64, 256]
for size in sizes:
    batch = 2
    sd = synth.make_state_dict(2, seed=0)
    u8 = synth.make_tiles_u8(batch, size, seed=1)
    x = synth.normalize_tiles(u8)
    t0 = time.time()
    with torch.no_grad():
        ref, feats = unet_oracle.unet_forward(sd, x, return_features=True)
    print("== whole network %dx3x%dx%d (oracle %.1fs)" % (batch, size, size, time.time() - t0))
    for prec in ("fast", "strict"):
        eng = UNetEngine(sd, 2, batch, size, size, device=dev, precision=prec)
        got = eng.forward(x.to(dev)).float().cpu()
        err = (got - ref).abs()
        mism = int((got.argmax(1) != ref.argmax(1)).sum())
        print("  %-6s logits rel_l2 %.3e  max|err|/max|logit| %.3e  argmax mismatches %d / %d" % (
            prec, (err.pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item(), err.max().item() / ref.abs().max().item(), mism, got[:, 0].numel()))
        line = []
        for name in ("stem", "enc1", "enc2", "enc3", "enc4", "center", "dec0", "dec1", "dec2", "dec3", "dec4"):
            a, b = eng.feature_nchw(name), feats[name]
            line.append("%s %.1e" % (name, ((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt()).item()))
        print("         " + "  ".join(line), flush=True)
        del eng

# step time, batch 32 x 512^2 (BASELINE cfg 2)
sd = synth.make_state_dict(2, seed=0)
xd = synth.make_tiles_u8(32, 512, seed=1).to(dev)
for prec in ("fast", "strict"):
    eng = UNetEngine(sd, 2, 32, 512, 512, device=dev, precision=prec)
    for _ in range(3):
        eng.forward(xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(xd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("== %s: %.3f ms / batch of 32 x 512^2 = %.0f tiles/s" % (prec, ms, 32e3 / ms), flush=True)
    # per-layer times
    rows = []
    for op in eng.ops:
        if op[0] != "conv":
            continue
        e0.record()
        for _ in range(3):
            op[1].run(_lib.current_stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        rows.append((op[1].name, e0.elapsed_time(e1) / 3 * 1e3, op[1].desc.block_n if hasattr(op[1].desc, "block_n") else 0, getattr(op[1].desc, "cta_pair", 0)))
    print("   per conv (us, back-to-back x3): " + "  ".join("%s[%d%s] %.0f" % (n.replace("resnet.", ""), bn, "p" if pr else "", t) for n, t, bn, pr in rows), flush=True)
    del eng
