"""Measure the tile choice (block_n, CTA pair) of every tcgen05 convolution of the inference plan on this GPU and write the layers
where a candidate beats the modelled choice (engine.choose_block_n / make_conv_desc) by >= 3 % to robosat_b200/plans_b200.json.

Every layer is timed alone (CUDA events, min of 5 launches) in an engine whose layers all use one candidate (layers the
candidate is invalid for keep the modelled choice); then the whole step is timed with and without the table, interleaved.
Run under gpurun: `python scripts/tune_plans.py [--out robosat_b200/plans_b200.json]`.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402  (layer_profile)
from robosat_b200 import synth  # noqa: E402
from robosat_b200.engine import UNetEngine  # noqa: E402

CANDS = [(32, 0), (64, 0), (128, 0), (128, 1), (256, 0), (256, 1)]
SHAPES = [("strict", 32, 512, 512), ("strict", 32, 576, 576), ("fast", 32, 512, 512), ("fast", 32, 576, 576), ("strict", 1, 512, 512)]


def profile(sd, dev, precision, n, h, w, overrides, x):
    eng = UNetEngine(sd, 2, n, h, w, device=dev, precision=precision, plan_overrides=overrides)
    rows = bench.layer_profile(eng, x, reps=5)
    out = {}
    for r, op in zip(rows, [o[1] for o in eng.ops if o[0] == "conv"]):
        d = op.desc
        out[r["name"]] = (r["ms"], getattr(d, "block_n", None), int(getattr(d, "cta_pair", 0)), hasattr(d, "taps_h"))
    del eng
    torch.cuda.empty_cache()
    return out


def step_ms(eng, x, steps=20):
    for _ in range(3):
        eng.forward(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        eng.forward(x)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "robosat_b200", "plans_b200.json"))
    ap.add_argument("--gain", type=float, default=0.03)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(2, seed=0)
    plans, report = {}, {}
    for precision, n, h, w in SHAPES:
        key = "%s:%dx%dx%d" % (precision, n, h, w)
        x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev)
        base = profile(sd, dev, precision, n, h, w, {}, x)
        names = [k for k, v in base.items() if not v[3]]  # tcgen05 implicit-GEMM layers (the line-buffer kernel has no tile choice)
        best = {k: (base[k][0], base[k][1], base[k][2]) for k in names}
        for bn, pair in CANDS:
            got = profile(sd, dev, precision, n, h, w, {k: {"block_n": bn, "cta_pair": pair} for k in names}, x)
            for k in names:
                ms, gbn, gpair, _ = got[k]
                if (gbn, gpair) == (bn, pair) and ms < best[k][0]:
                    best[k] = (ms, bn, pair)
        chosen = {}
        for k in names:
            ms, bn, pair = best[k]
            if (bn, pair) != (base[k][1], base[k][2]) and ms < (1.0 - args.gain) * base[k][0]:
                chosen[k] = {"block_n": bn, "cta_pair": pair, "ms": round(ms, 4), "model_ms": round(base[k][0], 4),
                             "model": [base[k][1], base[k][2]]}
        # whole step, interleaved A/B/A/B
        e0 = UNetEngine(sd, 2, n, h, w, device=dev, precision=precision, plan_overrides={})
        e1 = UNetEngine(sd, 2, n, h, w, device=dev, precision=precision, plan_overrides=chosen)
        t0, t1 = [], []
        for _ in range(3):
            t0.append(step_ms(e0, x))
            t1.append(step_ms(e1, x))
        same = bool((e0.forward(x).argmax(1) == e1.forward(x).argmax(1)).float().mean().item() > 0.9999)
        del e0, e1
        torch.cuda.empty_cache()
        report[key] = {"model_step_ms": round(min(t0), 4), "tuned_step_ms": round(min(t1), 4), "layers_changed": len(chosen),
                       "sum_layer_gain_ms": round(sum(v["model_ms"] - v["ms"] for v in chosen.values()), 4), "argmax_agrees": same}
        if chosen and min(t1) < min(t0) * 0.995:  # keep a table only where the whole step confirms it
            plans[key] = chosen
        print(key, json.dumps(report[key]), flush=True)
    with open(args.out, "w") as fp:
        json.dump({"device": torch.cuda.get_device_name(0), "note": "written by scripts/tune_plans.py; see engine.plan_table()", "report": report,
                   "plans": plans}, fp, indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
