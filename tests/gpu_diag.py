"""GPU bring-up diagnostics: run every single-convolution case in its own subprocess (so a hung kernel
costs one timeout, not the whole call) and append one JSON line per case to gpurun_out/diag.jsonl.

    python tests/gpu_diag.py            # driver: all cases, then the small whole-network check
    python tests/gpu_diag.py --case 3   # worker
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]  # a test-side tool: it may use the oracle
OUT = os.path.join(ROOT, "gpurun_out")


def emit(rec):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as fp:
        fp.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def stats(a, b):
    d = (a - b).abs()
    return {"max_abs": float(d.max()), "ref_absmax": float(b.abs().max()), "rel_l2": float(d.pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-30))}


def run_case(i):
    import torch
    import conv_cases
    from robosat_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    case = conv_cases.default_cases(dev)[i]()
    rec = {"case": i, "name": case.name, "block_n": case.desc.block_n, "tile": [case.desc.TW, case.desc.TH, case.desc.TN]}
    ref = case.ref()
    stream = _lib.current_stream_ptr()
    # 1) SIMT checker (plain loads): validates descriptor / packing logic
    _lib.check(lib.rsb_conv_run_simt_check(ctypes.byref(case.desc), stream), "simt")
    torch.cuda.synchronize()
    simt = case.result()
    rec["simt_vs_ref"] = stats(simt, ref)
    case.out.zero_()
    emit(dict(rec, stage="simt_done"))
    # 2) tensor-core path
    plan = ctypes.c_void_p()
    rc = lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan))
    if rc != 0:
        rec["plan_error"] = _lib.last_error()
        emit(dict(rec, stage="plan_failed"))
        return
    _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
    torch.cuda.synchronize()
    tc = case.result()
    rec["tc_vs_ref"] = stats(tc, ref)
    rec["tc_vs_simt"] = stats(tc, simt)
    # run twice more to catch pipeline-state bugs that only show on reuse
    case.out.zero_()
    _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
    _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
    torch.cuda.synchronize()
    rec["tc_rerun_vs_simt"] = stats(case.result(), simt)
    emit(dict(rec, stage="tc_done"))


def run_net(size, batch):
    import torch
    from robosat_b200 import synth
    from robosat_b200.engine import UNetEngine
    from oracle import unet_oracle

    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(2, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(batch, size, seed=1))
    t0 = time.time()
    eng = UNetEngine(sd, 2, batch, size, size, device=dev)
    rec = {"net": [batch, size], "build_s": time.time() - t0, "launches": eng.num_launches()}
    emit(dict(rec, stage="net_built"))
    logits = eng.forward(x.to(dev))
    torch.cuda.synchronize()
    got = logits.float().cpu()
    with torch.no_grad():
        ref, feats = unet_oracle.unet_forward(sd, x, return_features=True)
    rec["logits"] = stats(got, ref)
    rec["argmax_mismatch"] = int((got.argmax(1) != ref.argmax(1)).sum())
    rec["pixels"] = int(ref[:, 0].numel())
    layers = {}
    for name in eng.feats:
        key = name if name in feats else None
        if name.startswith("resnet.layer") or name == "pool4":
            continue
        if key:
            layers[name] = stats(eng.feature_nchw(name), feats[key])
    rec["layers"] = layers
    emit(dict(rec, stage="net_done"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--net", type=int, nargs=2, default=None, metavar=("SIZE", "BATCH"))
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    if args.case is not None:
        return run_case(args.case)
    if args.net is not None:
        return run_net(*args.net)
    import conv_cases

    n = len(conv_cases.default_cases(None))
    jobs = [["--case", str(i)] for i in range(n)] + [["--net", "64", "2"], ["--net", "256", "2"]]
    for job in jobs:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + job, timeout=args.timeout, capture_output=True, text=True)
            if r.returncode != 0:
                emit({"job": job, "rc": r.returncode, "stderr": r.stderr[-1500:], "stdout": r.stdout[-500:]})
        except subprocess.TimeoutExpired as e:
            emit({"job": job, "timeout": args.timeout, "stdout": (e.stdout or b"")[-500:].decode("utf-8", "replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-500:]})
        print("job", job, "%.1fs" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
