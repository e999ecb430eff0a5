"""Benchmark of the `rs predict` hot path (BASELINE.json configs[1]): ResNet50-UNet, 2 classes,
synthetic 3x512x512 tiles, batch 32 per GPU. One "step" = one tile batch through the whole path
(pre-pass, 60 convolution launches, 2 max-pools, softmax/crop/quantise head).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Prints ONE JSON line (rank 0). `value` = tiles/s with inputs resident in HBM, device-timed with CUDA events;
`e2e` = tiles/s through the public host API (TilePredictor: pinned host uint8 tiles in, uint8 foreground bins
out, copies inside the timed region); `roofline` = executed tensor FLOP/s of the dominant convolution kernel
against the measured bf16 peak; `cpu_baseline` = the reference's CPU algorithm (oracle port, torch CPU fp32)
on a bounded sample of the same workload. `--impl reference` times that CPU path alone.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 512
BATCH = 32
CLASSES = 2
FWD_GFLOP_DENSE = 167.160  # per 3x512x512 tile, dense-equivalent
WORKLOAD = "rs predict: ResNet50-UNet, 2-class, 3x512x512 synthetic tiles, batch=32 per GPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "tflops_burst": d.get("bf16_tflops"), "hbm_gbs": d.get("hbm_gbs"), "src": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(steps, warmup, tiles_per_step=2, threads=None):
    """The reference's CPU path (oracle restatement: same torch CPU fp32 ops, unet.py:110-141 + softmax) on a bounded sample."""
    import torch

    from oracle import unet_oracle
    from robosat_b200 import synth

    # torchrun exports OMP_NUM_THREADS=1: ask for the cores this process may run on, explicitly. More threads than the box can
    # really schedule (cgroup quotas, SMT) make oneDNN slower, not faster, so the thread count is the fastest of a few candidates
    # on a quick 256x256 probe -- the reference gets the best host configuration, not the nominal one.
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sd = synth.make_state_dict(CLASSES, seed=0)
    if threads is None:
        probe = synth.normalize_tiles(synth.make_tiles_u8(1, 256, seed=2))
        best = None
        for t in sorted({avail, max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16), min(avail, 8)}, reverse=True):
            torch.set_num_threads(t)
            unet_oracle.predict_probs(sd, probe)
            t0 = time.perf_counter()
            unet_oracle.predict_probs(sd, probe)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, t)
        threads = best[1]
    torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    x = synth.normalize_tiles(synth.make_tiles_u8(tiles_per_step, TILE, seed=1))
    for _ in range(warmup):
        unet_oracle.predict_probs(sd, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        unet_oracle.predict_probs(sd, x)
    dt = time.perf_counter() - t0
    return {"value": steps * tiles_per_step / dt, "unit": "tiles/s", "cores": cores, "kind": "port",
            "sample": "%d steps x %d tiles of 3x%dx%d, torch CPU fp32, %d threads (fastest of the candidates <= %d available)" % (
                steps, tiles_per_step, TILE, TILE, cores, avail)}, dt / steps


def run_reference(args, rank):
    if rank != 0:
        return
    cb, s_per_step = cpu_reference_leg(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "512x512 tiles/sec (predict fwd)", "value": cb["value"], "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def layer_profile(engine, x, reps=3):
    """Per-launch device times (CUDA events on the launch stream) -> dominant conv kernel's executed TFLOP/s."""
    import torch

    from robosat_b200 import _lib

    lib = _lib.load()
    stream = _lib.current_stream_ptr()
    engine.forward(x)
    torch.cuda.synchronize()
    rows = []
    for op in engine.ops:
        if op[0] != "conv":
            continue
        c = op[1]
        d = c.desc
        # executed MACs: every output pixel-phase x Cout x K (padded K blocks and tile padding are executed too,
        # but only the algorithmic part is counted as useful work)
        if hasattr(d, "taps_h"):  # line-buffer plan
            K, phases, kern = d.taps_h * d.taps_w * d.cin, d.nsub * d.nphase_a, "conv_row_kernel<%d,%d,%d>" % (32 if d.cin == 32 else 64, d.Cout, d.mode)
        else:
            K, phases, kern = 64 * sum(d.segs[i].cblocks for i in range(d.nseg)), d.phases, "conv_tc_kernel<%d,%d,%d,%d>" % (d.block_n, d.mode, 1 if d.residual else 0, d.cta_pair)
        flops = 2.0 * d.Nt * d.Ht * d.Wt * phases * d.Cout * K
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record()
            c.run(stream)
            b.record()
        torch.cuda.synchronize()
        ms = min(a.elapsed_time(b) for a, b in evs)
        rows.append({"name": c.name, "kernel": kern, "block_n": getattr(d, "block_n", d.Cout), "mode": d.mode, "ms": ms, "gflop": flops / 1e9,
                     "tflops": flops / ms / 1e9, "tiles": c.info()["tiles"], "kblocks": c.info()["kblocks"]})
    return rows


def serve_leg(dev, requests=100, warmup=10, size=512):
    """`rs serve` request latency (serve.py:135-172): one 3x512x512 tile per call through
    `SegmentEngine.run` -- pinned uint8 H2D, normalise, U-Net, argmax, uint8 D2H, host sync -- replayed as one CUDA graph,
    and the same kernels launched one by one on the stream. Secondary figure beside the headline."""
    import time

    import torch

    from robosat_b200 import synth
    from robosat_b200.serve import SegmentEngine

    sd = synth.make_state_dict(2, seed=0)
    tiles = synth.make_tiles_u8(4, size, seed=5).numpy()
    threads = torch.get_num_threads()
    torch.set_num_threads(1)  # the CPU-baseline leg leaves 100+ OpenMP workers spinning; a request is single-threaded host work
    out = {"metric": "rs serve latency per 512x512 tile (batch 1, host to host)", "unit": "ms", "requests": requests}
    for key, use_graph in (("graph_ms", True), ("stream_ms", False)):
        eng = SegmentEngine(sd, 2, size, size, device=dev, use_graph=use_graph)
        if use_graph and eng.graph is None:
            out["graph_error"] = eng.graph_error
            continue
        h_in = eng.h_in.numpy()
        for i in range(warmup):
            h_in[0] = tiles[i % 4]
            eng.run()
        t0 = time.perf_counter()
        for i in range(requests):
            h_in[0] = tiles[i % 4]
            eng.run()
        out[key] = (time.perf_counter() - t0) * 1e3 / requests
        del eng, h_in
        torch.cuda.empty_cache()
    torch.set_num_threads(threads)
    return out


def stitch_leg(dev, batch=32, size=512, overlap=32, reps=50):
    """Halo stitch on the device (tiles.py:162-227): one `rsb_stitch_halo` launch builds the buffered
    batch uint8 [32][576][576][3] from the device tile cache. HBM-bound byte work: algorithmic bytes = canvas read + written."""
    import torch

    from robosat_b200 import _lib

    lib = _lib.load()
    F = size + 2 * overlap
    store = torch.randint(0, 256, (9 * batch, size, size, 3), dtype=torch.uint8, device=dev)
    table = torch.arange(9 * batch, dtype=torch.int32, device=dev).reshape(batch, 9).contiguous()
    out = torch.empty((batch, F, F, 3), dtype=torch.uint8, device=dev)
    st = _lib.current_stream_ptr()
    for _ in range(3):
        _lib.check(lib.rsb_stitch_halo(store.data_ptr(), table.data_ptr(), out.data_ptr(), batch, size, overlap, st), "rsb_stitch_halo")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(lib.rsb_stitch_halo(store.data_ptr(), table.data_ptr(), out.data_ptr(), batch, size, overlap, st), "rsb_stitch_halo")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbps = 2 * out.numel() / ms / 1e6
    return {"metric": "halo stitch of 32 buffered 576x576 tiles on the device", "ms_per_batch": ms, "achieved_gbps": gbps,
            "peak_gbps": peaks()["hbm_gbs"], "frac": gbps / peaks()["hbm_gbs"], "tiles_per_s": batch / ms * 1e3}


def train_leg(dev, steps=8, warmup=3, batch=16):
    """BASELINE.json configs[2] (rs train: 2-class, Lovasz, 3x512x512, batch 16, 1 GPU): one step = zero_grad + train-mode forward
    + Lovasz loss + backward + Adam through the public module API. Reported beside the headline, never instead of it."""
    import torch

    from robosat_b200 import synth
    from robosat_b200.losses import LovaszLoss2d
    from robosat_b200.optim import Adam
    from robosat_b200.unet import UNet

    net = torch.nn.DataParallel(UNet(CLASSES, pretrained=False), device_ids=[dev.index]).to(dev)
    net.load_state_dict(synth.make_state_dict(CLASSES, seed=0))
    opt = Adam(net.parameters(), lr=1e-4)
    opt.mark_used([not n.startswith("module.resnet.fc.") for n, _ in net.named_parameters()])
    crit = LovaszLoss2d().to(dev)
    xs = [synth.normalize_tiles(synth.make_tiles_u8(batch, TILE, seed=300 + i)).to(dev) for i in range(2)]
    ms = [synth.make_masks(batch, TILE, CLASSES, seed=310 + i).to(dev) for i in range(2)]
    net.train()

    def step(i):
        opt.zero_grad()
        loss = crit(net(xs[i % 2]), ms[i % 2])
        loss.backward()
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    del net, opt
    torch.cuda.empty_cache()
    return {"metric": "512x512 tiles/sec (train fwd+bwd, Lovasz, Adam)", "value": batch * steps / (ms_total / 1e3), "unit": "tiles/s",
            "ms_per_step": ms_total / steps, "steps": steps, "warmup": warmup, "batch": batch, "dense_equiv_tflops": batch * steps * 500.246 / ms_total,
            "config": "rs train: ResNet50-UNet, 2-class, Lovasz loss, 3x512x512 synthetic tiles+masks, batch=16, 1xB200", "last_loss": float(loss.detach())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement (configs[2])")
    ap.add_argument("--layers-out", default=None, help="write the per-layer timing table (JSON) here")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist

    from robosat_b200 import synth
    from robosat_b200.predictor import TilePredictor

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weights: rank 0 materialises the checkpoint; ONE NCCL broadcast of the flat fp32 state_dict (the path's only collective)
    sd = synth.make_state_dict(CLASSES, seed=0)
    if world > 1:
        from robosat_b200.dist import broadcast_state_dict

        sd = broadcast_state_dict(sd if rank == 0 else None, template=sd, device=dev)  # rank 0 is the only one whose copy is used

    pred = TilePredictor(sd, CLASSES, BATCH, TILE, overlap=0, device=dev)
    n_in = 4  # rotate distinct input batches; activations (~2.5 GB per step) already exceed the 126 MB L2 many times over
    inputs = [synth.make_tiles_u8(BATCH, TILE, seed=100 + rank * 10 + i).to(dev) for i in range(n_in)]
    qbuf = torch.empty((BATCH, TILE, TILE), dtype=torch.uint8, device=dev)

    def step(i):
        pred.quantize(pred.logits(inputs[i % n_in]), qbuf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the host API: pinned host tiles in, uint8 bins out, copies inside the timed region
    host_batches = [synth.make_tiles_u8(BATCH, TILE, seed=200 + rank * 10 + i).pin_memory() for i in range(2)]
    for i in range(3):
        pred.predict_u8(host_batches[i % 2])
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    acc = 0
    for i in range(args.steps):
        pred.submit(host_batches[i % 2])
        if i >= 1:
            acc += int(pred.collect()[0, 0, 0])
    acc += int(pred.collect()[0, 0, 0])
    e3.record()
    barrier()
    e2e_ms = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3 * 0.0)  # device clock; wall clock kept for sanity below
    wall_ms = (time.perf_counter() - t0) * 1e3

    if world > 1:
        t = torch.tensor([ms, e2e_ms, wall_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, wall_ms = t.tolist()

    line = None
    if rank == 0:
        pk = peaks()
        rows = layer_profile(pred.engine, inputs[0])
        by_kernel = {}
        for r in rows:
            k = r["kernel"]
            a = by_kernel.setdefault(k, {"ms": 0.0, "gflop": 0.0, "launches": 0})
            a["ms"] += r["ms"]
            a["gflop"] += r["gflop"]
            a["launches"] += 1
        dom = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
        dk = by_kernel[dom]
        conv_ms = sum(r["ms"] for r in rows)
        conv_tf = sum(r["gflop"] for r in rows) / conv_ms
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            ent = json.load(open(tpath)).get("kernels", {}).get(dom)
            if ent:
                traffic = {"value": ent["mb_per_launch"], "unit": "MB per launch (ncu dram read+write)", "source": "profiles/ncu_traffic.json"}
        roof = {"bound": "tensor", "kernel": dom, "achieved": dk["gflop"] / dk["ms"], "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": dk["gflop"] / dk["ms"] / pk["tflops"], "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
                "launches_per_step": dk["launches"], "ms_per_step": dk["ms"], "all_conv_tflops": conv_tf, "all_conv_ms": conv_ms,
                "flops": "executed (sub-pixel decoder: 100.7 GFLOP/tile, not the 167.16 dense-equivalent)"}
        # every instantiation's share of the step and its executed tensor rate (the 1x1 layers with K <= 128 are HBM-bound)
        roof["by_kernel"] = {k: {"ms_per_step": round(v["ms"], 4), "launches": v["launches"], "tflops": round(v["gflop"] / v["ms"], 1)}
                             for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1]["ms"])}
        if args.layers_out:
            with open(args.layers_out, "w") as fp:
                json.dump({"layers": rows, "by_kernel": by_kernel}, fp, indent=1)
        tiles = world * BATCH * args.steps
        value = tiles / (ms / 1e3)
        line = {"metric": "512x512 tiles/sec (predict fwd)", "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "parallelism": "tile shards, dp%d, 1 weight broadcast" % world,
                           "l2": "inputs rotate over %d batches; per-step activations ~2.5 GB >> 126 MB L2" % n_in},
                "dense_equiv_tflops": value * FWD_GFLOP_DENSE / 1e3,
                "e2e": {"value": tiles / (e2e_ms / 1e3), "unit": "tiles/s", "h2d_bytes_per_step": pred.h2d_bytes, "d2h_bytes_per_step": pred.d2h_bytes,
                        "wall_ms": wall_ms, "api": "TilePredictor.submit/collect (pinned uint8 in, uint8 bins out)"},
                "gpu_launches": world * args.steps * pred.num_launches(), "clocks": clocks, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:  # the CPU baseline is an N=1 figure (rank 0 only)
            cb, _ = cpu_reference_leg(steps=2, warmup=1, tiles_per_step=2)
            line["cpu_baseline"] = cb
        if world == 1 and not args.no_train:
            del pred
            torch.cuda.empty_cache()
            try:
                line["train"] = train_leg(dev)
            except Exception as exc:  # the headline must survive a failure of the secondary measurement
                line["train"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            try:
                line["stitch"] = stitch_leg(dev)
            except Exception as exc:
                line["stitch"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            try:
                line["serve"] = serve_leg(dev)
            except Exception as exc:
                line["serve"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
