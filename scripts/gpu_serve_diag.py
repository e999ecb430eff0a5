"""Where does a batch-1 request spend its time? (run on the GPU box)"""
import time

import torch

from robosat_b200 import _lib, synth
from robosat_b200.engine import UNetEngine
from robosat_b200.serve import SegmentEngine

dev = torch.device("cuda:0")
sd = synth.make_state_dict(2, seed=0)
for batch in (1, 4):
    eng = UNetEngine(sd, 2, batch, 512, 512, device=dev)
    x = synth.make_tiles_u8(batch, 512, seed=1).to(dev)
    for _ in range(3):
        eng.forward(x)
    torch.cuda.synchronize()
    # whole forward: GPU time (events) and wall time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        eng.forward(x)
    e1.record()
    t_issue = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / 20 * 1e3
    print("batch %d forward: issue %.3f ms, wall %.3f ms, gpu %.3f ms per call" % (batch, t_issue, t_wall, e0.elapsed_time(e1) / 20), flush=True)
    # per-op GPU time, synchronised between ops
    stream = _lib.current_stream_ptr()
    lib = _lib.load()
    rows = []
    for op in eng.ops:
        if op[0] != "conv":
            continue
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        op[1].run(stream)
        b.record()
        torch.cuda.synchronize()
        rows.append((a.elapsed_time(b) * 1e3, op[1].name))
    rows.sort(reverse=True)
    print("  slowest ops (us):", ", ".join("%s %.0f" % (n, t) for t, n in rows[:8]), " sum %.0f us" % sum(t for t, _ in rows), flush=True)
    del eng
seg = SegmentEngine(sd, 2, 512, 512, device=dev, use_graph=True)
print("graph:", seg.graph is not None, seg.graph_error)
for _ in range(5):
    seg.run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(20):
    seg.graph.replay()
e1.record()
torch.cuda.synchronize()
print("graph replay: wall %.3f ms, gpu %.3f ms per replay" % ((time.perf_counter() - t0) / 20 * 1e3, e0.elapsed_time(e1) / 20))
