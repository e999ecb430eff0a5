"""Forward time per tile vs batch size (does a smaller batch keep the inter-layer traffic in L2?) -- run on the GPU box."""
import torch

from robosat_b200 import synth
from robosat_b200.engine import UNetEngine

dev = torch.device("cuda:0")
sd = synth.make_state_dict(2, seed=0)
for batch in (32, 16, 8, 4):
    eng = UNetEngine(sd, 2, batch, 512, 512, device=dev)
    xs = [synth.make_tiles_u8(batch, 512, seed=s).to(dev) for s in (1, 2)]
    for i in range(5):
        eng.forward(xs[i % 2])
    torch.cuda.synchronize()
    reps = 640 // batch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        eng.forward(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # per-segment: time the ops up to the end of layer1 / layer2 separately
    names = [op[1].name if op[0] == "conv" else op[0] for op in eng.ops]
    cut1 = max(i for i, n in enumerate(names) if n.startswith("resnet.layer1")) + 1
    cut2 = max(i for i, n in enumerate(names) if n.startswith("resnet.layer2")) + 1
    ops = eng.ops
    seg = []
    for lo, hi in ((0, cut1), (cut1, cut2), (cut2, len(ops))):
        eng.ops = ops[lo:hi]
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            eng.forward(xs[i % 2])
        b.record()
        torch.cuda.synchronize()
        seg.append(a.elapsed_time(b) / reps)
    eng.ops = ops
    print("batch %2d: %.3f ms/forward, %.4f ms/tile (%.0f tiles/s); per tile: stem..layer1 %.4f, layer2 %.4f, rest %.4f" % (
        batch, ms, ms / batch, batch / ms * 1e3, seg[0] / batch, seg[1] / batch, seg[2] / batch), flush=True)
    del eng
    torch.cuda.empty_cache()
