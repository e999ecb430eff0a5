"""Per-op device time of one training step (CUDA events around every op of the train engine's two op lists)."""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from robosat_b200 import synth  # noqa: E402
from robosat_b200.train_engine import UNetTrainEngine  # noqa: E402

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 512
dev = torch.device("cuda:0")
params = {k[7:]: v.to(dev) for k, v in synth.make_state_dict(2, seed=0).items()}
eng = UNetTrainEngine(params, 2, B, S, S, device=dev)
x = synth.normalize_tiles(synth.make_tiles_u8(B, S, seed=1)).to(dev)
dl = (torch.randn((B, 2, S, S), generator=torch.Generator().manual_seed(0)) * 1e-6).to(dev)
for _ in range(2):
    eng.forward(x)
    eng.backward(dl)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
agg = collections.defaultdict(float)
for label, use_graph in (("fwd+bwd step, kernel by kernel (10 back to back, no optimizer)", False),
                         ("fwd+bwd step, op lists replayed from CUDA graphs (default)", True)):
    eng.use_graph = use_graph
    for _ in range(4):  # a graph is captured on the third call
        eng.forward(x)
        eng.backward(dl)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        eng.forward(x)
        eng.backward(dl)
    b.record()
    torch.cuda.synchronize()
    agg[label] = a.elapsed_time(b) / 10
rows = []
for phase, ops, kw in (("fwd", eng.fwd_ops, {"x": x}), ("bwd", eng.bwd_ops, {"dlogits": dl})):
    for op in ops:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng._run([op], **kw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        name = op[1].name if op[0] in ("conv", "wgrad") else ""
        if op[0].startswith("bn_"):
            b_ = op[1]
            name = "%s M=%d C=%d (%.0f MB of fp16 z)" % (getattr(b_, "prefix", "?"), b_.M, b_.C, b_.M * b_.C * 2 / 1e6)
        agg[phase + ":" + op[0]] += ms
        if op[0] in ("conv", "wgrad") or op[0].startswith("bn_"):
            rows.append((ms, phase, op[0], name))
print(json.dumps({k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}))
rows.sort(reverse=True)
for ms, phase, kind, name in rows[:120]:
    print("%8.3f ms  %s %-6s %s" % (ms, phase, kind, name))
