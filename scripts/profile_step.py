"""One predict step (batch 32 of 3x512x512) after 3 warm-up steps: the target command for ncu captures.

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_tc|conv_row|maxpool|prepass|head_quant' -s 192 -c 64 \
        --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from robosat_b200 import synth  # noqa: E402
from robosat_b200.predictor import TilePredictor  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
pred = TilePredictor(synth.make_state_dict(2, seed=0), 2, batch, 512, overlap=0, device=dev)
x = synth.make_tiles_u8(batch, 512, seed=1).to(dev)
q = torch.empty((batch, 512, 512), dtype=torch.uint8, device=dev)
for _ in range(4):
    pred.quantize(pred.logits(x), q)
torch.cuda.synchronize()
names = [op[1].name if op[0] == "conv" else op[0] for op in pred.engine.ops] + ["head_quantize"]
print("LAUNCH_ORDER " + ",".join(names))
