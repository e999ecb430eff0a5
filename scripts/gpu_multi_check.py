"""Multi-GPU smoke of the two tools on ONE box (run under `gpurun --gpus 2`):
  * rs predict with RSB_GPUS=2 writes byte-identical PNGs to the single-GPU run (tiles are independent; one weight broadcast)
  * rs train  with RSB_GPUS=2 runs an epoch (one NCCL all-reduce of the flat gradient arena per step) and writes a checkpoint
"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from robosat_b200 import synth  # noqa: E402


def main():
    from robosat_b200.tools import predict, train
    from test_train_tool_gpu import _write_split

    tmp = tempfile.mkdtemp()
    tiles = os.path.join(tmp, "tiles")
    u8 = synth.make_tiles_u8(12, 256, seed=9).numpy()
    for i, arr in enumerate(u8):
        d = os.path.join(tiles, "17", str(100 + i % 4))
        os.makedirs(d, exist_ok=True)
        Image.fromarray(arr).save(os.path.join(d, "%d.png" % (200 + i // 4)))
    sd = synth.make_state_dict(2, seed=0)
    ckpt = os.path.join(tmp, "c.pth")
    torch.save({"epoch": 1, "state_dict": sd, "optimizer": {}}, ckpt)
    open(os.path.join(tmp, "model.toml"), "w").write("[common]\ncuda = true\nbatch_size = 2\nimage_size = 128\ncheckpoint = '%s/pth'\n[opt]\nepochs = 1\nlr = 0.0005\nloss = 'Lovasz'\n" % tmp)
    open(os.path.join(tmp, "dataset.toml"), "w").write("[common]\ndataset = '%s/ds'\nclasses = ['background', 'parking']\ncolors = ['denim', 'orange']\n" % tmp)
    outs = {}
    for world in (1, 2):
        os.environ["RSB_GPUS"] = str(world)
        outs[world] = os.path.join(tmp, "probs%d" % world)
        predict.main(argparse.Namespace(batch_size=2, checkpoint=ckpt, overlap=32, tile_size=256, workers=0, tiles=tiles, probs=outs[world],
                                        model=os.path.join(tmp, "model.toml"), dataset=os.path.join(tmp, "dataset.toml")))
    same = 0
    for root, _, files in os.walk(outs[1]):
        for f in files:
            a = os.path.join(root, f)
            b = a.replace(outs[1], outs[2])
            assert os.path.exists(b) and np.array_equal(np.array(Image.open(a)), np.array(Image.open(b))), f
            same += 1
    print("MULTI predict: %d tiles identical between 1 and 2 GPUs" % same)
    assert same == 12
    _write_split(os.path.join(tmp, "ds"), "training", 8, 128, 1)
    _write_split(os.path.join(tmp, "ds"), "validation", 4, 128, 5)
    os.environ["RSB_GPUS"] = "2"
    train.main(argparse.Namespace(model=os.path.join(tmp, "model.toml"), dataset=os.path.join(tmp, "dataset.toml"), checkpoint=ckpt, resume=False, workers=0))
    last = torch.load(os.path.join(tmp, "pth", "checkpoint-00001-of-00001.pth"), map_location="cpu")
    moved = sum(int(not torch.equal(last["state_dict"][k], sd[k])) for k in sd if "resnet.fc" not in k)
    print("MULTI train: checkpoint written by rank 0, %d tensors updated; log:" % moved)
    print(open(os.path.join(tmp, "pth", "log")).read())
    assert moved >= 320


if __name__ == "__main__":
    main()
