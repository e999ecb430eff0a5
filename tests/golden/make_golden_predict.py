"""Golden fixture for `rs predict` from the REAL reference tool (run in the build container only):

    python tests/golden/make_golden_predict.py

Builds the synthetic slippy-map directory of tests/test_predict_tool_gpu.py (five 256x256 tiles: a 2x2 block with neighbours and one
isolated tile), the seeded checkpoint and the TOML configs (cuda = false), runs the UNMODIFIED `robosat.tools.predict.main`
(/root/reference/robosat/tools/predict.py:43-113: buffered tiles -> ToTensor / Normalize -> UNet -> softmax -> unbuffer ->
np.digitize -> P-mode PNG) and stores the probability bins it wrote in tests/golden/predict_bins.npz. The GPU box has no
/root/reference: the test reads only the npz.
"""

import argparse
import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
shim = types.ModuleType("mercantile")
shim.Tile = namedtuple("Tile", ["x", "y", "z"])
sys.modules["mercantile"] = shim
sys.path.insert(0, "/root/reference")

import robosat.unet as ref_unet  # noqa: E402
from robosat.tools import predict as ref_predict  # noqa: E402

from robosat_b200 import synth  # noqa: E402

_resnet50 = ref_unet.resnet50
ref_unet.resnet50 = lambda pretrained=True: _resnet50(weights=None)

COORDS = [(100, 200), (101, 200), (100, 201), (101, 201), (103, 205)]  # keep in sync with tests/test_predict_tool_gpu.py


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        tiles_dir, probs_dir = os.path.join(tmp, "tiles"), os.path.join(tmp, "probs")
        u8 = synth.make_tiles_u8(5, 256, seed=9).numpy()
        for (x, y), arr in zip(COORDS, u8):
            os.makedirs(os.path.join(tiles_dir, "17", str(x)), exist_ok=True)
            Image.fromarray(arr).save(os.path.join(tiles_dir, "17", str(x), "%d.png" % y))
        ckpt = os.path.join(tmp, "checkpoint.pth")
        torch.save({"epoch": 1, "state_dict": synth.make_state_dict(2, seed=0), "optimizer": {}}, ckpt)
        open(os.path.join(tmp, "model.toml"), "w").write(
            "[common]\ncuda = false\nbatch_size = 2\nimage_size = 256\ncheckpoint = '%s'\n[opt]\nepochs = 1\nlr = 0.0001\nloss = 'Lovasz'\n" % tmp)
        open(os.path.join(tmp, "dataset.toml"), "w").write(
            "[common]\ndataset = '%s'\nclasses = ['background', 'parking']\ncolors = ['denim', 'orange']\n" % tmp)
        ref_predict.main(argparse.Namespace(batch_size=2, checkpoint=ckpt, overlap=32, tile_size=256, workers=0, tiles=tiles_dir, probs=probs_dir,
                                            model=os.path.join(tmp, "model.toml"), dataset=os.path.join(tmp, "dataset.toml")))
        for x, y in COORDS:
            im = Image.open(os.path.join(probs_dir, "17", str(x), "%d.png" % y))
            assert im.mode == "P" and im.size == (256, 256)
            out["bins_%d_%d" % (x, y)] = np.array(im)
            out["palette"] = np.array(im.getpalette()[:768], dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "predict_bins.npz"), **out)
    print("wrote predict_bins.npz:", sorted(out))


if __name__ == "__main__":
    main()
