"""Golden fixture for `rs masks` from the REAL reference (run in the build container only):

    python tests/golden/make_golden_masks.py

Runs the unmodified `robosat.tools.masks.main` (/root/reference/robosat/tools/masks.py:29-84: un-quantise -> np.average soft vote
-> argmax -> P-mode PNG) on seeded probability PNGs written the way `rs predict` writes them, and stores inputs and the masks it
produced in tests/golden/masks.npz. Cases: 1, 2, 3 and 5 models, weighted and unweighted, including the end anchors, the 0.5
crossing and exact ties (100/255 + 155/255 = 1). The GPU box has no /root/reference: tests read only the npz.
"""

import argparse
import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
shim = types.ModuleType("mercantile")
shim.Tile = namedtuple("Tile", ["x", "y", "z"])
sys.modules["mercantile"] = shim
sys.path.insert(0, "/root/reference")

from robosat.colors import continuous_palette_for_color  # noqa: E402
from robosat.tools import masks as ref_masks  # noqa: E402

CASES = [(1, None), (2, None), (3, [1.0, 2.0, 0.5]), (5, [0.1, 0.2, 0.3, 0.25, 0.15]), (2, [0.5, 0.5])]


def main():
    rng = np.random.RandomState(0)
    palette = continuous_palette_for_color("pink", 256)
    out = {}
    for ci, (K, w) in enumerate(CASES):
        q = rng.randint(0, 256, size=(K, 96, 80)).astype(np.uint8)
        q[:, 0, :8] = np.array([0, 255, 127, 128, 1, 254, 64, 191], dtype=np.uint8)
        if K == 2:
            q[0, 1, :], q[1, 1, :] = 100, 155
        with tempfile.TemporaryDirectory() as tmp:
            dirs = []
            for k in range(K):
                d = os.path.join(tmp, "probs%d" % k)
                os.makedirs(os.path.join(d, "18", "7"))
                im = Image.fromarray(q[k], mode="P")
                im.putpalette(palette)
                im.save(os.path.join(d, "18", "7", "9.png"), optimize=True)
                dirs.append(d)
            ref_masks.main(argparse.Namespace(masks=os.path.join(tmp, "masks"), probs=dirs, weights=w))
            m = np.array(Image.open(os.path.join(tmp, "masks", "18", "7", "9.png")))
        out["q%d" % ci] = q
        out["w%d" % ci] = np.array(w if w is not None else [], dtype=np.float64)
        out["mask%d" % ci] = m.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)
    print("wrote masks.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
