"""`rs predict` end to end on the GPU: synthetic slippy-map directory + checkpoint + TOML configs in, paletted PNGs out,
compared with the reference pipeline restated on the CPU (buffer_tile_image -> ToTensor/Normalize -> oracle forward ->
softmax -> unbuffer -> np.digitize, robosat/tools/predict.py:71-113)."""

import argparse
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import unet_oracle
from robosat_b200 import colors, synth
from robosat_b200.datasets import BufferedSlippyMapDirectory
from robosat_b200.transforms import ImageToUint8Tensor

pytestmark = pytest.mark.gpu


def test_rs_predict_end_to_end(tmp_path, cuda_device, monkeypatch):
    from robosat_b200.tools import predict

    tiles_dir, probs_dir = tmp_path / "tiles", tmp_path / "probs"
    u8 = synth.make_tiles_u8(5, 256, seed=9).numpy()
    coords = [(100, 200), (101, 200), (100, 201), (101, 201), (103, 205)]  # a 2x2 block with neighbours + one isolated tile
    for (x, y), arr in zip(coords, u8):
        os.makedirs(tiles_dir / "17" / str(x), exist_ok=True)
        Image.fromarray(arr).save(tiles_dir / "17" / str(x) / ("%d.png" % y))
    sd = synth.make_state_dict(2, seed=0)
    ckpt = tmp_path / "checkpoint-00001-of-00001.pth"
    torch.save({"epoch": 1, "state_dict": sd, "optimizer": {}}, ckpt)
    (tmp_path / "model.toml").write_text("[common]\ncuda = true\nbatch_size = 2\nimage_size = 256\ncheckpoint = '%s'\n[opt]\nepochs = 1\nlr = 0.0001\nloss = 'Lovasz'\n" % tmp_path)
    (tmp_path / "dataset.toml").write_text("[common]\ndataset = '%s'\nclasses = ['background', 'parking']\ncolors = ['denim', 'orange']\n" % tmp_path)
    monkeypatch.setenv("RSB_GPUS", "1")
    args = argparse.Namespace(batch_size=2, checkpoint=str(ckpt), overlap=32, tile_size=256, workers=0, tiles=str(tiles_dir), probs=str(probs_dir),
                              model=str(tmp_path / "model.toml"), dataset=str(tmp_path / "dataset.toml"))
    predict.main(args)  # default input path: tiles decoded once, halo stitched on the device (robosat_b200/stitch.py)

    # the reference-shaped input path (buffered tiles assembled on the host) must give byte-identical masks
    probs_host = tmp_path / "probs_host"
    monkeypatch.setenv("RSB_HOST_STITCH", "1")
    host_args = argparse.Namespace(**{**vars(args), "probs": str(probs_host)})
    predict.main(host_args)
    monkeypatch.delenv("RSB_HOST_STITCH")
    for (x, y) in coords:
        a = np.array(Image.open(probs_dir / "17" / str(x) / ("%d.png" % y)))
        b = np.array(Image.open(probs_host / "17" / str(x) / ("%d.png" % y)))
        assert np.array_equal(a, b), (x, y)

    directory = BufferedSlippyMapDirectory(str(tiles_dir), transform=ImageToUint8Tensor(), size=256, overlap=32)
    palette = colors.continuous_palette_for_color("pink", 256)
    worst, total_diff, total_px, argmax_flips = 0, 0, 0, 0
    for i in range(len(directory)):
        image, xyz = directory[i]
        x, y, z = (int(v) for v in xyz)
        out = Image.open(probs_dir / str(z) / str(x) / ("%d.png" % y))
        assert out.mode == "P" and out.size == (256, 256) and out.getpalette()[: 3 * 256] == palette
        with torch.no_grad():
            probs = unet_oracle.predict_probs(sd, synth.normalize_tiles(image.unsqueeze(0))).numpy()[0]
        fg = directory.unbuffer(probs)[1]
        ref = np.digitize(fg, np.linspace(0, 1, 256)).astype(np.uint8)
        diff = np.abs(np.array(out).astype(np.int32) - ref.astype(np.int32))
        worst = max(worst, int(diff.max()))
        total_diff += int((diff > 0).sum())
        total_px += diff.size
        # the argmax the probability image encodes: foreground iff p > 0.5, i.e. bin >= 129 (anchor 128/255 > 0.5)
        argmax_flips += int(((np.array(out) >= 129) != (ref >= 129)).sum())
    print("rs predict vs reference pipeline: worst bin difference %d, pixels whose bin differs %d / %d, argmax flips %d" % (
        worst, total_diff, total_px, argmax_flips))
    # Default (strict) precision: the foreground probability agrees to ~1e-4, so a bin (width 1/255) can only differ where the
    # reference probability sits that close to a bin edge -- by exactly one bin, on a small fraction of the pixels.
    assert worst <= 1 and total_diff <= 0.02 * total_px
    assert argmax_flips <= max(2, 8 * total_px // 131072)

    # and against the files the UNMODIFIED reference tool wrote for the same directory / checkpoint (tests/golden/make_golden_predict.py
    # ran robosat.tools.predict.main in the build container): same palette, same bins up to one bin at bin edges
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "predict_bins.npz"))
    worst_ref, diff_ref = 0, 0
    for (x, y) in coords:
        out = Image.open(probs_dir / "17" / str(x) / ("%d.png" % y))
        assert out.getpalette()[:768] == gold["palette"].tolist()
        d = np.abs(np.array(out).astype(np.int32) - gold["bins_%d_%d" % (x, y)].astype(np.int32))
        worst_ref, diff_ref = max(worst_ref, int(d.max())), diff_ref + int((d > 0).sum())
    print("rs predict vs the reference tool's own output: worst bin difference %d, pixels whose bin differs %d / %d" % (worst_ref, diff_ref, total_px))
    assert worst_ref <= 1 and diff_ref <= 0.02 * total_px


def test_tile_predictor_graph_replay_equals_kernel_by_kernel(cuda_device):
    """`TilePredictor(use_graph=True)` (what `rs predict` uses: network + head captured once per slot, one driver call per batch)
    returns the same bins, bit for bit, as launching the 60 kernels one by one -- for several batches through both slots."""
    from robosat_b200.predictor import TilePredictor

    sd = synth.make_state_dict(2, seed=0)
    tiles = [synth.make_tiles_u8(2, 128, seed=40 + i) for i in range(5)]
    outs = {}
    for use_graph in (False, True):
        pred = TilePredictor(sd, 2, 2, 128, overlap=16, device=cuda_device, use_graph=use_graph)
        if use_graph:
            assert pred.graph_error is None, pred.graph_error
        res = []
        for i, t in enumerate(tiles):
            pred.submit(t)
            if i >= 1:
                res.append(pred.collect().clone())
        res.append(pred.collect().clone())
        outs[use_graph] = res
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a, b)
