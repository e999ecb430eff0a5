"""`rs train` end to end on one GPU: synthetic slippy-map dataset + TOML configs in; log lines, history and a
reference-format checkpoint out (robosat/tools/train.py:115-160), then `--checkpoint ... --resume` continues from it."""

import argparse
import os

import numpy as np
import pytest
import torch
from PIL import Image

from robosat_b200 import colors, synth

pytestmark = pytest.mark.gpu


def _write_split(root, split, n, size, seed):
    rng = np.random.RandomState(seed)
    tiles = synth.make_tiles_u8(n, size, seed=seed).numpy()
    masks = synth.make_masks(n, size, 2, seed=seed + 1).numpy().astype(np.uint8)
    for i in range(n):
        y = 105000 + int(rng.randint(0, 5))  # image and label of a tile share z/x/y
        for sub, arr, mode in (("images", tiles[i], "RGB"), ("labels", masks[i], "P")):
            d = os.path.join(root, split, sub, "18", str(69000 + i))
            os.makedirs(d, exist_ok=True)
            img = Image.fromarray(arr, mode=mode if mode == "P" else None)
            if mode == "P":
                img.putpalette(colors.make_palette("denim", "orange"))
            img.save(os.path.join(d, "%d.png" % y))


def test_rs_train_end_to_end_and_resume(tmp_path, cuda_device, monkeypatch):
    from robosat_b200.tools import train

    ds = tmp_path / "ds"
    _write_split(str(ds), "training", 4, 128, 1)
    _write_split(str(ds), "validation", 2, 128, 5)
    ck = tmp_path / "pth"
    (tmp_path / "model.toml").write_text("[common]\ncuda = true\nbatch_size = 2\nimage_size = 128\ncheckpoint = '%s'\n[opt]\nepochs = 2\nlr = 0.0005\nloss = 'Lovasz'\n" % ck)
    (tmp_path / "dataset.toml").write_text("[common]\ndataset = '%s'\nclasses = ['background', 'parking']\ncolors = ['denim', 'orange']\n[weights]\nvalues = [1.6248, 5.762827]\n" % ds)
    monkeypatch.setenv("RSB_GPUS", "1")
    start = tmp_path / "start.pth"
    torch.save({"epoch": 0, "state_dict": synth.make_state_dict(2, seed=0), "optimizer": {}}, start)
    args = argparse.Namespace(model=str(tmp_path / "model.toml"), dataset=str(tmp_path / "dataset.toml"), checkpoint=str(start), resume=False, workers=0)
    train.main(args)
    log = open(ck / "log").read()
    assert "--- Hyper Parameters on Dataset:" in log and "Epoch: 2/2" in log and "Train    loss:" in log and "Validate loss:" in log and "parking IoU:" in log
    last = torch.load(ck / "checkpoint-00002-of-00002.pth", map_location="cpu")
    assert sorted(last.keys()) == ["epoch", "optimizer", "state_dict"] and last["epoch"] == 2
    ref = synth.make_state_dict(2, seed=0)
    assert list(last["state_dict"].keys()) == list(ref.keys())
    assert len(last["optimizer"]["state"]) == 168 and last["optimizer"]["param_groups"][0]["lr"] == 0.0005
    moved = sum(int(not torch.equal(last["state_dict"][k], ref[k])) for k in ref if "resnet.fc" not in k)
    assert moved >= 320  # every trained tensor and every BN buffer changed
    assert torch.equal(last["state_dict"]["module.resnet.fc.weight"], ref["module.resnet.fc.weight"])
    # resume: picks up epoch and optimiser state; asking for epochs already reached is an error like the reference's
    (tmp_path / "model.toml").write_text(open(tmp_path / "model.toml").read().replace("epochs = 2", "epochs = 3"))
    args = argparse.Namespace(model=str(tmp_path / "model.toml"), dataset=str(tmp_path / "dataset.toml"), checkpoint=str(ck / "checkpoint-00002-of-00002.pth"), resume=True, workers=0)
    train.main(args)
    assert os.path.exists(ck / "checkpoint-00003-of-00003.pth")
    # cross-entropy branch of the loss selection works through the same loop
    (tmp_path / "model.toml").write_text(open(tmp_path / "model.toml").read().replace("loss = 'Lovasz'", "loss = 'CrossEntropy'").replace("epochs = 3", "epochs = 1"))
    args = argparse.Namespace(model=str(tmp_path / "model.toml"), dataset=str(tmp_path / "dataset.toml"), checkpoint=None, resume=False, workers=0)
    train.main(args)
