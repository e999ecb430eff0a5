"""`rs masks` / `rs weights`: the per-pixel work on the GPU is bit-identical to the reference's numpy
(robosat/tools/masks.py:42-84 un-quantise + np.average soft vote + argmax; robosat/tools/weights.py:39-55 np.bincount + 1/ln(1.02+p))."""

import argparse
import os

import numpy as np
import pytest
import torch
from PIL import Image

from robosat_b200 import colors
from robosat_b200.tools import masks, weights


def _softvote_reference(quantised, w):
    """masks.py:47-62 + :72-84 restated: anchors[q] -> [background, foreground] -> np.average over models -> argmax"""
    anchors = np.linspace(0, 1, 256)
    probs = []
    for q in quantised:
        foreground = np.rollaxis(np.expand_dims(anchors[q], axis=0), axis=0)
        background = np.rollaxis(1. - foreground, axis=0)
        probs.append(np.concatenate((background, foreground), axis=0))
    return np.argmax(np.average(probs, axis=0, weights=w), axis=0).astype(np.uint8)


def test_parsers_and_weight_formula():
    parser = argparse.ArgumentParser()
    sub = parser.add_subparsers()
    masks.add_parser(sub)
    weights.add_parser(sub)
    a = parser.parse_args(["masks", "out", "p1", "p2", "--weights", "0.25", "0.75"])
    assert a.masks == "out" and a.probs == ["p1", "p2"] and a.weights == [0.25, 0.75] and a.func is masks.main
    b = parser.parse_args(["weights", "--dataset", "d.toml"])
    assert b.dataset == "d.toml" and b.func is weights.main
    counts, n = np.array([900, 100]), 1000
    want = 1 / np.log(1.02 + counts / n)
    assert weights.weights_from_counts(counts, n) == [round(float(v), 6) for v in want]


@pytest.mark.gpu
def test_softvote_bit_identical_to_numpy(cuda_device):
    rng = np.random.RandomState(0)
    for K, w in [(1, None), (2, None), (3, [1.0, 2.0, 0.5]), (5, [0.1, 0.2, 0.3, 0.25, 0.15]), (2, [0.5, 0.5])]:
        q = rng.randint(0, 256, size=(K, 96, 80)).astype(np.uint8)
        q[:, 0, :8] = np.array([0, 255, 127, 128, 1, 254, 64, 191], dtype=np.uint8)  # end anchors and the 0.5 crossing
        if K == 2:
            q[0, 1, :], q[1, 1, :] = 100, 155                                          # exact ties: 100/255 + 155/255 = 1
        got = masks.softvote_device(torch.from_numpy(q).to(cuda_device), w).cpu().numpy()
        assert np.array_equal(got, _softvote_reference(list(q), w)), (K, w)


def test_softvote_restatement_is_pinned_to_the_real_reference():
    """tests/golden/masks.npz holds masks written by the unmodified `robosat.tools.masks.main` (make_golden_masks.py); the numpy
    restatement used by the GPU test above reproduces them exactly, so that test is pinned to the reference, not to this file."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masks.npz"))
    for ci in range(5):
        w = list(gold["w%d" % ci]) or None
        assert np.array_equal(_softvote_reference(list(gold["q%d" % ci]), w), gold["mask%d" % ci]), ci


@pytest.mark.gpu
def test_softvote_matches_reference_fixture(cuda_device):
    """`rsb_softvote` against the masks the real reference tool produced from the same probability PNGs (bit-identical, ties included)"""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masks.npz"))
    for ci in range(5):
        w = list(gold["w%d" % ci]) or None
        got = masks.softvote_device(torch.from_numpy(gold["q%d" % ci]).to(cuda_device), w).cpu().numpy()
        assert np.array_equal(got, gold["mask%d" % ci]), ci


@pytest.mark.gpu
def test_class_histogram_matches_bincount(cuda_device):
    rng = np.random.RandomState(1)
    counts = None
    total = np.zeros(6, dtype=np.int64)
    for n in (1, 255, 4096, 512 * 512 * 3 + 7):
        lab = rng.randint(0, 6, size=n).astype(np.uint8)
        total += np.bincount(lab, minlength=6)
        counts = weights.class_counts_device(torch.from_numpy(lab).to(cuda_device), 6, counts)
    assert np.array_equal(counts.cpu().numpy(), total)


@pytest.mark.gpu
def test_masks_and_weights_tools_end_to_end(tmp_path, cuda_device, capsys):
    rng = np.random.RandomState(2)
    coords = [(5, 7), (5, 8), (6, 7)]
    dirs = []
    quant = {}
    for k in range(2):
        root = tmp_path / ("probs%d" % k)
        dirs.append(str(root))
        for (x, y) in coords:
            os.makedirs(root / "18" / str(x), exist_ok=True)
            q = rng.randint(0, 256, size=(64, 64)).astype(np.uint8)
            quant[(k, x, y)] = q
            img = Image.fromarray(q, mode="P")
            img.putpalette(colors.continuous_palette_for_color("pink", 256))  # what rs predict writes (predict.py:105-108)
            img.save(root / "18" / str(x) / ("%d.png" % y), optimize=True)
    out = tmp_path / "masks"
    masks.main(argparse.Namespace(masks=str(out), probs=dirs, weights=[0.3, 0.7]))
    for (x, y) in coords:
        got = Image.open(out / "18" / str(x) / ("%d.png" % y))
        assert got.mode == "P"
        assert np.array_equal(np.array(got), _softvote_reference([quant[(0, x, y)], quant[(1, x, y)]], [0.3, 0.7]))
    # rs weights
    labels = tmp_path / "ds" / "training" / "labels"
    allpix = []
    for (x, y) in coords:
        os.makedirs(labels / "18" / str(x), exist_ok=True)
        m = (rng.rand(64, 64) < 0.2).astype(np.uint8)
        allpix.append(m.ravel())
        img = Image.fromarray(m, mode="P")
        img.putpalette(colors.make_palette("denim", "orange"))
        img.save(labels / "18" / str(x) / ("%d.png" % y), optimize=True)
    (tmp_path / "dataset.toml").write_text("[common]\ndataset = '%s'\nclasses = ['background', 'parking']\ncolors = ['denim', 'orange']\n" % (tmp_path / "ds"))
    weights.main(argparse.Namespace(dataset=str(tmp_path / "dataset.toml")))
    printed = capsys.readouterr().out.strip().splitlines()[-1]
    cat = np.concatenate(allpix)
    counts = np.bincount(cat, minlength=2)
    want = 1 / np.log(1.02 + counts / cat.size)
    want.round(6, out=want)
    assert printed == str(want.tolist())
