"""Whole-network parity on the GPU: UNetEngine (librsb200.so) vs the CPU fp32 oracle and the committed
reference fixtures.

Tolerances (written here as the task requires). The product computes with fp16 operands and fp32
accumulation -- the same 10-bit-mantissa operand class as the reference's own CUDA path
(torch.backends.cudnn.allow_tf32 defaults to True) -- so against the fp32 CPU reference we require
    relative L2 error of the logits        <= 5e-3
    max |err| / max |logit|                <= 1e-2
    per-pixel argmax: identical wherever the reference margin |l1 - l0| exceeds 4 * max|err|; total mismatches <= 0.5 %
The measured values are printed (and recorded in DESIGN.md); the north-star's 1e-3 is met in the L2 sense only
on uncentred logits -- see DESIGN.md "Numerics".
"""

import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle
from robosat_b200 import synth
from robosat_b200.engine import UNetEngine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(got, ref, what):
    err = (got - ref).abs()
    max_err = err.max().item()
    rel_l2 = (err.pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    rel_max = max_err / ref.abs().max().item()
    mism = got.argmax(1) != ref.argmax(1)
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    n_mism = int(mism.sum())
    print("%s: rel_l2 %.3e rel_max %.3e max_abs %.4f argmax mismatches %d / %d" % (what, rel_l2, rel_max, max_err, n_mism, mism.numel()))
    assert rel_l2 <= 5e-3 and rel_max <= 1e-2, (what, rel_l2, rel_max)
    assert n_mism <= 0.005 * mism.numel(), (what, n_mism)
    if n_mism:
        assert margin[mism].max().item() <= 4 * max_err, "argmax differs on a pixel that is not a near-tie"


@pytest.mark.parametrize("classes", [2, 6])
def test_logits_match_reference_fixture_64(classes, cuda_device):
    gold = np.load(os.path.join(GOLD, "unet_64.npz"))
    sd = synth.make_state_dict(classes, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(2, 64, seed=1))
    eng = UNetEngine(sd, classes, 2, 64, 64, device=cuda_device)
    got = eng.forward(x.to(cuda_device)).float().cpu()
    _check(got, torch.from_numpy(gold["logits_c%d" % classes]), "fixture64 c%d" % classes)


def test_layerwise_and_logits_match_oracle_256(cuda_device):
    sd = synth.make_state_dict(2, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(2, 256, seed=1))
    eng = UNetEngine(sd, 2, 2, 256, 256, device=cuda_device)
    got = eng.forward(x.to(cuda_device)).float().cpu()
    with torch.no_grad():
        ref, feats = unet_oracle.unet_forward(sd, x, return_features=True)
    for name in ("stem", "enc0", "enc1", "enc2", "enc3", "enc4", "center", "dec0", "dec1", "dec2", "dec3", "dec4"):
        a, b = eng.feature_nchw(name), feats[name]
        r = ((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt()).item()
        assert r < 3e-3, (name, r)
    _check(got, ref, "oracle256")
    gold = np.load(os.path.join(GOLD, "unet_stats.npz"))
    assert np.abs(got[:, :, ::16, ::16].numpy() - gold["sample"]).max() <= 1e-2 * np.abs(gold["sample"]).max()


def test_uint8_input_path_equals_float_path(cuda_device):
    """raw uint8 NHWC tiles (normalised on the device) give the same logits as the reference-style fp32 NCHW input"""
    sd = synth.make_state_dict(2, seed=0)
    u8 = synth.make_tiles_u8(2, 128, seed=4)
    eng = UNetEngine(sd, 2, 2, 128, 128, device=cuda_device)
    a = eng.forward(synth.normalize_tiles(u8).to(cuda_device)).clone()
    b = eng.forward(u8.to(cuda_device)).clone()
    assert torch.equal(a, b)


def test_rectangular_and_overlap_sized_input(cuda_device):
    """320 = 256 + 2*32 is what `rs predict --tile_size 256` feeds the net (predict.py:75); non-square works too."""
    sd = synth.make_state_dict(2, seed=0)
    for (n, h, w) in [(1, 320, 320), (3, 64, 192)]:
        g = torch.Generator().manual_seed(h)
        x = torch.randn((n, 3, h, w), generator=g)
        eng = UNetEngine(sd, 2, n, h, w, device=cuda_device)
        got = eng.forward(x.to(cuda_device)).float().cpu()
        with torch.no_grad():
            ref = unet_oracle.unet_forward(sd, x)
        _check(got, ref, "%dx%dx%d" % (n, h, w))


def test_full_size_batch_properties_512(cuda_device):
    """BASELINE config 2 shape (batch 32 of 3x512x512): size-independent properties instead of a CPU re-run.
    (a) batch independence: tile i alone == tile i inside the batch, bit for bit;
    (b) replay idempotence: the same input twice gives identical logits;
    (c) a 2-tile subset agrees with the oracle."""
    sd = synth.make_state_dict(2, seed=0)
    u8 = synth.make_tiles_u8(32, 512, seed=1)
    eng = UNetEngine(sd, 2, 32, 512, 512, device=cuda_device)
    xd = u8.to(cuda_device)
    a = eng.forward(xd).clone()
    b = eng.forward(xd).clone()
    assert torch.equal(a, b)
    eng2 = UNetEngine(sd, 2, 2, 512, 512, device=cuda_device)
    sub = eng2.forward(xd[4:6].contiguous()).clone()
    assert torch.equal(sub, a[4:6])
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, synth.normalize_tiles(u8[4:6]))
    _check(sub.float().cpu(), ref, "512 subset")
