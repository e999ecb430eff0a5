"""Device-side training augmentation against the reference's PIL joint transforms
(robosat/transforms.py:127-221 as composed by robosat/tools/train.py:253-258)."""

import random

import numpy as np
import pytest
import torch
from PIL import Image

from robosat_b200 import augment, synth
from robosat_b200.transforms import JointRandomHorizontalFlip, JointRandomRotation


def _reference_pipeline(image, mask, rng_seed):
    """the reference's four random transforms in its order, driven by Python's `random` (same draws as augment.draw_ops)"""
    random.seed(rng_seed)
    images = [image]
    for t in (JointRandomHorizontalFlip(0.5), JointRandomRotation(0.5, 90), JointRandomRotation(0.5, 90), JointRandomRotation(0.5, 90)):
        images, mask = t(images, mask)
    return images[0], mask


def test_draw_ops_consumes_random_like_the_reference():
    """one draw per transform and sample, in the reference's order: flip, rotate, rotate, rotate"""
    random.seed(7)
    ops = augment.draw_ops(5)
    random.seed(7)
    want = []
    for _ in range(5):
        flip = random.random() < 0.5
        k = sum(1 for _ in range(3) if random.random() < 0.5) % 4
        want.append(int(flip) | (k << 1))
    assert ops == want and all(0 <= o < 8 for o in ops)


@pytest.mark.gpu
def test_device_augmentation_bit_exact_vs_pil(cuda_device):
    S = 64
    u8 = synth.make_tiles_u8(8, S, seed=11)
    masks = (synth.make_masks(8, S, 6, seed=12)).to(torch.uint8)
    aug = augment.DeviceAugmenter(8, S, device=cuda_device)
    # all eight dihedral outcomes explicitly ...
    ops = list(range(8))
    out_img, out_mask = aug.augment(u8.to(cuda_device), masks.to(cuda_device), ops=ops)
    torch.cuda.synchronize()
    for n, op in enumerate(ops):
        im, mk = Image.fromarray(u8[n].numpy()), Image.fromarray(masks[n].numpy(), mode="P")
        if op & 1:
            im, mk = im.transpose(Image.FLIP_LEFT_RIGHT), mk.transpose(Image.FLIP_LEFT_RIGHT)
        for _ in range(op >> 1):
            im, mk = im.transpose(Image.ROTATE_90), mk.transpose(Image.ROTATE_90)
        assert np.array_equal(out_img[n].cpu().numpy(), np.asarray(im)), op
        assert np.array_equal(out_mask[n].cpu().numpy(), np.asarray(mk).astype(np.int64)), op
    # ... and the random pipeline, sample by sample, against the reference's transforms under the same seed
    for n in range(4):
        random.seed(100 + n)
        op = augment.draw_ops(1)
        oi, om = aug.augment(u8[n:n + 1].contiguous().to(cuda_device), masks[n:n + 1].contiguous().to(cuda_device), ops=op)
        ri, rm = _reference_pipeline(Image.fromarray(u8[n].numpy()), Image.fromarray(masks[n].numpy(), mode="P"), 100 + n)
        assert np.array_equal(oi[0].cpu().numpy(), np.asarray(ri)) and np.array_equal(om[0].cpu().numpy(), np.asarray(rm).astype(np.int64))
    assert out_mask.dtype == torch.int64
