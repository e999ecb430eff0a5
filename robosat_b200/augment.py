"""Training-side augmentation on the device: the random left-right flip and the three random quarter turns
of the reference's joint transform (robosat/tools/train.py:253-258 -> robosat/transforms.py:127-221), applied to a whole uint8
batch by ONE kernel (`rsb_augment_dihedral`) instead of four PIL transposes per sample in the DataLoader workers.

The random decisions are drawn on the host with Python's `random` exactly like the reference (one `random.random() < p` per
transform and sample, in the reference's order: flip, rotate, rotate, rotate), so the distribution of the eight dihedral
outcomes is the same; only the pixel shuffling moves to the GPU. Bit-exact against PIL (tests/test_augment.py)."""

import random

import torch

from robosat_b200 import _lib


def draw_ops(n, p_flip=0.5, p_rot=0.5, rotations=3, rng=random):
    """op per sample = flip | (quarter turns << 1), drawn like JointRandomHorizontalFlip(p) + `rotations` x JointRandomRotation(p, 90)"""
    ops = []
    for _ in range(n):
        flip = 1 if rng.random() < p_flip else 0
        k = sum(1 for _ in range(rotations) if rng.random() < p_rot) % 4
        ops.append(flip | (k << 1))
    return ops


class DeviceAugmenter:
    """`augment(images u8 [N, S, S, 3], masks u8 [N, S, S]) -> (images u8 [N, S, S, 3], masks int64 [N, S, S])` on the device."""

    def __init__(self, batch, size, device="cuda"):
        self.device = torch.device(device)
        self.out_img = torch.empty((batch, size, size, 3), dtype=torch.uint8, device=self.device)
        self.out_mask = torch.empty((batch, size, size), dtype=torch.int64, device=self.device)
        self._ops_host = torch.zeros(batch, dtype=torch.int32, pin_memory=self.device.type == "cuda")
        self._ops = torch.zeros(batch, dtype=torch.int32, device=self.device)

    def augment(self, images, masks, ops=None):
        n, s = images.shape[0], images.shape[1]
        assert images.dtype == torch.uint8 and tuple(images.shape) == (n, s, s, 3) and images.is_contiguous() and images.device == self.device
        assert masks.dtype == torch.uint8 and tuple(masks.shape) == (n, s, s) and masks.is_contiguous()
        assert n <= self.out_img.shape[0] and s == self.out_img.shape[1]
        ops = draw_ops(n) if ops is None else ops
        self._ops_host[:n] = torch.tensor(ops, dtype=torch.int32)
        self._ops[:n].copy_(self._ops_host[:n], non_blocking=True)
        _lib.check(_lib.load().rsb_augment_dihedral(images.data_ptr(), masks.data_ptr(), self._ops.data_ptr(), self.out_img.data_ptr(),
                                                    self.out_mask.data_ptr(), n, s, _lib.current_stream_ptr()), "rsb_augment_dihedral")
        return self.out_img[:n], self.out_mask[:n]
