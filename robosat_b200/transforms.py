"""Image / mask transformations with the reference's names (robosat/transforms.py:14-221)."""

import random

import numpy as np
import torch
from PIL import Image


class ImageToTensor:
    """PIL RGB image -> fp32 CHW in [0, 1] (what torchvision's ToTensor does for uint8 images)."""

    def __call__(self, image):
        arr = np.array(image, dtype=np.uint8)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        return torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1).float().div(255)


class ImageToUint8Tensor:
    """PIL RGB image -> uint8 HWC tensor: the raw form the B200 pre-pass normalises on the device."""

    def __call__(self, image):
        return torch.from_numpy(np.array(image.convert("RGB"), dtype=np.uint8))


class MaskToTensor:
    """PIL mask -> int64 HW tensor of class indices."""

    def __call__(self, image):
        return torch.from_numpy(np.array(image, dtype=np.uint8)).long()


class ConvertImageMode:
    def __init__(self, mode):
        self.mode = mode

    def __call__(self, image):
        return image.convert(self.mode)


class JointCompose:
    """Chain of joint (images, mask) transformations."""

    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, images, mask):
        for t in self.transforms:
            images, mask = t(images, mask)
        return images, mask


class JointTransform:
    """Lift independent image / mask transformations (either may be None) to a joint one."""

    def __init__(self, image_transform, mask_transform):
        self.image_transform = image_transform
        self.mask_transform = mask_transform

    def __call__(self, images, mask):
        if self.image_transform is not None:
            images = [self.image_transform(v) for v in images]
        if self.mask_transform is not None:
            mask = self.mask_transform(mask)
        return images, mask


class _JointRandomTranspose:
    def __init__(self, p, method):
        self.p = p
        self.method = method

    def __call__(self, images, mask):
        if random.random() < self.p:
            return [v.transpose(self.method) for v in images], mask.transpose(self.method)
        return images, mask


class JointRandomVerticalFlip(_JointRandomTranspose):
    def __init__(self, p):
        super().__init__(p, Image.FLIP_TOP_BOTTOM)


class JointRandomHorizontalFlip(_JointRandomTranspose):
    def __init__(self, p):
        super().__init__(p, Image.FLIP_LEFT_RIGHT)


class JointRandomRotation(_JointRandomTranspose):
    _METHODS = {90: Image.ROTATE_90, 180: Image.ROTATE_180, 270: Image.ROTATE_270}

    def __init__(self, p, degree):
        if degree not in self._METHODS:
            raise NotImplementedError("We only support multiple of 90 degree rotations for now")
        super().__init__(p, self._METHODS[degree])
