"""Seeded synthetic weights, tiles and masks for parity tests and bench.py.

There is no network (no ImageNet weights, no imagery), so every test and
benchmark runs on synthetic data of the reference's shapes.

`make_state_dict` produces a checkpoint `state_dict` with exactly the key
names, order, shapes and dtypes of the reference model
(`robosat/unet.py:94-108` on top of torchvision `resnet50`): 329 entries when
prefixed with `module.` (the `nn.DataParallel` wrapper, `robosat/tools/train.py:69`).
It does NOT call any reference code, so the GPU box (which has no
`/root/reference`) regenerates bit-identical weights from the seed.

Weights are scaled like a trained network rather than a fresh init: He fan-in
convolutions, BatchNorm statistics away from the identity (so the eval-mode BN
folding is exercised) and a damped last BN per bottleneck (so activations do
not grow through the 16 residual blocks).
"""

from collections import OrderedDict

import torch

# torchvision resnet50 topology: Bottleneck x (3, 4, 6, 3), expansion 4, stride on conv2 (v1.5)
RESNET50_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# Per-class median logit of the seed-0 network on seed-1 512x512 tiles (measured once with the
# fp32 reference). Subtracted from `final.bias` so that synthetic predictions contain every class
# and the per-pixel argmax parity test is not vacuous. Other seeds are left uncentred.
_SEED0_LOGIT_MEDIAN = (4.174513, -3.441606, 1.38988, 0.577779, -1.552098, 0.199064)


def unet_param_shapes(num_classes, num_filters=32):
    """Ordered (name, shape, kind) for every state_dict entry of the reference UNet.

    kind: "conv" | "bn_weight" | "bn_bias" | "bn_mean" | "bn_var" | "bn_count" | "fc_w" | "fc_b" | "bias"
    """

    out = []

    def bn(prefix, c):
        out.append((prefix + ".weight", (c,), "bn_weight"))
        out.append((prefix + ".bias", (c,), "bn_bias"))
        out.append((prefix + ".running_mean", (c,), "bn_mean"))
        out.append((prefix + ".running_var", (c,), "bn_var"))
        out.append((prefix + ".num_batches_tracked", (), "bn_count"))

    out.append(("resnet.conv1.weight", (64, 3, 7, 7), "conv"))
    bn("resnet.bn1", 64)

    inplanes = 64
    for li, (planes, blocks, _stride) in enumerate(RESNET50_LAYERS, start=1):
        for b in range(blocks):
            p = "resnet.layer{}.{}".format(li, b)
            out.append((p + ".conv1.weight", (planes, inplanes, 1, 1), "conv"))
            bn(p + ".bn1", planes)
            out.append((p + ".conv2.weight", (planes, planes, 3, 3), "conv"))
            bn(p + ".bn2", planes)
            out.append((p + ".conv3.weight", (planes * 4, planes, 1, 1), "conv"))
            bn(p + ".bn3", planes * 4)
            if b == 0:
                out.append((p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1), "conv"))
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4

    out.append(("resnet.fc.weight", (1000, 2048), "fc_w"))
    out.append(("resnet.fc.bias", (1000,), "fc_b"))

    nf = num_filters
    out.append(("center.block.block.weight", (nf * 8, 2048, 3, 3), "conv"))
    out.append(("dec0.block.block.weight", (nf * 8, 2048 + nf * 8, 3, 3), "conv"))
    out.append(("dec1.block.block.weight", (nf * 8, 1024 + nf * 8, 3, 3), "conv"))
    out.append(("dec2.block.block.weight", (nf * 2, 512 + nf * 8, 3, 3), "conv"))
    out.append(("dec3.block.block.weight", (nf * 4, 256 + nf * 2, 3, 3), "conv"))
    out.append(("dec4.block.block.weight", (nf, nf * 4, 3, 3), "conv"))
    out.append(("dec5.block.weight", (nf, nf, 3, 3), "conv"))
    out.append(("final.weight", (num_classes, nf, 1, 1), "conv"))
    out.append(("final.bias", (num_classes,), "bias"))
    return out


def make_state_dict(num_classes=2, seed=0, prefix="module.", num_filters=32):
    """Deterministic reference-format `state_dict` (fp32 / int64, OIHW convolutions)."""

    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = OrderedDict()

    for name, shape, kind in unet_param_shapes(num_classes, num_filters):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif kind == "bn_weight":
            lo, hi = (0.15, 0.45) if name.endswith("bn3.weight") else (0.6, 1.4)
            t = torch.rand(shape, generator=g) * (hi - lo) + lo
        elif kind == "bn_bias":
            t = torch.randn(shape, generator=g) * 0.15
        elif kind == "bn_mean":
            t = torch.randn(shape, generator=g) * 0.2
        elif kind == "bn_var":
            t = torch.rand(shape, generator=g) * 1.0 + 0.5
        elif kind == "bn_count":
            t = torch.tensor(1000, dtype=torch.int64)
        elif kind == "fc_w":
            t = torch.randn(shape, generator=g) * 0.01
        elif kind == "fc_b":
            t = torch.zeros(shape)
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.1
            if int(seed) == 0 and num_filters == 32 and shape[0] <= len(_SEED0_LOGIT_MEDIAN):
                t = t - torch.tensor(_SEED0_LOGIT_MEDIAN[: shape[0]], dtype=torch.float32)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[prefix + name] = t

    return sd


def make_tiles_u8(n, size, seed=1):
    """Synthetic RGB tiles as uint8 NHWC: smooth low-frequency blobs plus pixel noise."""

    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    coarse = torch.rand((n, 3, max(size // 32, 1), max(size // 32, 1)), generator=g)
    smooth = torch.nn.functional.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
    noise = torch.rand((n, 3, size, size), generator=g)
    img = (0.75 * smooth + 0.25 * noise).clamp_(0, 1)
    return (img * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def normalize_tiles(tiles_u8):
    """uint8 NHWC -> fp32 NCHW exactly as the reference's predict transform does.

    `ToTensor` (u8 -> f32 / 255) then `Normalize(mean, std)` (`robosat/tools/predict.py:71-73`).
    """

    x = tiles_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub(mean).div(std).contiguous()


def make_masks(n, size, num_classes=2, seed=3):
    """Synthetic int64 label masks [n, size, size]: blocky blobs, every class present."""

    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    coarse = torch.randint(0, num_classes, (n, 1, max(size // 16, 1), max(size // 16, 1)), generator=g)
    return torch.nn.functional.interpolate(coarse.float(), size=(size, size), mode="nearest").long().squeeze(1)


def write_slippy_tiles(root, z, x_range, y_range, size=512, seed=7, workers=16, fmt="png"):
    """A synthetic slippy-map directory `root/z/x/y.png` over a contiguous x/y grid (so that every interior tile has its 8
    neighbours, SURVEY.md 8(d) cfg 4): lossless PNGs of `make_tiles_u8` imagery. Returns the number of tiles written."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    xs, ys = list(x_range), list(y_range)
    pool_imgs = make_tiles_u8(64, size, seed=seed).numpy()  # 64 distinct images, reused round-robin with a flip/rotation per tile

    def one(job):
        i, x, y = job
        a = pool_imgs[i % 64]
        k = (i // 64) % 8
        a = a[::-1] if k & 1 else a
        a = a[:, ::-1] if k & 2 else a
        a = a.transpose(1, 0, 2) if k & 4 else a
        d = os.path.join(root, str(z), str(x))
        os.makedirs(d, exist_ok=True)
        Image.fromarray(a.copy()).save(os.path.join(d, "%d.%s" % (y, fmt)), compress_level=1)

    jobs = [(i, x, y) for i, (x, y) in enumerate((x, y) for x in xs for y in ys)]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(one, jobs))
    return len(jobs)
