"""Generate the golden fixtures from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/robosat (CPU torch) with the survey's import recipe (mercantile shim, resnet50
patched to pretrained=False -- there is no network) and stores small seeded input/output vectors:

    unet_64.npz      logits of UNet(2) / UNet(6) on 2 tiles of 3x64x64, weights from robosat_b200.synth (seed 0)
    unet_stats.npz   logit checksums at 3x256x256 (too large to commit in full)
    losses.npz       LovaszLoss2d / CrossEntropyLoss2d values + autograd gradients, Metrics counts, Adam steps

The GPU box has no /root/reference: tests read only these files.
"""

import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

shim = types.ModuleType("mercantile")
shim.Tile = namedtuple("Tile", ["x", "y", "z"])
sys.modules["mercantile"] = shim
sys.path.insert(0, "/root/reference")

import robosat.unet as ref_unet  # noqa: E402
from robosat.losses import CrossEntropyLoss2d, FocalLoss2d, LovaszLoss2d, mIoULoss2d  # noqa: E402
from robosat.metrics import Metrics  # noqa: E402

from robosat_b200 import synth  # noqa: E402

_resnet50 = ref_unet.resnet50
ref_unet.resnet50 = lambda pretrained=True: _resnet50(weights=None)


def ref_net(num_classes):
    sd = synth.make_state_dict(num_classes, seed=0)
    net = torch.nn.DataParallel(ref_unet.UNet(num_classes))
    net.load_state_dict(sd)
    net.eval()
    return net


def main():
    torch.manual_seed(0)
    out = {}
    for C in (2, 6):
        net = ref_net(C)
        x = synth.normalize_tiles(synth.make_tiles_u8(2, 64, seed=1))
        with torch.no_grad():
            out["logits_c%d" % C] = net.module(x).numpy()
    np.savez_compressed(os.path.join(HERE, "unet_64.npz"), **out)

    net = ref_net(2)
    x = synth.normalize_tiles(synth.make_tiles_u8(2, 256, seed=1))
    with torch.no_grad():
        lo = net.module(x)
    np.savez_compressed(
        os.path.join(HERE, "unet_stats.npz"),
        sum=np.float64(lo.double().sum().item()),
        abs_sum=np.float64(lo.double().abs().sum().item()),
        sample=lo[:, :, ::16, ::16].numpy(),
        fg_pixels=np.int64((lo.argmax(1) == 1).sum().item()),
        argmax_packed=np.packbits(lo.argmax(1).numpy().astype(np.uint8)),
    )

    # losses / metrics / Adam on seeded logits
    g = torch.Generator().manual_seed(7)
    res = {}
    for tag, (N, C, S) in {"a": (2, 2, 32), "b": (3, 6, 16), "c": (1, 2, 64)}.items():
        logits = (torch.randn((N, C, S, S), generator=g) * 2.0).requires_grad_(True)
        targets = synth.make_masks(N, S, num_classes=C, seed=11)
        loss = LovaszLoss2d()(logits, targets)
        loss.backward()
        res["lovasz_%s_logits" % tag] = logits.detach().numpy()
        res["lovasz_%s_targets" % tag] = targets.numpy()
        res["lovasz_%s_loss" % tag] = np.float32(loss.item())
        res["lovasz_%s_grad" % tag] = logits.grad.numpy().copy()
        logits.grad = None
        w = torch.tensor([1.6248, 5.762827, 1.0, 2.0, 0.5, 3.0][:C])
        loss = CrossEntropyLoss2d(weight=w)(logits, targets)
        loss.backward()
        res["ce_%s_weight" % tag] = w.numpy()
        res["ce_%s_loss" % tag] = np.float32(loss.item())
        res["ce_%s_grad" % tag] = logits.grad.numpy().copy()
        logits.grad = None
        loss = FocalLoss2d(gamma=2, weight=w)(logits, targets)
        loss.backward()
        res["focal_%s_loss" % tag] = np.float32(loss.item())
        res["focal_%s_grad" % tag] = logits.grad.numpy().copy()
        # mIoU loss: once with the logits as they are and once scaled up (confident predictions make the CE term win / lose)
        onehot = torch.zeros(N, C, S, S).scatter_(1, targets.view(N, 1, S, S), 1.0)
        res["miou_aligned_%s_logits" % tag] = (logits.detach() * 0.15 + onehot * 1.2).numpy()
        for sub, scale in (("", 1.0), ("_sharp", 4.0), ("_aligned", None)):
            lg = (logits.detach() * scale if scale is not None else torch.from_numpy(res["miou_aligned_%s_logits" % tag])).requires_grad_(True)
            loss = mIoULoss2d(weight=w)(lg, targets)
            loss.backward()
            res["miou%s_%s_loss" % (sub, tag)] = np.float32(loss.item())
            res["miou%s_%s_grad" % (sub, tag)] = lg.grad.numpy().copy()
        m = Metrics(range(C))
        for mask, output in zip(targets, logits.detach()):
            m.add(mask, output)
        res["metrics_%s" % tag] = np.array([m.tn, m.fn, m.fp, m.tp], dtype=np.int64)
    # Adam: 3 steps of torch.optim.Adam(lr=1e-4) exactly as train.py:81 constructs it
    p = torch.nn.Parameter(torch.randn(1000, generator=g))
    opt = torch.optim.Adam([p], lr=1e-4)
    res["adam_p0"] = p.detach().numpy().copy()
    grads = torch.randn(3, 1000, generator=g)
    res["adam_grads"] = grads.numpy()
    for s in range(3):
        p.grad = grads[s].clone()
        opt.step()
        res["adam_p%d" % (s + 1)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **res)
    # halo stitching (tiles.py:162-227) and the probability palette (colors.py:70-95) on a synthetic 3x3 grid with a hole
    import tempfile

    from PIL import Image
    from robosat.colors import continuous_palette_for_color, make_palette
    from robosat.tiles import buffer_tile_image, tiles_from_slippy_map

    rng = np.random.RandomState(5)
    with tempfile.TemporaryDirectory() as tmp:
        grid = {}
        for x in range(10, 13):
            for y in range(20, 23):
                if (x, y) == (12, 20):
                    continue  # missing neighbour -> nodata border
                arr = rng.randint(0, 256, (16, 16, 3)).astype(np.uint8)
                os.makedirs(os.path.join(tmp, "7", str(x)), exist_ok=True)
                Image.fromarray(arr).save(os.path.join(tmp, "7", str(x), "%d.png" % y))
                grid["tile_%d_%d" % (x, y)] = arr
        tiles = list(tiles_from_slippy_map(tmp))
        centre = [t for t, _ in tiles if (t.x, t.y) == (11, 21)][0]
        corner = [t for t, _ in tiles if (t.x, t.y) == (10, 20)][0]
        grid["buffered_centre_o4"] = np.array(buffer_tile_image(centre, tiles, overlap=4, tile_size=16))
        grid["buffered_corner_o4"] = np.array(buffer_tile_image(corner, tiles, overlap=4, tile_size=16))
        grid["buffered_centre_o0"] = np.array(buffer_tile_image(centre, tiles, overlap=0, tile_size=16))
    grid["palette_pink_256"] = np.array(continuous_palette_for_color("pink", 256), dtype=np.int64)
    grid["palette_denim_orange"] = np.array(make_palette("denim", "orange"), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), **grid)

    for f in ("unet_64.npz", "unet_stats.npz", "losses.npz", "tiles.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
