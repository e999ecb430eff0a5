"""Test infrastructure: fp32 torch forward of the U-Net in TRAIN mode with every ReLU mask and max-pool argmax FROZEN to
the ones a UNetTrainEngine run actually used. Autograd through it is the reference for the engine's backward: it removes
the only discontinuities (mask / argmax flips caused by fp16 forward rounding), so the remaining difference is pure
fp16 rounding and a wrong dgrad / wgrad formulation shows up as an O(1) error.
Topology follows oracle/unet_oracle.py (robosat/unet.py:110-141)."""

import torch
import torch.nn.functional as F

BLOCKS = (3, 4, 6, 3)


def _mask(eng, name):
    t, (n, h, w, c) = eng.relu_outs[name]
    m = (t.detach().float().cpu().reshape(n, h, w, c) > 0).permute(0, 3, 1, 2)
    if name == "dec4":
        m = m[:, :, :, 1:w - 3]
    return m.float()


def _nchw(t, shape):
    n, h, w, c = shape
    return t.detach().float().cpu().reshape(n, h, w, c).permute(0, 3, 1, 2)


def _pool_frozen(x, x_eng, k, s, p):
    """max pool whose argmax comes from x_eng (the engine's fp16 values) but whose values / gradient use x"""
    _, idx = F.max_pool2d(x_eng, k, s, p, return_indices=True)
    return x.flatten(2).gather(2, idx.flatten(2)).view_as(idx)


def _bn(x, sd, p):
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=1e-5)


def forward(eng, sd, x):
    def relu(t, name):
        return t * _mask(eng, name)

    y0 = relu(_bn(F.conv2d(x, sd["resnet.conv1.weight"], None, 2, 3), sd, "resnet.bn1"), "stem")
    cur = _pool_frozen(y0, _nchw(*eng.relu_outs["stem"]), 3, 2, 1)
    encs = []
    for li, blocks in enumerate(BLOCKS, start=1):
        for b in range(blocks):
            p = "resnet.layer%d.%d" % (li, b)
            stride = 2 if (b == 0 and li > 1) else 1
            o = relu(_bn(F.conv2d(cur, sd[p + ".conv1.weight"]), sd, p + ".bn1"), p + ".relu1")
            o = relu(_bn(F.conv2d(o, sd[p + ".conv2.weight"], None, stride, 1), sd, p + ".bn2"), p + ".relu2")
            o = _bn(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            idt = cur
            if (p + ".downsample.0.weight") in sd:
                idt = _bn(F.conv2d(cur, sd[p + ".downsample.0.weight"], None, stride), sd, p + ".downsample.1")
            cur = relu(o + idt, p + ".out")
        encs.append(cur)
    e1, e2, e3, e4 = encs

    def dec(name, t):
        return relu(F.conv2d(F.interpolate(t, scale_factor=2, mode="nearest"), sd[name + ".block.block.weight"], None, 1, 1), name)

    last = "resnet.layer4.2.out"
    c = dec("center", _pool_frozen(e4, _nchw(*eng.relu_outs[last]), 2, 2, 0))
    d0 = dec("dec0", torch.cat([e4, c], 1))
    d1 = dec("dec1", torch.cat([e3, d0], 1))
    d2 = dec("dec2", torch.cat([e2, d1], 1))
    d3 = dec("dec3", torch.cat([e1, d2], 1))
    d4 = dec("dec4", d3)
    d5 = relu(F.conv2d(d4, sd["dec5.block.weight"], None, 1, 1), "dec5")
    return F.conv2d(d5, sd["final.weight"], sd["final.bias"])
