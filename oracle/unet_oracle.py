"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference U-Net forward.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product path (`robosat_b200/`) never does.

What it restates
    `UNet.forward`                      /root/reference/robosat/unet.py:110-141
    `ConvRelu.forward`                  /root/reference/robosat/unet.py:44   (3x3, pad 1, no bias, ReLU)
    `DecoderBlock.forward`              /root/reference/robosat/unet.py:73   (nearest x2 upsample -> ConvRelu)
    torchvision `resnet50` (v1.5)       torchvision/models/resnet.py `Bottleneck.forward`, `_make_layer`
                                        (third-party; the reference pins torchvision~=0.3, setup.py:38)

The arithmetic itself (conv2d, batch_norm, max_pool2d, interpolate) is delegated to torch's
CPU fp32 kernels exactly as the reference does (the reference contains no arithmetic of its
own); what is restated here is the network topology, written against a plain
`state_dict` so that it needs neither `/root/reference` nor torchvision's module classes.

Pinning: `tests/golden/make_golden.py` imports the real reference (`/root/reference`, CPU) in
the build container, loads the same seeded `state_dict`, and stores its logits as fixtures;
`tests/test_oracle.py` checks this restatement against them bit-for-bit.
"""

import torch
import torch.nn.functional as F

RESNET50_BLOCKS = (3, 4, 6, 3)
BN_EPS = 1e-5  # nn.BatchNorm2d default used by torchvision resnet


def _strip(sd):
    """Accept both `module.`-prefixed (DataParallel, train.py:69) and bare keys."""
    if any(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return dict(sd)


_TRAIN = [False]


def _bn(x, sd, p):
    # eval-mode BatchNorm2d: (x - running_mean) / sqrt(running_var + eps) * weight + bias
    # train mode (net.train(), train.py:168): batch statistics, running stats updated in place with momentum 0.1
    if _TRAIN[0]:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            training=True, momentum=0.1, eps=BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def _bottleneck(x, sd, p, stride, conv):
    # torchvision Bottleneck.forward: 1x1 -> bn -> relu -> 3x3(stride) -> bn -> relu -> 1x1 -> bn -> (+identity) -> relu
    identity = x
    out = F.relu(_bn(conv(x, sd[p + ".conv1.weight"], None, 1, 0), sd, p + ".bn1"))
    out = F.relu(_bn(conv(out, sd[p + ".conv2.weight"], None, stride, 1), sd, p + ".bn2"))
    out = _bn(conv(out, sd[p + ".conv3.weight"], None, 1, 0), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        identity = _bn(conv(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    return F.relu(out + identity)


def _conv(x, w, b, stride, padding):
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def unet_forward_train(state_dict, x, return_features=False):
    """train-mode forward (batch-statistics BatchNorm; `running_mean` / `running_var` of `state_dict` are updated in place,
    like net.train() + net(images) at train.py:168,180). Tensors with requires_grad=True give autograd gradients."""
    _TRAIN[0] = True
    try:
        return unet_forward(state_dict, x, return_features=return_features)
    finally:
        _TRAIN[0] = False


def unet_forward(state_dict, x, conv=_conv, return_features=False):
    """fp32 logits [N, C, H, W] for fp32 NCHW input x, eval mode (unet.py:110-141).

    `conv` is injectable so a study script can emulate reduced-precision operands; the
    default is plain fp32 `F.conv2d`.
    """

    sd = _strip(state_dict)
    assert x.size(-1) % 32 == 0 and x.size(-2) % 32 == 0, "image resolution has to be divisible by 32 for resnet"

    feats = {}
    # unet.py:122-125
    enc0 = conv(x, sd["resnet.conv1.weight"], None, 2, 3)
    enc0 = F.relu(_bn(enc0, sd, "resnet.bn1"))
    feats["stem"] = enc0
    enc0 = F.max_pool2d(enc0, kernel_size=3, stride=2, padding=1)
    feats["enc0"] = enc0

    # unet.py:127-130
    cur = enc0
    encs = []
    for li, blocks in enumerate(RESNET50_BLOCKS, start=1):
        for b in range(blocks):
            stride = 2 if (b == 0 and li > 1) else 1
            cur = _bottleneck(cur, sd, "resnet.layer{}.{}".format(li, b), stride, conv)
        encs.append(cur)
        feats["enc{}".format(li)] = cur
    enc1, enc2, enc3, enc4 = encs

    def dec(name, t):
        # DecoderBlock (unet.py:73) -> ConvRelu (unet.py:44)
        up = F.interpolate(t, scale_factor=2, mode="nearest")
        return F.relu(conv(up, sd[name + ".block.block.weight"], None, 1, 1))

    center = dec("center", F.max_pool2d(enc4, kernel_size=2, stride=2))  # unet.py:132
    dec0 = dec("dec0", torch.cat([enc4, center], dim=1))  # unet.py:134
    dec1 = dec("dec1", torch.cat([enc3, dec0], dim=1))  # unet.py:135
    dec2 = dec("dec2", torch.cat([enc2, dec1], dim=1))  # unet.py:136
    dec3 = dec("dec3", torch.cat([enc1, dec2], dim=1))  # unet.py:137
    dec4 = dec("dec4", dec3)  # unet.py:138
    dec5 = F.relu(conv(dec4, sd["dec5.block.weight"], None, 1, 1))  # unet.py:139
    logits = conv(dec5, sd["final.weight"], sd["final.bias"], 1, 0)  # unet.py:141

    if return_features:
        feats.update(center=center, dec0=dec0, dec1=dec1, dec2=dec2, dec3=dec3, dec4=dec4, dec5=dec5)
        return logits, feats
    return logits


def predict_probs(state_dict, x):
    """softmax over classes as in the predict loop (predict.py:84-87)."""
    with torch.no_grad():
        return F.softmax(unet_forward(state_dict, x), dim=1)
