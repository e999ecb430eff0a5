"""Builders for single-convolution parity cases shared by the GPU tests and tests/gpu_diag.py.

Each case builds an `rsb_conv_desc` exactly the way `robosat_b200.engine.UNetEngine` does for that layer
kind, plus a CPU fp32 reference of the same operation (torch functional ops on the fp16-rounded operands,
i.e. the reference's own call sites: F.conv2d / F.interpolate / torch.cat, robosat/unet.py:44,73,134-141).
"""

import torch
import torch.nn.functional as F

from robosat_b200 import engine as E
from robosat_b200._lib import ConvSrc


def _rand(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen) * scale)


def _nhwc_half(x_nchw, device):
    return x_nchw.permute(0, 2, 3, 1).contiguous().half().to(device)


def _out_nchw(t, shape_nhwc):
    t = t.cpu()
    if t.dim() == len(shape_nhwc) + 1:  # split precision: (hi, lo) planes
        t = (t[0].double() + t[1].double())
    return t.float().reshape(shape_nhwc).permute(0, 3, 1, 2).contiguous()


def _pair_nhwc(x_nchw, device):
    """fp32 NCHW -> [2, N, H, W, C] fp16 (hi, lo) planes"""
    v = x_nchw.permute(0, 2, 3, 1).contiguous().float()
    hi = v.half()
    return torch.stack([hi, (v - hi.float()).half()]).to(device)


def _pair_value(x_nchw):
    """what the (hi, lo) pair of an fp32 tensor represents (float64)"""
    hi = x_nchw.float().half()
    lo = (x_nchw.float() - hi.float()).half()
    return hi.double() + lo.double()


def _wsplit(wp64, device):
    hi, lo, scale = E.split_fp16(wp64.double())
    return torch.stack([hi, lo]).to(device), scale, (hi.double() + lo.double()) * scale


def _act_in(x, device, split):
    """activation operand + the value it represents"""
    if split:
        return _pair_nhwc(x, device), _pair_value(x)
    return _nhwc_half(x, device), x.half().float()


def _plane(t, split):
    return t.numel() // 2 if split else 0


def _scratch(device, split):
    """fp32 scratch for the K-chunked accumulation of long strict-precision K loops (the plan decides whether it is used)"""
    return torch.zeros(148 * 128 * 256, dtype=torch.float32, device=device) if split else None


class Case:
    def __init__(self, name, desc, keep, out, out_view, ref):
        self.name, self.desc, self.keep, self.out, self.out_view, self.ref = name, desc, keep, out, out_view, ref

    def result(self):
        """device output as fp32 NCHW on the CPU"""
        return self.out_view(self.out)


def conv_case(kind, N, H, W, cin, cout, device, seed=0, residual=False, relu=True, bias=True, block_n=None, split=False, cta_pair=None):
    """kind: '1x1' | '3x3' | '3x3s2' | '1x1s2'. split=True: strict precision (hi/lo planes, fp64 reference)"""
    g = torch.Generator().manual_seed(seed)
    k = 3 if kind.startswith("3x3") else 1
    stride = 2 if kind.endswith("s2") else 1
    x = _rand((N, cin, H, W), g)
    w = _rand((cout, cin, k, k), g, (2.0 / (cin * k * k)) ** 0.5)
    b = _rand((cout,), g, 0.1) if bias else None
    oH, oW = H // stride, W // stride
    xd, x = _act_in(x, device, split)
    scale = 1.0
    if split:
        wp, scale, wv = _wsplit(E.pack_conv(w.double()), device)
        w = wv.reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous()
    else:
        wp = E.pack_conv(w).half().to(device)
        w = w.half().float()
    bd = b.float().to(device) if bias else None
    out = torch.zeros(*((2,) if split else ()), N, oH, oW, cout, dtype=torch.float16, device=device)
    res = _rand((N, cout, oH, oW), g) if residual else None
    resd = None
    if residual:
        resd, res = _act_in(res, device, split)
    pl = _plane(xd, split)
    scr = _scratch(device, split)
    if stride == 1:
        srcs = [E._src_dense(xd, N, H, W, cin, pl)]
        segs = [(0, kh - k // 2, kw - k // 2, cin // 64) for kh in range(k) for kw in range(k)]
    else:
        srcs = [E._src_parity(xd, N, H, W, cin, ph, pw, pl) for ph in range(2) for pw in range(2)]
        segs = []
        for kh in range(k):
            for kw in range(k):
                ph, dh = (kh - k // 2) % 2, (kh - k // 2) // 2
                pw, dw = (kw - k // 2) % 2, (kw - k // 2) // 2
                segs.append((ph * 2 + pw, dh, dw, cin // 64))
        if k == 1:
            srcs = srcs[:1]
    desc = E.make_conv_desc(srcs, segs, wp, bd, cout, 1, (oW, oH, N), out, (cout, oW * cout, oH * oW * cout),
                            residual=resd, relu=relu, block_n=block_n, split=split, acc_scale=scale, out_plane=_plane(out, split),
                            res_plane=_plane(resd, split) if residual else 0, cta_pair=cta_pair, scratch=scr)

    def ref():
        y = F.conv2d(x, w, b.to(x.dtype) if bias else None, stride=stride, padding=k // 2)
        if residual:
            y = y + res
        return (F.relu(y) if relu else y).float()

    return Case("%s%s_%dx%dx%d_%d-%d" % ("split_" if split else "", kind, N, H, W, cin, cout), desc, (xd, wp, bd, resd, scr), out,
                lambda t: _out_nchw(t, (N, oH, oW, cout)), ref)


def decoder_case(N, lh, lw, cins, cout, device, seed=0, block_n=None, split=False, cta_pair=None):
    """DecoderBlock on cat(sources): nearest x2 -> 3x3 conv -> relu, as 4 phases on the low-res inputs."""
    g = torch.Generator().manual_seed(seed)
    xs = [_rand((N, c, lh, lw), g) for c in cins]
    ctot = sum(cins)
    w = _rand((cout, ctot, 3, 3), g, (2.0 / (ctot * 9)) ** 0.5)
    pairs = [_act_in(x, device, split) for x in xs]
    xds, xs = [p[0] for p in pairs], [p[1] for p in pairs]
    scale = 1.0
    if split:
        wp, scale, _ = _wsplit(E.pack_upsample_phases(w.double()), device)
        w = w.double()
    else:
        wp = E.pack_upsample_phases(w).half().to(device)
    oH, oW = 2 * lh, 2 * lw
    out = torch.zeros(*((2,) if split else ()), N, oH, oW, cout, dtype=torch.float16, device=device)
    srcs = [E._src_dense(t, N, lh, lw, c, _plane(t, split)) for t, c in zip(xds, cins)]
    segs = [(si, th - 1, tw - 1, c // 64) for th in range(2) for tw in range(2) for si, c in enumerate(cins)]
    scr = _scratch(device, split)
    desc = E.make_conv_desc(srcs, segs, wp, None, cout, 4, (lw, lh, N), out, (cout, oW * cout, oH * oW * cout),
                            out_scale=(2, 2), block_n=block_n, split=split, acc_scale=scale, out_plane=_plane(out, split), cta_pair=cta_pair,
                            scratch=scr)

    def ref():
        up = F.interpolate(torch.cat(xs, dim=1), scale_factor=2, mode="nearest")
        return F.relu(F.conv2d(up, w, None, padding=1)).float()

    return Case("%sdecoder_%dx%dx%d_%s-%d" % ("split_" if split else "", N, lh, lw, "+".join(map(str, cins)), cout), desc, (xds, wp, scr), out,
                lambda t: _out_nchw(t, (N, oH, oW, cout)), ref)


def stem_case(N, H, W, device, seed=0, split=False):
    """resnet conv1 7x7/2 pad 3 + folded bn + relu through the pre-pass and the overlapped-window view."""
    from robosat_b200 import _lib

    g = torch.Generator().manual_seed(seed)
    x = _rand((N, 3, H, W), g)
    w = _rand((64, 3, 7, 7), g, (2.0 / 147) ** 0.5)
    b = _rand((64,), g, 0.1)
    H2, W2, Wp = H // 2, W // 2, W // 2 + 4
    xd = x.contiguous().to(device)
    if torch.device(device).type == "cpu":
        import emulate

        s2d = emulate.prepass_s2d_split_cpu(x) if split else emulate.prepass_s2d_cpu(x)
    else:
        s2d = torch.zeros(*((2,) if split else ()), N, H2, Wp, 16, dtype=torch.float16, device=device)
        lib = _lib.load()
        if split:
            _lib.check(lib.rsb_prepass_s2d_split(xd.data_ptr(), 0, s2d.data_ptr(), s2d.numel() // 2, N, H, W, None, None, _lib.current_stream_ptr()), "prepass")
        else:
            _lib.check(lib.rsb_prepass_s2d(xd.data_ptr(), 0, s2d.data_ptr(), N, H, W, None, None, _lib.current_stream_ptr()), "prepass")
    scale = 1.0
    if split:
        wp, scale, _ = _wsplit(E.pack_stem(w.double()), device)
    else:
        wp = E.pack_stem(w).half().to(device)
    bd = b.to(device)
    out = torch.zeros(*((2,) if split else ()), N, H2, W2, 64, dtype=torch.float16, device=device)
    src = ConvSrc(s2d.data_ptr(), 16, Wp * 16, H2 * Wp * 16, 64, W2, H2, N, _plane(s2d, split))
    segs = [(0, t - 2, 0, 1) for t in range(4)]
    desc = E.make_conv_desc([src], segs, wp, bd, 64, 1, (W2, H2, N), out, (64, W2 * 64, H2 * W2 * 64), split=split, acc_scale=scale,
                            out_plane=_plane(out, split))

    def ref():
        if split:
            return F.relu(F.conv2d(_pair_value(x), w.double(), b.double(), stride=2, padding=3)).float()
        return F.relu(F.conv2d(x.half().float(), w.half().float(), b, stride=2, padding=3))

    return Case("%sstem_%dx%dx%d" % ("split_" if split else "", N, H, W), desc, (xd, s2d, wp, bd), out, lambda t: _out_nchw(t, (N, H2, W2, 64)), ref)


def head_case(N, H, W, classes, device, seed=0, split=False):
    """dec5 (3x3 32->32 + relu) fused with final (1x1 32->classes + bias) reading the W-padded dec4 buffer."""
    g = torch.Generator().manual_seed(seed)
    x = _rand((N, 32, H, W), g)
    w5 = _rand((32, 32, 3, 3), g, (2.0 / 288) ** 0.5)
    wf = _rand((classes, 32, 1, 1), g, 0.3)
    bf = _rand((classes,), g, 0.1)
    Wq = W + 4
    xd, x = _act_in(x, device, split)
    buf = torch.zeros(*((2,) if split else ()), N, H, Wq, 32, dtype=torch.float16, device=device)
    buf[..., 1:W + 1, :] = xd
    scale = 1.0
    if split:
        wp, scale, _ = _wsplit(E.pack_window3(w5.double()), device)
        w5 = w5.double()
    else:
        wp = E.pack_window3(w5).half().to(device)
        w5 = w5.half().float()
    hw = wf.reshape(classes, 32).contiguous().to(device)
    hb = bf.to(device)
    logits = torch.zeros(N, classes, H, W, dtype=torch.float32, device=device)
    src = ConvSrc(buf.data_ptr(), 32, Wq * 32, H * Wq * 32, 128, W, H, N, _plane(buf, split))
    segs = [(0, kh - 1, 0, 2) for kh in range(3)]
    desc = E.make_conv_desc([src], segs, wp, None, 32, 1, (W, H, N), None, None, head=(hw, hb, logits, classes), split=split, acc_scale=scale)

    def ref():
        return F.conv2d(F.relu(F.conv2d(x, w5, None, padding=1)), wf.to(x.dtype), bf.to(x.dtype)).float()

    return Case("%shead_%dx%dx%d_c%d" % ("split_" if split else "", N, H, W, classes), desc, (buf, wp, hw, hb), logits, lambda t: t.float().cpu(), ref)


def default_cases(device):
    """One case per layer kind the U-Net uses, small enough for the CPU reference to take milliseconds."""
    return [
        lambda: conv_case("1x1", 2, 16, 16, 64, 64, device, seed=1),
        lambda: conv_case("1x1", 1, 16, 32, 256, 128, device, seed=2, residual=True, block_n=128),
        lambda: conv_case("3x3", 2, 16, 16, 64, 64, device, seed=3),
        lambda: conv_case("3x3", 1, 24, 40, 128, 256, device, seed=4, block_n=256),
        lambda: conv_case("3x3s2", 2, 32, 32, 128, 128, device, seed=5),
        lambda: conv_case("1x1s2", 2, 32, 32, 256, 512, device, seed=6, relu=False),
        lambda: conv_case("1x1", 3, 8, 8, 512, 2048, device, seed=7, residual=True, block_n=256),
        lambda: decoder_case(2, 8, 8, [128, 64], 64, device, seed=8),
        lambda: decoder_case(1, 16, 16, [128], 32, device, seed=9),
        lambda: decoder_case(3, 4, 4, [256, 256], 256, device, seed=10, block_n=128),
        lambda: stem_case(2, 64, 64, device, seed=11),
        lambda: head_case(2, 32, 32, 2, device, seed=12),
        lambda: head_case(1, 32, 64, 6, device, seed=13),
    ]


def split_cases(device):
    """The same layer kinds in strict precision (hi/lo operand planes, three MMAs per K step); references in float64."""
    return [
        lambda: conv_case("1x1", 2, 16, 16, 64, 64, device, seed=1, split=True),
        lambda: conv_case("1x1", 1, 16, 32, 256, 128, device, seed=2, residual=True, block_n=128, split=True),
        lambda: conv_case("1x1", 2, 16, 16, 64, 256, device, seed=14, residual=True, block_n=64, split=True),
        lambda: conv_case("3x3", 2, 16, 16, 64, 64, device, seed=3, split=True),
        lambda: conv_case("3x3", 1, 24, 40, 128, 256, device, seed=4, block_n=256, split=True, cta_pair=False),
        lambda: conv_case("3x3", 1, 24, 40, 128, 256, device, seed=4, block_n=256, split=True, cta_pair=True),
        lambda: conv_case("3x3", 8, 64, 64, 64, 256, device, seed=21, block_n=128, split=True, cta_pair=True),   # multi-wave
        lambda: conv_case("3x3s2", 2, 32, 32, 128, 128, device, seed=5, split=True),
        lambda: conv_case("1x1s2", 2, 32, 32, 256, 512, device, seed=6, relu=False, split=True),
        lambda: conv_case("1x1", 3, 8, 8, 512, 2048, device, seed=7, residual=True, block_n=256, split=True, cta_pair=True),
        lambda: conv_case("1x1", 3, 8, 8, 512, 2048, device, seed=7, residual=True, block_n=128, split=True, cta_pair=False),
        lambda: decoder_case(2, 8, 8, [128, 64], 64, device, seed=8, split=True),
        lambda: decoder_case(1, 16, 16, [128], 32, device, seed=9, split=True),
        lambda: decoder_case(3, 4, 4, [256, 256], 256, device, seed=10, block_n=128, split=True),
        lambda: decoder_case(2, 4, 4, [2048, 256], 256, device, seed=15, split=True),     # dec0's K = 9216 per phase: 9 K chunks
        lambda: decoder_case(2, 4, 4, [2048, 256], 256, device, seed=15, split=True, cta_pair=False, block_n=128),  # chunked, one CTA per tile
        lambda: conv_case("3x3", 2, 16, 16, 256, 256, device, seed=16, split=True),       # layer3 conv2: 36 K blocks -> 3 chunks, several tiles per CTA
        lambda: stem_case(2, 64, 64, device, seed=11, split=True),
        lambda: head_case(2, 32, 32, 2, device, seed=12, split=True),
        lambda: head_case(1, 32, 64, 6, device, seed=13, split=True),
    ]


# --------------------------------------------------------------------------------------------------
# line-buffer kernel cases (csrc/rsb_conv_row.cu): same arithmetic, different operand staging
# --------------------------------------------------------------------------------------------------
def row_conv_case(N, H, W, cin, cout, device, seed=0, bias=True, relu=True, rows_per_unit=0):
    g = torch.Generator().manual_seed(seed)
    x = _rand((N, cin, H, W), g).half().float()
    w = _rand((cout, cin, 3, 3), g, (2.0 / (cin * 9)) ** 0.5)
    b = _rand((cout,), g, 0.1) if bias else None
    xd = _nhwc_half(x, device)
    wp = E.pack_conv(w).half().to(device)
    bd = b.to(device) if bias else None
    out = torch.zeros(N, H, W, cout, dtype=torch.float16, device=device)
    desc = E.make_rowconv_desc(E._src_dense(xd, N, H, W, cin), cin, wp, bd, cout, (W, H, N), out, (cout, W * cout, H * W * cout), relu=relu,
                               rows_per_unit=rows_per_unit)

    def ref():
        y = F.conv2d(x, w.half().float(), b, padding=1)
        return F.relu(y) if relu else y

    return Case("row3x3_%dx%dx%d_%d-%d" % (N, H, W, cin, cout), desc, (xd, wp, bd), out, lambda t: _out_nchw(t, (N, H, W, cout)), ref)


def row_up_case(N, lh, lw, cin, cout, device, seed=0, rows_per_unit=0):
    g = torch.Generator().manual_seed(seed)
    x = _rand((N, cin, lh, lw), g).half().float()
    w = _rand((cout, cin, 3, 3), g, (2.0 / (cin * 9)) ** 0.5)
    xd = _nhwc_half(x, device)
    wp = E.pack_upsample_phases(w).half().to(device)
    oH, oW = 2 * lh, 2 * lw
    out = torch.zeros(N, oH, oW, cout, dtype=torch.float16, device=device)
    desc = E.make_rowconv_desc(E._src_dense(xd, N, lh, lw, cin), cin, wp, None, cout, (lw, lh, N), out, (cout, oW * cout, oH * oW * cout), upsample=True,
                               rows_per_unit=rows_per_unit)

    def ref():
        return F.relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, padding=1))

    return Case("rowup_%dx%dx%d_%d-%d" % (N, lh, lw, cin, cout), desc, (xd, wp), out, lambda t: _out_nchw(t, (N, oH, oW, cout)), ref)


def row_head_case(N, H, W, classes, device, seed=0, rows_per_unit=0, split=False):
    g = torch.Generator().manual_seed(seed)
    x = _rand((N, 32, H, W), g)
    w5 = _rand((32, 32, 3, 3), g, (2.0 / 288) ** 0.5)
    wf = _rand((classes, 32, 1, 1), g, 0.3)
    bf = _rand((classes,), g, 0.1)
    xd, x = _act_in(x, device, split)
    scale = 1.0
    if split:
        wp, scale, _ = _wsplit(E.pack_conv(w5.double()), device)
        w5 = w5.double()
    else:
        wp = E.pack_conv(w5).half().to(device)
        w5 = w5.half().float()
    hw, hb = wf.reshape(classes, 32).contiguous().to(device), bf.to(device)
    logits = torch.zeros(N, classes, H, W, dtype=torch.float32, device=device)
    desc = E.make_rowconv_desc(E._src_dense(xd, N, H, W, 32, _plane(xd, split)), 32, wp, None, 32, (W, H, N), None, None, head=(hw, hb, logits, classes),
                               rows_per_unit=rows_per_unit, split=split, acc_scale=scale)

    def ref():
        return F.conv2d(F.relu(F.conv2d(x, w5, None, padding=1)), wf.to(x.dtype), bf.to(x.dtype)).float()

    return Case("%srowhead_%dx%dx%d_c%d" % ("split_" if split else "", N, H, W, classes), desc, (xd, wp, hw, hb), logits, lambda t: t.float().cpu(), ref)


def row_cases(device):
    return [
        lambda: row_conv_case(2, 40, 160, 64, 64, device, seed=31),                    # partial strip + unit boundary (32-row units)
        lambda: row_conv_case(1, 9, 128, 64, 64, device, seed=32, rows_per_unit=4),    # several short units: ring wrap-around
        lambda: row_conv_case(2, 20, 256, 32, 32, device, seed=33, bias=False),        # 64-byte pixel rows (SWIZZLE_64B)
        lambda: row_conv_case(1, 12, 128, 64, 32, device, seed=34, relu=False),        # narrow output, no ReLU
        lambda: row_up_case(2, 24, 160, 128, 32, device, seed=35, rows_per_unit=8),    # dec4: fused upsample, 2 sub-tiles, 2 row phases
        lambda: row_up_case(1, 8, 128, 64, 64, device, seed=36),
        lambda: row_head_case(2, 36, 192, 2, device, seed=37),                         # dec5 + final
        lambda: row_head_case(1, 16, 128, 6, device, seed=38, rows_per_unit=5),
    ]


def row_split_cases(device):
    """strict-precision line buffer (dec5 + final): hi/lo planes in the row ring, N-concatenated MMAs"""
    return [
        lambda: row_head_case(2, 36, 192, 2, device, seed=37, split=True),
        lambda: row_head_case(1, 16, 128, 6, device, seed=38, rows_per_unit=5, split=True),
        lambda: row_head_case(1, 40, 300, 2, device, seed=39, split=True),   # partial last strip
    ]
