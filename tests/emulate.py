"""CPU emulation of the semantics of `rsb_conv_run` for descriptors whose pointers are HOST memory.

Test infrastructure only. It reads the raw pointers of an `rsb_conv_desc` exactly the way the kernel's TMA
boxes do (views with arbitrary -- possibly overlapping -- pitches, zero fill outside the view extents) so the
host-side logic (weight packing, parity views, upsample phases, window views, output addressing) can be
verified without a GPU.
"""

import ctypes

import numpy as np


def _view(ptr, shape, strides_elems, dtype=np.float16, writable=False):
    item = np.dtype(dtype).itemsize
    span = 1 + sum((s - 1) * abs(st) for s, st in zip(shape, strides_elems))
    raw = (ctypes.c_uint8 * (span * item)).from_address(ptr)
    base = np.frombuffer(raw, dtype=dtype)
    v = np.lib.stride_tricks.as_strided(base, shape=shape, strides=[st * item for st in strides_elems], writeable=writable)
    return v


def prepass_s2d_cpu(x):
    """torch fp32 NCHW -> fp16 [N, H/2, W/2+4, 16] (what rsb_prepass_s2d writes)"""
    import torch

    N, C, H, W = x.shape
    out = torch.zeros(N, H // 2, W // 2 + 4, 16, dtype=torch.float16)
    for ph in range(2):
        for pw in range(2):
            out[:, :, 2:2 + W // 2, (ph * 2 + pw) * 3:(ph * 2 + pw) * 3 + 3] = x[:, :, ph::2, pw::2].permute(0, 2, 3, 1).half()
    return out


def run_desc(d):
    """Execute descriptor `d` (host pointers) and write its outputs like the device kernel would."""
    K = 64 * sum(d.segs[i].cblocks for i in range(d.nseg))
    wts = _view(d.weights, (d.phases * d.Cout, K), (K, 1)).astype(np.float32)
    bias = _view(d.bias, (d.Cout,), (1,), np.float32) if d.bias else None
    Nt, Ht, Wt = d.Nt, d.Ht, d.Wt
    hh0 = np.arange(Ht)[:, None]
    ww0 = np.arange(Wt)[None, :]
    for phase in range(d.phases):
        pa, pb = phase >> 1, phase & 1
        acc = np.zeros((Nt, Ht, Wt, d.Cout), dtype=np.float32)
        k0 = 0
        for si in range(d.nseg):
            seg = d.segs[si]
            src = d.srcs[seg.src]
            width = seg.cblocks * 64
            sv = _view(src.ptr, (src.N, src.H, src.W, min(width, src.C)), (src.pitch_n, src.pitch_h, src.pitch_w, 1))
            hh = hh0 + seg.dh + pa
            ww = ww0 + seg.dw + pb
            inb = (hh >= 0) & (hh < src.H) & (ww >= 0) & (ww < src.W)
            a = np.zeros((Nt, Ht, Wt, width), dtype=np.float32)
            g = sv[:, np.clip(hh, 0, src.H - 1), np.clip(ww, 0, src.W - 1), :].astype(np.float32)  # [N, Ht, Wt, c]
            g = g * inb[None, :, :, None]
            nn = min(Nt, src.N)
            a[:nn, :, :, :g.shape[-1]] = g[:nn]
            wseg = wts[phase * d.Cout:(phase + 1) * d.Cout, k0:k0 + width]
            acc += np.tensordot(a, wseg, axes=([3], [1]))
            k0 += width
        if bias is not None:
            acc += bias
        if d.mode == 0:
            base = d.out + 2 * (pa * d.out_pitch_h + pb * d.out_pitch_w)
            strides = (d.out_pitch_n, d.out_sy * d.out_pitch_h, d.out_sx * d.out_pitch_w, 1)
            if d.residual:
                rbase = d.residual + 2 * (pa * d.out_pitch_h + pb * d.out_pitch_w)
                acc += _view(rbase, (Nt, Ht, Wt, d.Cout), strides).astype(np.float32)
            if d.relu:
                acc = np.maximum(acc, 0)
            ov = _view(base, (Nt, Ht, Wt, d.Cout), strides, writable=True)
            ov[...] = acc.astype(np.float16)
        else:
            if d.relu:
                acc = np.maximum(acc, 0)
            hw = _view(d.head_w, (d.head_classes, 32), (32, 1), np.float32)
            hb = _view(d.head_b, (d.head_classes,), (1,), np.float32)
            logits = np.tensordot(acc, hw, axes=([3], [1])) + hb  # [N, H, W, classes]
            ov = _view(d.head_out, (Nt, d.head_classes, Ht, Wt), (d.head_classes * Ht * Wt, Ht * Wt, Wt, 1), np.float32, writable=True)
            ov[...] = np.transpose(logits, (0, 3, 1, 2))


def run_engine(engine, x):
    """Execute a plan_only UNetEngine (CPU buffers) op by op; returns fp32 NCHW logits (torch)."""
    import torch
    import torch.nn.functional as F

    assert engine.plan_only and engine.device.type == "cpu"
    for op in engine.ops:
        if op[0] == "prepass":
            engine.s2d.copy_(prepass_s2d_cpu(x))
        elif op[0] == "maxpool":
            _, src, dst, n, h, w, c, k, s, p = op
            y = F.max_pool2d(src.float().reshape(n, h, w, c).permute(0, 3, 1, 2), kernel_size=k, stride=s, padding=p)
            dst.copy_(y.permute(0, 2, 3, 1).half())
        else:
            run_desc(op[1].desc)
    return engine.logits.clone()
