"""Loss modules with the reference's class names and call signature (robosat/losses.py:8-119):
`forward(inputs fp32 [N, C, H, W], targets int64 [N, H, W]) -> 0-dim tensor` supporting `.backward()` / `.item()`.

All four losses selectable in `rs train` (train.py:97-102) run entirely in librsb200.so: the forward kernel pipeline also
produces the closed-form gradient, which `backward` just scales by the incoming gradient.
"""

import torch
import torch.nn as nn

from robosat_b200 import _lib


class _LovaszFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets):
        if not inputs.is_cuda:
            raise _lib.RsbError("LovaszLoss2d runs on the GPU kernels only (no CPU fallback)")
        lib = _lib.load()
        n, c, h, w = inputs.shape
        x = inputs.detach().contiguous().float()
        t = targets.contiguous().long()
        nbytes = lib.rsb_lovasz_workspace_bytes(n, c, h * w)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        _lib.check(lib.rsb_lovasz(x.data_ptr(), t.data_ptr(), loss.data_ptr(), grad.data_ptr(), ws.data_ptr(), nbytes, n, c, h * w,
                                  _lib.current_stream_ptr()), "rsb_lovasz")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, weight):
        if not inputs.is_cuda:
            raise _lib.RsbError("CrossEntropyLoss2d runs on the GPU kernels only (no CPU fallback)")
        lib = _lib.load()
        n, c, h, w = inputs.shape
        x = inputs.detach().contiguous().float()
        t = targets.contiguous().long()
        wt = weight.to(x.device).contiguous().float() if weight is not None else None
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        scratch = torch.empty(2, dtype=torch.float64, device=x.device)
        _lib.check(lib.rsb_cross_entropy(x.data_ptr(), t.data_ptr(), wt.data_ptr() if wt is not None else None, loss.data_ptr(), grad.data_ptr(),
                                         scratch.data_ptr(), n, c, h * w, _lib.current_stream_ptr()), "rsb_cross_entropy")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


class CrossEntropyLoss2d(nn.Module):
    """Class-weighted cross entropy (losses.py:8-25)."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", weight if weight is None else torch.as_tensor(weight, dtype=torch.float32))

    def forward(self, inputs, targets):
        return _CrossEntropyFn.apply(inputs, targets, self.weight)


class LovaszLoss2d(nn.Module):
    """Lovasz hinge over the flattened C*H*W vector of every image (losses.py:86-119)."""

    def __init__(self):
        super().__init__()

    def forward(self, inputs, targets):
        return _LovaszFn.apply(inputs, targets)


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, weight, gamma):
        if not inputs.is_cuda:
            raise _lib.RsbError("FocalLoss2d runs on the GPU kernels only (no CPU fallback)")
        lib = _lib.load()
        n, c, h, w = inputs.shape
        x = inputs.detach().contiguous().float()
        t = targets.contiguous().long()
        wt = weight.to(x.device).contiguous().float() if weight is not None else None
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        scratch = torch.empty(2, dtype=torch.float64, device=x.device)
        _lib.check(lib.rsb_focal(x.data_ptr(), t.data_ptr(), wt.data_ptr() if wt is not None else None, float(gamma), loss.data_ptr(), grad.data_ptr(),
                                 scratch.data_ptr(), n, c, h * w, _lib.current_stream_ptr()), "rsb_focal")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None, None


class _MIoUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, weight):
        if not inputs.is_cuda:
            raise _lib.RsbError("mIoULoss2d runs on the GPU kernels only (no CPU fallback)")
        lib = _lib.load()
        n, c, h, w = inputs.shape
        x = inputs.detach().contiguous().float()
        t = targets.contiguous().long()
        wt = weight.to(x.device).contiguous().float() if weight is not None else None
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        scratch = torch.empty(lib.rsb_miou_scratch_doubles(n, c), dtype=torch.float64, device=x.device)
        _lib.check(lib.rsb_miou(x.data_ptr(), t.data_ptr(), wt.data_ptr() if wt is not None else None, loss.data_ptr(), grad.data_ptr(),
                                scratch.data_ptr(), n, c, h * w, _lib.current_stream_ptr()), "rsb_miou")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


class FocalLoss2d(nn.Module):
    """Focal loss (losses.py:28-50): class-weighted NLL of (1 - softmax)^gamma * log_softmax."""

    def __init__(self, gamma=2, weight=None):
        super().__init__()
        self.gamma = gamma
        self.register_buffer("weight", weight if weight is None else torch.as_tensor(weight, dtype=torch.float32))

    def forward(self, inputs, targets):
        return _FocalFn.apply(inputs, targets, self.weight, self.gamma)


class mIoULoss2d(nn.Module):
    """max(soft mIoU loss, class-weighted cross entropy) (losses.py:53-83); the larger term also provides the gradient."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", weight if weight is None else torch.as_tensor(weight, dtype=torch.float32))

    def forward(self, inputs, targets):
        return _MIoUFn.apply(inputs, targets, self.weight)
