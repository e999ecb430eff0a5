"""Segmentation metrics with the reference's accumulation semantics (robosat/metrics.py:13-84).

`Metrics.add(actual, predicted)` keeps the reference's per-sample signature. On CUDA tensors the counting runs in
one kernel (`rsb_metrics_count`) into a device-side int64[4] that is read back only when a score is requested,
instead of four `.item()` synchronisations per sample. `add_batch` takes a whole batch at once.
"""

import math

import numpy as np
import torch

from robosat_b200 import _lib


class Metrics:
    def __init__(self, labels):
        self.labels = labels
        self._host = [0, 0, 0, 0]  # tn, fn, fp, tp (the reference's naming, metrics.py:36-41)
        self._dev = None

    # ---------------------------------------------------------------- accumulation
    def add(self, actual, predicted):
        """actual: [H, W] labels, predicted: [C, H, W] scores for one sample"""
        self.add_batch(actual.unsqueeze(0), predicted.unsqueeze(0))

    def add_batch(self, actual, predicted):
        """actual: int64 [N, H, W], predicted: fp32 [N, C, H, W]"""
        if not predicted.is_cuda:
            raise _lib.RsbError("Metrics runs on the GPU kernels only; move the tensors to a CUDA device")
        n, c, h, w = predicted.shape
        if self._dev is None:
            self._dev = torch.zeros(4, dtype=torch.int64, device=predicted.device)
        lib = _lib.load()
        _lib.check(lib.rsb_metrics_count(predicted.contiguous().float().data_ptr(), actual.contiguous().long().data_ptr(), self._dev.data_ptr(),
                                         n, c, h * w, _lib.current_stream_ptr()), "rsb_metrics_count")

    def _counts(self):
        if self._dev is not None:
            for i, v in enumerate(self._dev.cpu().tolist()):
                self._host[i] += v
            self._dev.zero_()
        return self._host

    tn = property(lambda self: self._counts()[0])
    fn = property(lambda self: self._counts()[1])
    fp = property(lambda self: self._counts()[2])
    tp = property(lambda self: self._counts()[3])

    # ---------------------------------------------------------------- scores (metrics.py:43-84)
    def get_miou(self):
        tn, fn, fp, tp = self._counts()
        try:
            return np.nanmean([tn / (tn + fn + fp), tp / (tp + fn + fp)])
        except ZeroDivisionError:
            return float("NaN")

    def get_fg_iou(self):
        tn, fn, fp, tp = self._counts()
        try:
            return tp / (tp + fn + fp)
        except ZeroDivisionError:
            return float("NaN")

    def get_mcc(self):
        tn, fn, fp, tp = self._counts()
        try:
            return (tp * tn - fp * fn) / math.sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))
        except ZeroDivisionError:
            return float("NaN")
