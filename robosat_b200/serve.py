"""Single-image latency path: the B200 replacement for `Predictor.segment` of `rs serve`
(robosat/tools/serve.py:135-192).

Reference: PIL image -> ConvertImageMode("RGB") -> ImageToTensor -> Normalize -> net(batch of 1) -> .cpu().numpy()
-> argmax(axis=0).astype(uint8) -> P-mode image with the dataset's palette.

Here the whole device side of that call -- H2D copy of the raw uint8 pixels from a pinned buffer, normalisation, the U-Net plan
(60 launches), the per-pixel argmax and the D2H copy of one byte per pixel -- is captured ONCE into a CUDA graph and replayed per
request. Measured on B200 (`bench.py` `serve` leg, scripts/gpu_serve_diag.py): graph replay and launching the same kernels one by
one on the stream are within 5 % of each other (0.97 vs 1.00 ms per 512x512 request in strict precision, 0.65 vs 0.70 ms in fast):
the kernels already chain through programmatic dependent launch, so the stream version is not launch-bound and the graph buys
little GPU time -- what it does buy is one driver call instead of 63 ctypes launches (0.17 ms of single-threaded host work per
request) and immunity to host jitter. The latency floor is ~10 us of prologue / pipeline fill per layer. Weights, activations
and both pinned staging buffers are static. `use_graph=False` serves from the stream with identical results (tests/test_serve_gpu.py).
"""

import numpy as np
import torch
from PIL import Image

from robosat_b200 import _lib
from robosat_b200.colors import make_palette
from robosat_b200.engine import UNetEngine


class SegmentEngine:
    """uint8 RGB tiles [B, H, W, 3] (host) -> uint8 class-index masks [B, H, W] (host), graph-replayed."""

    def __init__(self, state_dict, num_classes, height, width, batch=1, device="cuda", use_graph=True, precision=None):
        assert num_classes <= 255
        self.device = torch.device(device)
        self.batch, self.H, self.W, self.classes = batch, height, width, num_classes
        self.engine = UNetEngine(state_dict, num_classes, batch, height, width, device=self.device, precision=precision)
        self.h_in = torch.empty((batch, height, width, 3), dtype=torch.uint8, pin_memory=True)
        self.d_in = torch.empty((batch, height, width, 3), dtype=torch.uint8, device=self.device)
        self.d_mask = torch.empty((batch, height, width), dtype=torch.uint8, device=self.device)
        self.h_mask = torch.empty((batch, height, width), dtype=torch.uint8, pin_memory=True)
        self.graph = None
        self.graph_error = None
        if use_graph:
            self._capture()

    def _enqueue(self):
        """H2D, forward, argmax, D2H on the current stream (also what the graph records)."""
        self.d_in.copy_(self.h_in, non_blocking=True)
        logits = self.engine.forward(self.d_in)
        _lib.check(_lib.load().rsb_head_argmax(logits.data_ptr(), self.d_mask.data_ptr(), self.batch, self.classes, self.H * self.W,
                                               _lib.current_stream_ptr()), "rsb_head_argmax")
        self.h_mask.copy_(self.d_mask, non_blocking=True)

    def _capture(self):
        # warm up on a side stream first: one-time cudaFuncSetAttribute calls and lazy module loading must not happen
        # while the stream is capturing
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._enqueue()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph):
                self._enqueue()
        except Exception as exc:  # keep serving from the stream (same kernels, launch-bound) and say why
            self.graph_error = "%s: %s" % (type(exc).__name__, exc)
            torch.cuda.synchronize(self.device)
            return
        self.graph = graph

    def run(self):
        """Process what is in `h_in`; returns `h_mask` (pinned, valid until the next call)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_mask

    def segment_u8(self, tiles):
        """tiles: uint8 [H, W, 3] or [B, H, W, 3] (numpy or torch, host) -> numpy uint8 [H, W] / [B, H, W] class indices."""
        t = tiles.numpy() if torch.is_tensor(tiles) else np.asarray(tiles)
        single = t.ndim == 3
        if single:
            t = t[None]
        assert t.shape == (self.batch, self.H, self.W, 3) and t.dtype == np.uint8, "expected uint8 [B, H, W, 3]"
        self.h_in.numpy()[...] = t  # plain memcpy into the pinned staging buffer
        out = self.run().numpy().copy()
        return out[0] if single else out


class Predictor:
    """Same constructor and `segment(image) -> PIL.Image` contract as robosat/tools/serve.py:135-172."""

    def __init__(self, checkpoint, model, dataset):
        cuda = model["common"]["cuda"]
        assert torch.cuda.is_available() or not cuda, "cuda is available when requested"
        if not cuda:
            raise _lib.RsbError("robosat_b200 serves from a B200 only: set common.cuda = true (there is no CPU path)")
        self.cuda = cuda
        self.device = torch.device("cuda")
        self.checkpoint = checkpoint
        self.model = model
        self.dataset = dataset
        self.num_classes = len(dataset["common"]["classes"])
        chkpt = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, str) else checkpoint
        self.state_dict = chkpt["state_dict"]
        self.palette = make_palette(*dataset["common"]["colors"])
        self._engines = {}  # one captured graph per image extent

    def _engine_for(self, height, width):
        eng = self._engines.get((height, width))
        if eng is None:
            eng = self._engines[(height, width)] = SegmentEngine(self.state_dict, self.num_classes, height, width, device=self.device)
        return eng

    def segment(self, image):
        rgb = np.asarray(image.convert("RGB"), dtype=np.uint8)  # ConvertImageMode("RGB"), serve.py:154
        mask = self._engine_for(rgb.shape[0], rgb.shape[1]).segment_u8(rgb)
        out = Image.fromarray(mask, mode="P")
        out.putpalette(self.palette)
        return out
