"""Halo stitching on the device: the step immediately before the `rs predict` hot path.

Reference (`BufferedSlippyMapDirectory.__getitem__`, robosat/datasets.py:110-131 -> `buffer_tile_image`,
robosat/tiles.py:162-227): for EVERY tile, open the centre image and up to 8 neighbour images, crop `overlap`-wide strips
and paste them on a nodata=0 canvas -- i.e. every tile file is decoded up to 9 times, on one CPU thread per worker.

Here every tile is decoded ONCE (thread pool), uploaded once into a device-resident cache of raw RGB tiles
(`uint8 [slots][S][S][3]`, least-recently-used replacement) and the buffered batch `uint8 [B][S+2o][S+2o][3]` is
assembled by one kernel (`rsb_stitch_halo`) straight into the predictor's input buffer: bit-identical to the reference's
canvas, 9x fewer decodes, and the (S+2o)^2 buffered images never cross PCIe.
"""

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image

from robosat_b200 import _lib
from robosat_b200.tiles import Tile

NEIGHBOURS = [(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]  # row-major over (dy, dx): entry 4 is the tile itself


def neighbour_keys(tile):
    """The 9 tiles (itself included) whose pixels appear in the buffered image of `tile`, in table order."""
    return [Tile(x=int(tile.x) + dx, y=int(tile.y) + dy, z=int(tile.z)) for dx, dy in NEIGHBOURS]


NATIVE_PNG = os.environ.get("RSB_PNG_DECODER", "native") != "pil"


def decode_rgb(path, size, out=None):
    """`Image.open(path).convert("RGB")` as uint8 [S, S, 3] (robosat/tiles.py:150-159, 181), optionally straight into `out`
    (a [S, S, 3] uint8 view of pinned staging memory: no intermediate copy).

    8-bit PNG tiles go through the library's own decoder (`rsb_png_read_rgb`: file read + inflate + unfilter in C, no
    interpreter lock held, so the pool threads decode in parallel); every other format (JPEG, WebP, 16-bit / interlaced PNG)
    is PIL's job exactly as in the reference."""
    if NATIVE_PNG and path.lower().endswith(".png"):
        arr = out if out is not None else np.empty((size, size, 3), dtype=np.uint8)
        rc = _lib.load().rsb_png_read_rgb(os.fsencode(path), arr.ctypes.data, size, size)
        if rc == 0:
            return arr
        if rc != _lib.RSB_E_UNSUPPORTED:
            raise _lib.RsbError("decoding %s failed: %s" % (path, _lib.last_error()))
    arr = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
    assert arr.shape == (size, size, 3), "tile %s is %s, expected %dx%d" % (path, arr.shape, size, size)
    if out is not None:
        out[...] = arr
        return out
    return arr


def decode_many(jobs, size, threads):
    """jobs: [(path, out uint8 [S, S, 3])]. 8-bit PNG files are decoded by ONE library call that fans them out over `threads`
    threads of its own (no interpreter lock, no per-file Python work); anything it reports as unsupported, and every non-PNG
    file, goes through `decode_rgb` (PIL) like in the reference."""
    import ctypes

    rest = []
    png = [(p, o) for p, o in jobs if NATIVE_PNG and p.lower().endswith(".png")]
    rest.extend((p, o) for p, o in jobs if not (NATIVE_PNG and p.lower().endswith(".png")))
    if png:
        n = len(png)
        paths = (ctypes.c_char_p * n)(*[os.fsencode(p) for p, _ in png])
        outs = (ctypes.c_void_p * n)(*[o.ctypes.data for _, o in png])
        rcs = (ctypes.c_int32 * n)()
        rc = _lib.load().rsb_png_read_rgb_batch(paths, n, outs, size, size, threads, rcs)
        if rc != 0:
            bad = [png[i][0] for i in range(n) if rcs[i] not in (0, _lib.RSB_E_UNSUPPORTED)]
            raise _lib.RsbError("decoding %s failed (rc=%d)" % (bad[:3], rc))
        rest.extend(png[i] for i in range(n) if rcs[i] == _lib.RSB_E_UNSUPPORTED)
    for p, o in rest:
        arr = np.asarray(Image.open(p).convert("RGB"), dtype=np.uint8)
        assert arr.shape == (size, size, 3), "tile %s is %s, expected %dx%d" % (p, arr.shape, size, size)
        o[...] = arr


class DeviceTileCache:
    """Decoded tiles resident on `device`; `ensure` decodes + uploads what is missing, evicting least-recently-used slots."""

    def __init__(self, index, tile_size, capacity, device="cuda", workers=8):
        self.index = index  # Tile -> path
        self.size = tile_size
        self.capacity = capacity
        self.device = torch.device(device)
        pin = self.device.type == "cuda"
        self.store = torch.zeros((capacity, tile_size, tile_size, 3), dtype=torch.uint8, device=self.device)
        self._slot = {}      # Tile -> slot
        self._owner = [None] * capacity
        self._stamp = [0] * capacity
        self._clock = 0
        self._free = list(range(capacity - 1, -1, -1))
        # pinned staging, three regions used by successive tickets: the pool threads decode STRAIGHT into one (no host copy)
        # while an earlier ticket waits to be committed in the second and the uploads of a third are still in flight
        self._half = max(16, min(capacity, int(os.environ.get("RSB_STAGING_TILES", "128"))))
        self._regions = 3
        self._staging = torch.empty((self._regions * self._half, tile_size, tile_size, 3), dtype=torch.uint8, pin_memory=pin)
        self._staging_np = self._staging.numpy()
        self._uploaded = [torch.cuda.Event() if pin else None for _ in range(self._regions)]
        self._tickets = 0
        self._protected = {}  # ticket number -> tiles a prefetched-but-not-yet-stitched batch needs (never evicted meanwhile)
        # uploads run on their own stream, behind the last kernel that read the store, so they overlap the previous batch's network
        self._copy_stream = torch.cuda.Stream(device=self.device) if pin else None
        self._store_read = torch.cuda.Event() if pin else None
        self.decode_threads = workers  # threads the library uses per ticket (C++ threads; the Python pool only carries the call)
        self._pool = ThreadPoolExecutor(max_workers=2)
        self.decodes = 0
        self.hits = 0
        self.decode_wait_s = 0.0  # time the caller was blocked waiting for decodes (what is left on the critical path)

    def slot(self, tile):
        return self._slot.get(tile, -1)

    def _take_slot(self, keep):
        if self._free:
            return self._free.pop()
        victim = min((s for s in range(self.capacity) if self._owner[s] not in keep), key=lambda s: self._stamp[s], default=None)
        assert victim is not None, "tile cache too small for one batch: raise capacity (need >= 9 x batch)"
        del self._slot[self._owner[victim]]
        return victim

    def ensure(self, tiles):
        """Make every tile of `tiles` that exists in the store's index resident. Uploads are enqueued on the current stream."""
        self.commit(self.prefetch(tiles))

    def prefetch(self, tiles):
        """First half of `ensure`: decide which tiles are missing, reserve their slots and START decoding them on the pool
        threads; returns a ticket for `commit`. Between the two calls the caller is free to enqueue GPU work and wait for
        earlier batches, so the decode of batch i+1 runs while the GPU computes batch i (decode off the critical path)."""
        self._clock += 1
        want = [t for t in dict.fromkeys(tiles) if t in self.index]
        keep = set(want).union(*self._protected.values()) if self._protected else set(want)
        missing = []
        for t in want:
            s = self._slot.get(t)
            if s is None:
                missing.append(t)
            else:
                self._stamp[s] = self._clock
                self.hits += 1
        number = self._tickets
        half = number % self._regions
        self._tickets += 1
        self._protected[number] = set(want)
        if self._uploaded[half] is not None:
            self._uploaded[half].synchronize()  # the uploads of the ticket that last used this half have left the staging memory
        items, jobs = [], []
        for j, t in enumerate(missing):
            s = self._take_slot(keep)
            self._slot[t], self._owner[s], self._stamp[s] = s, t, self._clock
            # the first `_half` tiles decode into this ticket's half of the pinned staging; an oversized ticket (the first batch
            # of a run can miss 9 x batch tiles) decodes the rest into ordinary arrays and stages them at commit time
            dst = self._staging_np[half * self._half + j] if j < self._half else np.empty((self.size, self.size, 3), dtype=np.uint8)
            items.append((s, dst, j < self._half))
            jobs.append((self.index[t], dst))
        self.decodes += len(missing)
        # ONE pool job per ticket: the library fans the PNG files out over its own threads (rsb_png_read_rgb_batch)
        return {"number": number, "half": half, "items": items,
                "future": self._pool.submit(decode_many, jobs, self.size, self.decode_threads) if jobs else None}

    def commit(self, ticket):
        """Second half of `ensure`: wait for the decodes of `ticket` and enqueue their uploads on the current stream (i.e. after
        every kernel already enqueued that still reads the slots being replaced)."""
        import time

        half, items = ticket["half"], ticket["items"]
        base = half * self._half
        if ticket["future"] is not None:
            t0 = time.perf_counter()
            ticket["future"].result()
            self.decode_wait_s += time.perf_counter() - t0
        main = torch.cuda.current_stream(self.device) if self._copy_stream is not None else None

        def upload(pairs):
            if self._copy_stream is None:
                for s, j in pairs:
                    self.store[s].copy_(self._staging[base + j], non_blocking=True)
                return
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._store_read)  # every stitch kernel enqueued so far has finished reading the store
                for s, j in pairs:
                    self.store[s].copy_(self._staging[base + j], non_blocking=True)
                self._uploaded[half].record(self._copy_stream)
            main.wait_event(self._uploaded[half])  # the next stitch kernel (on the caller's stream) sees the new tiles

        upload([(s, j) for j, (s, _, direct) in enumerate(items) if direct])
        extra = [(s, arr) for s, arr, direct in items if not direct]
        for i in range(0, len(extra), self._half):  # oversized ticket (first batch of a run): reuse the region, synchronously
            part = extra[i:i + self._half]
            if self._uploaded[half] is not None:
                self._uploaded[half].synchronize()
            for j, (s, arr) in enumerate(part):
                self._staging_np[base + j] = arr
            upload([(s, j) for j, (s, _) in enumerate(part)])
        self._protected.pop(ticket["number"], None)  # the caller builds its slot table next; later prefetches may evict again

    def mark_store_read(self):
        """call after enqueuing a kernel that reads the store (the stitch kernel): later uploads wait for it"""
        if self._store_read is not None:
            self._store_read.record(torch.cuda.current_stream(self.device))

    def table(self, tiles):
        """int32 [len(tiles), 9] slot table for `rsb_stitch_halo` (-1 where the store has no such neighbour)."""
        return torch.tensor([[self.slot(k) for k in neighbour_keys(t)] for t in tiles], dtype=torch.int32).reshape(len(tiles), 9)

    def close(self):
        self._pool.shutdown(wait=True)


class HaloStitcher:
    """Buffered batches straight from the tile cache: `stitch(tiles, out)` fills uint8 [B, S+2o, S+2o, 3] on the device."""

    def __init__(self, cache, overlap, batch):
        assert cache.capacity >= 9 * batch, "tile cache must hold one batch with all its neighbours"
        self.cache, self.overlap, self.batch = cache, overlap, batch
        self.full = cache.size + 2 * overlap
        self._tables = [torch.empty((batch, 9), dtype=torch.int32, pin_memory=cache.device.type == "cuda") for _ in range(2)]
        self._dtable = torch.empty((batch, 9), dtype=torch.int32, device=cache.device)
        self._used = [torch.cuda.Event() if cache.device.type == "cuda" else None for _ in range(2)]
        self._n = 0
        self._pending = {}

    def prefetch(self, tiles):
        """Start decoding what `stitch(tiles, ...)` will need. Up to two batches may be prefetched ahead of the one being
        stitched (the cache keeps their tiles resident until their own `stitch`)."""
        self._pending[id(tiles)] = (tiles, self.cache.prefetch([k for t in tiles for k in neighbour_keys(t)]))

    def stitch(self, tiles, out):
        """tiles: <= batch Tile keys (missing rows of a ragged last batch become black); out: device uint8 [batch, F, F, 3]."""
        assert len(tiles) <= self.batch and tuple(out.shape) == (self.batch, self.full, self.full, 3) and out.dtype == torch.uint8
        pending = self._pending.pop(id(tiles), None)
        if pending is not None and pending[0] is tiles:
            self.cache.commit(pending[1])  # decodes started by prefetch(tiles)
        else:
            self.cache.ensure([k for t in tiles for k in neighbour_keys(t)])
        i = self._n % 2
        self._n += 1
        if self._used[i] is not None:
            self._used[i].synchronize()
        host = self._tables[i]
        host.fill_(-1)
        host[:len(tiles)] = self.cache.table(tiles)
        self._dtable.copy_(host, non_blocking=True)
        _lib.check(_lib.load().rsb_stitch_halo(self.cache.store.data_ptr(), self._dtable.data_ptr(), out.data_ptr(), self.batch, self.cache.size,
                                               self.overlap, _lib.current_stream_ptr()), "rsb_stitch_halo")
        self.cache.mark_store_read()
        if self._used[i] is not None:
            self._used[i].record(torch.cuda.current_stream(self.cache.device))
        return out
