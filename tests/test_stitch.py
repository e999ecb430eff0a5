"""Device-side halo stitching (robosat_b200/stitch.py, `rsb_stitch_halo`) == the reference's `buffer_tile_image`
(robosat/tiles.py:162-227; our restatement in robosat_b200/tiles.py is pinned to the reference by tests/golden/tiles.npz),
bit for bit -- byte work. CPU: cache / table logic + a numpy restatement of the kernel's index arithmetic; GPU: the kernel."""

import os

import numpy as np
import pytest
import torch
from PIL import Image

import emulate
from robosat_b200 import synth
from robosat_b200.stitch import DeviceTileCache, HaloStitcher, neighbour_keys
from robosat_b200.tiles import Tile, buffer_tile_image, tiles_from_slippy_map

SIZE = 64


def _make_store(root, coords, seed=11):
    u8 = synth.make_tiles_u8(len(coords), SIZE, seed=seed).numpy()
    for (x, y), arr in zip(coords, u8):
        os.makedirs(os.path.join(root, "18", str(x)), exist_ok=True)
        Image.fromarray(arr).save(os.path.join(root, "18", str(x), "%d.png" % y))
    return dict(tiles_from_slippy_map(root))


# a 3x3 block with a hole, a tile whose only neighbour is diagonal, and an isolated tile
COORDS = [(10, 20), (11, 20), (12, 20), (10, 21), (12, 21), (10, 22), (11, 22), (12, 22), (13, 23), (30, 40)]


def _reference(index, tile, overlap):
    return np.asarray(buffer_tile_image(tile, index, overlap=overlap, tile_size=SIZE))


@pytest.mark.parametrize("overlap", [0, 6, 8, 32, 64])
def test_table_and_index_arithmetic_match_buffer_tile_image(tmp_path, overlap):
    index = _make_store(str(tmp_path), COORDS)
    cache = DeviceTileCache(index, SIZE, capacity=32, device="cpu", workers=2)
    tiles = sorted(index)
    cache.ensure([k for t in tiles for k in neighbour_keys(t)])
    assert cache.decodes == len(COORDS)  # every file decoded exactly once although it is needed for up to 9 buffered tiles
    table = cache.table(tiles).numpy()
    assert (table[:, 4] >= 0).all() and (table[tiles.index(Tile(30, 40, 18))] == [-1, -1, -1, -1, table[tiles.index(Tile(30, 40, 18)), 4], -1, -1, -1, -1]).all()
    got = emulate.stitch_halo_cpu(cache.store.numpy(), table, SIZE, overlap)
    for i, t in enumerate(tiles):
        assert np.array_equal(got[i], _reference(index, t, overlap)), (t, overlap)
    cache.close()


def test_cache_evicts_least_recently_used_and_redecodes(tmp_path):
    index = _make_store(str(tmp_path), [(x, 5) for x in range(8)])
    cache = DeviceTileCache(index, SIZE, capacity=4, device="cpu", workers=1)
    t = [Tile(x, 5, 18) for x in range(8)]
    cache.ensure(t[:4])
    assert cache.decodes == 4 and all(cache.slot(k) >= 0 for k in t[:4])
    cache.ensure([t[0], t[1]])            # refresh 0 and 1
    cache.ensure([t[4], t[5]])            # must evict 2 and 3, the least recently used
    assert cache.slot(t[2]) == -1 and cache.slot(t[3]) == -1 and cache.slot(t[0]) >= 0 and cache.slot(t[5]) >= 0
    cache.ensure([t[2]])                  # comes back with a second decode
    assert cache.decodes == 7 and cache.slot(t[2]) >= 0
    slots = [cache.slot(k) for k in t if cache.slot(k) >= 0]
    assert len(slots) == len(set(slots)) == 4
    with pytest.raises(AssertionError):
        cache.ensure(t[:5])               # more than the cache can hold at once
    cache.close()


@pytest.mark.gpu
@pytest.mark.parametrize("overlap,batch", [(32, 4), (0, 3), (64, 2), (6, 4)])  # overlap 6: the byte-wise kernel
def test_device_stitch_bit_exact(tmp_path, overlap, batch, cuda_device):
    index = _make_store(str(tmp_path), COORDS)
    cache = DeviceTileCache(index, SIZE, capacity=9 * batch, device=cuda_device, workers=4)  # minimal capacity: evictions happen
    stitcher = HaloStitcher(cache, overlap, batch)
    F = SIZE + 2 * overlap
    out = torch.zeros((batch, F, F, 3), dtype=torch.uint8, device=cuda_device)
    tiles = sorted(index)
    for i in range(0, len(tiles), batch):
        part = tiles[i:i + batch]
        out.fill_(7)
        stitcher.stitch(part, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for j, t in enumerate(part):
            assert np.array_equal(got[j], _reference(index, t, overlap)), (t, overlap)
        assert (got[len(part):] == 0).all()  # ragged last batch: black padding tiles
    cache.close()


@pytest.mark.gpu
def test_stitch_kernel_bandwidth(cuda_device):
    """HBM-bound byte work: report achieved GB/s of the stitch kernel at the predict shape (batch 32, 512 + 2*32)."""
    from robosat_b200 import _lib

    lib = _lib.load()
    B, S, o = 32, 512, 32
    F = S + 2 * o
    store = torch.randint(0, 256, (9 * B, S, S, 3), dtype=torch.uint8, device=cuda_device)
    table = torch.arange(9 * B, dtype=torch.int32, device=cuda_device).reshape(B, 9).contiguous()
    out = torch.empty((B, F, F, 3), dtype=torch.uint8, device=cuda_device)
    st = _lib.current_stream_ptr()
    for _ in range(3):
        _lib.check(lib.rsb_stitch_halo(store.data_ptr(), table.data_ptr(), out.data_ptr(), B, S, o, st), "stitch")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.check(lib.rsb_stitch_halo(store.data_ptr(), table.data_ptr(), out.data_ptr(), B, S, o, st), "stitch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = 2 * out.numel() / 1e9  # every canvas byte is read once and written once
    print("rsb_stitch_halo: %.3f ms per batch of %d, %.0f GB/s algorithmic (read + write), %.0f tiles/s" % (ms, B, gb / ms * 1e3, B / ms * 1e3))
    assert gb / ms * 1e3 > 500
