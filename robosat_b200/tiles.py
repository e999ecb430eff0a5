"""Slippy-map tile enumeration and halo stitching (`z/x/y.ext` directories).

Interface kept from the reference (robosat/tiles.py:65-227) because the dataset classes and the `rs predict`
loop are built on it: `tiles_from_slippy_map`, `tiles_from_csv`, `adjacent_tile`, `buffer_tile_image`.
Differences: `mercantile` is optional (a namedtuple with the same fields stands in), and the tile -> path lookup
is a dictionary built ONCE by the caller instead of `dict(tiles)` on every call (the reference's O(N^2),
robosat/tiles.py:177) -- passing a list still works.
"""

import csv
import os
from collections import namedtuple

from PIL import Image

try:  # the reference imports mercantile unconditionally; only its Tile type is used on this path
    from mercantile import Tile
except ImportError:  # pragma: no cover - depends on the environment
    Tile = namedtuple("Tile", ["x", "y", "z"])


def _is_int(text):
    try:
        int(text)
    except ValueError:
        return False
    return True


def tiles_from_slippy_map(root):
    """Yield (Tile, path) for every `root/z/x/y.*` file whose three components parse as integers."""
    for z in os.listdir(root):
        if not _is_int(z):
            continue
        zdir = os.path.join(root, z)
        for x in os.listdir(zdir):
            if not _is_int(x):
                continue
            xdir = os.path.join(zdir, x)
            for name in os.listdir(xdir):
                stem = os.path.splitext(name)[0]
                if _is_int(stem):
                    yield Tile(x=int(x), y=int(stem), z=int(z)), os.path.join(xdir, name)


def tiles_from_csv(path):
    """Yield a Tile per non-empty `x,y,z` row."""
    with open(path) as fp:
        for row in csv.reader(fp):
            if row:
                yield Tile(*(int(v) for v in row))


def adjacent_tile(tile, dx, dy, tiles):
    """RGB image of the neighbour at offset (dx, dy), or None when the store has no such tile."""
    key = Tile(x=int(tile.x) + dx, y=int(tile.y) + dy, z=int(tile.z))
    try:
        return Image.open(tiles[key]).convert("RGB")
    except KeyError:
        return None


def buffer_tile_image(tile, tiles, overlap, tile_size, nodata=0):
    """Tile plus an `overlap`-pixel border cut from its 8 neighbours; `nodata` where a neighbour is missing.

    Returns a (tile_size + 2*overlap)^2 RGB image. `tiles` is a mapping Tile -> path (a list of pairs is converted).
    """
    if not isinstance(tiles, dict):
        tiles = dict(tiles)
    o, s = overlap, tile_size
    full = s + 2 * o
    canvas = Image.new(mode="RGB", size=(full, full), color=nodata)
    canvas.paste(Image.open(tiles[tile]).convert("RGB"), box=(o, o))
    if o == 0:
        return canvas
    # (dx, dy) -> (destination box on the canvas, source box inside the neighbour), both (left, upper, right, lower)
    spans = {-1: ((0, o), (s - o, s)), 0: ((o, o + s), (0, s)), 1: ((o + s, full), (0, o))}
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            other = adjacent_tile(tile, dx, dy, tiles)
            if other is None:
                continue
            (cx0, cx1), (sx0, sx1) = spans[dx]
            (cy0, cy1), (sy0, sy1) = spans[dy]
            canvas.paste(other.crop(box=(sx0, sy0, sx1, sy1)), box=(cx0, cy0, cx1, cy1))
    return canvas
