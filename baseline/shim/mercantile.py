"""Three-line stand-in for the `mercantile` package (absent from this image, no network): the reference's hot path only uses
`mercantile.Tile` as a record type (robosat/tiles.py, robosat/datasets.py). Bench / test infrastructure, never on the product path."""
from collections import namedtuple

Tile = namedtuple("Tile", ["x", "y", "z"])
