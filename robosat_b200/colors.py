"""Colour palettes for the paletted PNG outputs (robosat/colors.py:45-95)."""

import colorsys

_HEX = {
    "dark": "#404040", "gray": "#eeeeee", "light": "#f8f8f8", "white": "#ffffff", "cyan": "#3bb2d0", "blue": "#3887be",
    "bluedark": "#223b53", "denim": "#50667f", "navy": "#28353d", "navydark": "#222b30", "purple": "#8a8acb",
    "teal": "#41afa5", "green": "#56b881", "yellow": "#f1f075", "mustard": "#fbb03b", "orange": "#f9886c",
    "red": "#e55e5e", "pink": "#ed6498",
}
MAPBOX = {name: tuple(int(h[i:i + 2], 16) for i in (1, 3, 5)) for name, h in _HEX.items()}


def make_palette(*colors):
    """Flat [r0, g0, b0, r1, ...] list for PIL's putpalette."""
    return [channel for name in colors for channel in MAPBOX[name]]


def continuous_palette_for_color(color, bins=256):
    """`bins` shades of one colour with saturation ramping from 1/bins to 1 (HSV), flattened for putpalette."""
    r, g, b = (c / 255 for c in MAPBOX[color])
    h, _, v = colorsys.rgb_to_hsv(r, g, b)
    palette = []
    for i in range(bins):
        palette.extend(int(c * 255) for c in colorsys.hsv_to_rgb(h, (1 / bins) * (i + 1), v))
    assert len(palette) == 3 * bins
    return palette
