"""Tile-code helpers (host side; mirrors the tile-code part of
``upcp.utils.las_utils``, reference src/upcp/utils/las_utils.py:12-53).
LAS/LAZ I/O is the "next" row of the scope table and is not provided here."""
import pathlib
import re


def get_tilecode_from_filename(filename):
    """Extract the tile code (``XXXX_YYYY``) from a file name."""
    return re.match(r'.*(\d{4}_\d{4}).*', filename)[1]


def get_tilecodes_from_folder(las_folder, las_prefix=''):
    """Set of unique tile codes of the ``.laz`` files in a folder."""
    files = pathlib.Path(las_folder).glob(f'{las_prefix}*.laz')
    return set(get_tilecode_from_filename(file.name) for file in files)


def get_bbox_from_tile_code(tile_code, padding=0, width=50, height=50):
    """<X,Y> bounding box of a tile; the code encodes the lower-left corner / 50.

    Returns ``((x_min, y_max), (x_max, y_min))`` (inverted y-axis).  Note that, as in
    the reference (las_utils.py:52-53), the x extent uses ``height``.
    """
    tile_split = tile_code.split('_')
    x_min = int(tile_split[0]) * 50
    y_min = int(tile_split[1]) * 50
    return ((x_min - padding, y_min + height + padding),
            (x_min + height + padding, y_min - padding))
