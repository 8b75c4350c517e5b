"""Tile-code helpers and LAS I/O (host side; mirrors ``upcp.utils.las_utils``, reference
src/upcp/utils/las_utils.py:12-53,118-130).  ``laspy`` is used when it is installed; without it ``las_io`` reads
and writes ``.las`` and -- through the native LASzip decoder / encoder of csrc/laz.cu -- ``.laz`` (point formats 0-3)."""
import logging
import pathlib
import re

from . import las_io


logger = logging.getLogger(__name__)


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


def read_las(las_file):
    """Read a las file and return the las object (las_utils.py:118-120)."""
    try:
        import laspy
        laspy_read = laspy.read                 # (a stand-in module without the reader counts as "not installed")
    except (ImportError, AttributeError):
        # native reader: plain .las and LASzip-compressed .laz (point formats 0-3, csrc/laz.cu)
        return las_io.read(las_file)
    return laspy_read(las_file)


def label_and_save_las(las, labels, outfile):
    """Label a las file using the provided class labels and save to outfile (las_utils.py:123-130)."""
    assert len(labels) == las.header.point_count
    if isinstance(las, las_io.LasData):
        las.label = labels          # adds the uint8 "label" extra dimension when missing
        if str(outfile).lower().endswith('.laz') and las.header.point_format_id > 3:
            # the native LASzip encoder covers point formats 0-3; LAS 1.4 formats are written uncompressed under
            # the requested name (readers go by the header's compression bit, not by the extension)
            logger.warning(f'{outfile}: point format {las.header.point_format_id} is written uncompressed')
        las.write(outfile)          # `.laz`: LASzip-compressed through csrc/laz.cu, as laspy[lazrs] would
        return
    import laspy
    if 'label' not in las.point_format.extra_dimension_names:
        las.add_extra_dim(laspy.ExtraBytesParams(name='label', type='uint8', description='Labels'))
    las.label = labels
    las.write(outfile)
