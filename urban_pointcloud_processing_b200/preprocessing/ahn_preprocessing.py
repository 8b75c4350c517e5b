"""AHN preprocessing: ground and building surface rasters of one tile from an AHN point cloud -- the
producer of the ``ahn_<tilecode>.npz`` files the labelling path reads (drop-in for the surface part of
``upcp.preprocessing.ahn_preprocessing``: ``_get_ahn_surface`` :129-175, ``process_ahn_las_tile`` :178-236;
SURVEY 8f "next" row 4).  The interpolation itself (8 nearest within 1 m, inverse squared distance, for the
ground; maximum of the 8 nearest within 0.5 m for buildings) runs on the GPU (``upcp_interpolate``); the
rounding to 2 decimals and float16 of the reference is kept, so the files are interchangeable.

Reading goes through ``utils.las_io`` (plain ``.las``) or ``laspy`` when it is installed (``.laz``); clipping the
country-wide AHN cloud into tiles (``clip_ahn_las_tile``) is file plumbing and not part of this module."""
import os
import pathlib

import numpy as np

from ..utils import las_utils
from ..utils.interpolation import SpatialInterpolator

# AHN classification codes (see https://www.ahn.nl/4-classificatie)
AHN_OTHER = 1
AHN_GROUND = 2
AHN_BUILDING = 6
AHN_WATER = 9
AHN_ARTIFACT = 26


def _get_ahn_surface(ahn_las, grid_x, grid_y, si_method, ahn_label, n_neighbors=8, max_dist=1, power=2.,
                     fill_value=np.nan):
    """Surface raster for the points of one AHN class: float16 array of the shape of ``grid_x``, NaN where no
    point lies within ``max_dist`` (ahn_preprocessing.py:129-175)."""
    mask = np.asarray(ahn_las.classification) == ahn_label
    if np.count_nonzero(mask) <= 1:
        return np.full(grid_x.shape, np.nan, dtype='float16')
    points = np.vstack((np.asarray(ahn_las.x), np.asarray(ahn_las.y), np.asarray(ahn_las.z))).T[mask]
    positions = np.vstack((grid_x.reshape(-1), grid_y.reshape(-1))).T
    idw = SpatialInterpolator(points[:, 0:2], points[:, 2], method=si_method)
    ahn_gnd_grid = idw(positions, n_neighbors=n_neighbors, max_dist=max_dist, power=power, fill_value=fill_value)
    return np.around(ahn_gnd_grid.reshape(grid_x.shape), decimals=2).astype('float16')


def ahn_surfaces(ahn_las, tile_code, resolution=0.1):
    """(x, y, ground, building) of the tile -- what ``process_ahn_las_tile`` writes."""
    ((x_min, y_max), (x_max, y_min)) = las_utils.get_bbox_from_tile_code(tile_code)
    grid_y, grid_x = np.mgrid[y_max - resolution / 2:y_min:-resolution, x_min + resolution / 2:x_max:resolution]
    ground_surface = _get_ahn_surface(ahn_las, grid_x, grid_y, 'idw', AHN_GROUND)
    building_surface = _get_ahn_surface(ahn_las, grid_x, grid_y, 'max', AHN_BUILDING, max_dist=0.5)
    return grid_x[0, :], grid_y[:, 0], ground_surface, building_surface


def _read_ahn(ahn_las_file):
    try:
        import laspy
        return laspy.read(ahn_las_file)
    except ImportError:
        from ..utils import las_io
        return las_io.read(ahn_las_file)


def process_ahn_las_tile(ahn_las_file, out_folder='', resolution=0.1):
    """Generate ground and building surfaces for a given AHN point cloud file and save them as
    ``ahn_<tilecode>.npz`` (keys x, y, ground, building).  Returns the path of the output file."""
    if isinstance(ahn_las_file, pathlib.PurePath):
        ahn_las_file = ahn_las_file.as_posix()
    tile_code = las_utils.get_tilecode_from_filename(ahn_las_file)
    if out_folder != '':
        pathlib.Path(out_folder).mkdir(parents=True, exist_ok=True)
    x, y, ground, building = ahn_surfaces(_read_ahn(ahn_las_file), tile_code, resolution)
    filename = os.path.join(out_folder, 'ahn_' + tile_code + '.npz')
    np.savez_compressed(filename, x=x, y=y, ground=ground, building=building)
    return filename
