"""AHN elevation readers for the labelling path (drop-ins for the ★ part of
``upcp.utils.ahn_utils``: AHNReader :25-102, NPZReader :105-138, load_ahn_tile
:287-302, AHNFileNotFoundError :511).  The readers stay on the host (file I/O and
per-tile caching); the per-point lookup runs on the GPU through ``DeviceRaster``.
GeoTIFF reading and gap filling are out of scope (offline raster preparation)."""
import logging
import os
from abc import ABC, abstractmethod
from pathlib import Path

import numpy as np

from .. import device as dev

logger = logging.getLogger(__name__)


class AHNFileNotFoundError(Exception):
    """Exception raised for missing AHN files."""


def load_ahn_tile(ahn_file):
    """Load the ground and building surface grids of an AHN ``.npz`` file as a dict with
    keys 'x', 'y', 'ground_surface', 'building_surface' (surfaces as float64)."""
    if not os.path.isfile(ahn_file):
        raise AHNFileNotFoundError(f'Tried loading {ahn_file} but file does not exist.')
    ahn = np.load(ahn_file)
    return {'x': ahn['x'], 'y': ahn['y'],
            'ground_surface': ahn['ground'].astype(float),
            'building_surface': ahn['building'].astype(float)}


class AHNReader(ABC):
    """Abstract class for reading AHN data."""

    NAME = None

    def __init__(self, data_folder, caching):
        super().__init__()
        self.path = Path(data_folder)
        self.set_caching(caching)
        self._clear_cache()
        if not self.path.exists():
            print('Input folder does not exist.')
            raise ValueError

    @abstractmethod
    def filter_tile(self, tilecode):
        pass

    def set_caching(self, state):
        if not hasattr(self, 'caching') or (self.caching is not state):
            self.caching = state
            if not self.caching:
                self._clear_cache()
                logger.debug('Caching disabled.')
            else:
                logger.debug('Caching enabled.')

    def _clear_cache(self):
        self.cache = {'tilecode': ''}
        self._rasters = {}

    # -- device side -----------------------------------------------------------
    def device_raster(self, tilecode, surface='ground_surface'):
        """HBM-resident raster (bin edges + values) of one surface of one tile."""
        key = (tilecode, surface, str(dev.current_device()), dev.stream_key())
        raster = self._rasters.get(key)
        if raster is None:
            ahn_tile = self.filter_tile(tilecode)
            if surface not in ahn_tile:
                logger.error(f'Unknown surface: {surface}.')
                raise ValueError
            if len(self._rasters) > 8:
                self._rasters.clear()
            raster = dev.DeviceRaster(ahn_tile['x'], ahn_tile['y'], ahn_tile[surface])
            self._rasters[key] = raster
        return raster

    # -- reference API -----------------------------------------------------------
    def cache_interpolator(self, tilecode, points, surface='ground_surface'):
        logger.info(f'Caching {surface} for tile {tilecode}.')
        self.set_caching(True)
        if self.cache['tilecode'] != tilecode:
            self._clear_cache()
        ahn_tile = self.filter_tile(tilecode)
        if surface not in ahn_tile:
            logger.error(f'Unknown surface: {surface}.')
            raise ValueError
        self.cache['tilecode'] = tilecode
        tile = dev.get_tile(points)
        self.cache[surface] = self.device_raster(tilecode, surface).lookup(tile).cpu().numpy()

    def interpolate(self, tilecode, points=None, mask=None, surface='ground_surface'):
        if points is None and mask is None:
            logger.error('Must provide either points or mask.')
            raise ValueError
        if self.caching and mask is not None:
            if self.cache['tilecode'] == tilecode:
                if surface in self.cache:
                    return self.cache[surface][mask]
                else:
                    logger.debug(f'Surface {surface} not in cache for tile {tilecode}.')
        elif self.caching:
            logger.debug('Caching enabled but no mask provided.')
        if self.cache['tilecode'] != tilecode and points is None:
            logger.error(f'Tile {tilecode} not cached and no points provided.')
            raise ValueError
        ahn_tile = self.filter_tile(tilecode)
        if surface not in ahn_tile:
            logger.error(f'Unknown surface: {surface}.')
            raise ValueError
        pts = np.asarray(points, dtype=np.float64)
        tile = dev.DeviceTile(np.ascontiguousarray(pts[:, :2]))
        return self.device_raster(tilecode, surface).lookup(tile).cpu().numpy()


class NPZReader(AHNReader):
    """Reader for pre-processed AHN ``.npz`` tiles (``ahn_<tilecode>.npz``).

    Parameters
    ----------
    data_folder : str or Path
        Folder containing the .npz files.
    caching : bool (default: True)
        Enable caching of the current ahn tile and interpolation data.
    """

    NAME = 'npz'

    def __init__(self, data_folder, caching=True):
        super().__init__(data_folder, caching)

    def filter_tile(self, tilecode):
        """AHN tile dict for the area of the given CycloMedia tile code."""
        fname = os.path.join(self.path, 'ahn_' + tilecode + '.npz')
        if self.caching:
            if self.cache['tilecode'] != tilecode:
                self._clear_cache()
                self.cache['tilecode'] = tilecode
            if 'ahn_tile' not in self.cache:
                self.cache['ahn_tile'] = load_ahn_tile(fname)
            return self.cache['ahn_tile']
        return load_ahn_tile(fname)


class ArrayAHNReader(AHNReader):
    """In-memory AHN reader (synthetic tiles, tests): ``tiles`` maps a tile code to a dict
    with keys 'x', 'y', 'ground_surface', 'building_surface'."""

    NAME = 'npz'

    def __init__(self, tiles, caching=True):
        self.tiles = tiles
        self.path = Path('.')
        self.set_caching(caching)
        self._clear_cache()

    def filter_tile(self, tilecode):
        if tilecode not in self.tiles:
            raise AHNFileNotFoundError(f'No AHN data for tile {tilecode}.')
        if self.caching and self.cache['tilecode'] != tilecode:
            self._clear_cache()
            self.cache['tilecode'] = tilecode
        return self.tiles[tilecode]


def device_raster_for(ahn_reader, tilecode, surface):
    """DeviceRaster for ANY reader object honouring the reference reader protocol
    (``filter_tile(tilecode)`` returning the tile dict) -- including the reference's
    own ``upcp.utils.ahn_utils.NPZReader`` instances."""
    if hasattr(ahn_reader, 'device_raster'):
        return ahn_reader.device_raster(tilecode, surface)
    cache = getattr(ahn_reader, '_upcp_b200_rasters', None)
    if cache is None:
        cache = {}
        try:
            ahn_reader._upcp_b200_rasters = cache
        except AttributeError:
            pass
    key = (tilecode, surface, str(dev.current_device()), dev.stream_key())
    raster = cache.get(key)
    if raster is None:
        ahn_tile = ahn_reader.filter_tile(tilecode)
        if surface not in ahn_tile:
            logger.error(f'Unknown surface: {surface}.')
            raise ValueError
        if len(cache) > 8:
            cache.clear()
        raster = dev.DeviceRaster(ahn_tile['x'], ahn_tile['y'], ahn_tile[surface])
        cache[key] = raster
    return raster
