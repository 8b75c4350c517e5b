"""AHN elevation readers for the labelling path (drop-ins for the ★ part of
``upcp.utils.ahn_utils``: AHNReader :25-102, NPZReader :105-138, load_ahn_tile
:287-302, AHNFileNotFoundError :511).  The readers stay on the host (file I/O and
per-tile caching); the per-point lookup runs on the GPU through ``DeviceRaster``.
GeoTIFF reading and gap filling are out of scope (offline raster preparation)."""
import logging
import os
import threading
from abc import ABC, abstractmethod
from collections import OrderedDict
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


_RASTER_SLOTS = 32          # device rasters kept per reader (tile x surface x stream), least recently used out
_TILE_SLOTS = 8             # host tiles kept per reader


class _LRU(OrderedDict):
    """Tiny least-recently-used map; the owner serialises access."""

    def __init__(self, slots):
        super().__init__()
        self.slots = slots

    def fetch(self, key):
        val = self.get(key)
        if val is not None:
            self.move_to_end(key)
        return val

    def store(self, key, val):
        self[key] = val
        self.move_to_end(key)
        while len(self) > self.slots:
            self.popitem(last=False)


class AHNReader(ABC):
    """Abstract class for reading AHN data.

    One reader is shared by every worker thread of ``Pipeline.process_clouds`` /
    ``process_folder(workers=W)``, each labelling ITS OWN tile.  Therefore nothing that a lookup
    returns is ever read back out of a "current tile" slot: host tiles and device rasters live in
    per-tilecode LRU maps behind one lock, and ``self.cache`` -- the reference's single-tile
    interpolation cache (ahn_utils.py:56-102) -- is only ever swapped as a whole."""

    NAME = None

    def __init__(self, data_folder, caching):
        super().__init__()
        self.path = Path(data_folder)
        self._init_state()
        self.set_caching(caching)
        if not self.path.exists():
            print('Input folder does not exist.')
            raise ValueError

    def _init_state(self):
        self._lock = threading.RLock()
        self._tiles = _LRU(_TILE_SLOTS)
        self._rasters = _LRU(_RASTER_SLOTS)
        self.cache = {'tilecode': ''}

    @abstractmethod
    def _load_tile(self, tilecode):
        """Read one tile from its source (no caching here)."""

    def filter_tile(self, tilecode):
        """AHN tile dict for the area of the given CycloMedia tile code (ahn_utils.py:123-138).
        Always the tile that was asked for, whatever other threads are doing."""
        if not self.caching:
            return self._load_tile(tilecode)
        with self._lock:
            tile = self._tiles.fetch(tilecode)
        if tile is None:
            tile = self._load_tile(tilecode)          # (outside the lock: file reads of different tiles overlap)
            with self._lock:
                self._tiles.store(tilecode, tile)
        return tile

    def set_caching(self, state):
        if not hasattr(self, 'caching') or (self.caching is not state):
            self.caching = state
            if not self.caching:
                self._clear_cache()
                with self._lock:
                    self._tiles.clear()
                    self._rasters.clear()
                logger.debug('Caching disabled.')
            else:
                logger.debug('Caching enabled.')

    def _clear_cache(self):
        # the reference's interpolation cache only; tiles and rasters are per-tilecode LRU maps
        self.cache = {'tilecode': ''}

    # -- device side -----------------------------------------------------------
    def device_raster(self, tilecode, surface='ground_surface'):
        """HBM-resident raster (bin edges + values) of one surface of one tile, per device and
        CUDA stream (tiles in flight on different streams never share a buffer)."""
        key = (tilecode, surface, str(dev.current_device()), dev.stream_key())
        if self.caching:
            with self._lock:
                raster = self._rasters.fetch(key)
            if raster is not None:
                return raster
        ahn_tile = self.filter_tile(tilecode)
        if surface not in ahn_tile:
            logger.error(f'Unknown surface: {surface}.')
            raise ValueError
        raster = dev.DeviceRaster(ahn_tile['x'], ahn_tile['y'], ahn_tile[surface])
        if self.caching:
            with self._lock:
                self._rasters.store(key, raster)
        return raster

    # -- reference API -----------------------------------------------------------
    def cache_interpolator(self, tilecode, points, surface='ground_surface'):
        logger.info(f'Caching {surface} for tile {tilecode}.')
        self.set_caching(True)
        ahn_tile = self.filter_tile(tilecode)
        if surface not in ahn_tile:
            logger.error(f'Unknown surface: {surface}.')
            raise ValueError
        tile = dev.get_tile(points)
        values = self.device_raster(tilecode, surface).lookup(tile).cpu().numpy()
        with self._lock:
            cache = dict(self.cache) if self.cache['tilecode'] == tilecode else {'tilecode': tilecode}
            cache[surface] = values
            self.cache = cache                        # swapped as a whole: readers hold a consistent snapshot

    def interpolate(self, tilecode, points=None, mask=None, surface='ground_surface'):
        if points is None and mask is None:
            logger.error('Must provide either points or mask.')
            raise ValueError
        cache = self.cache                            # one snapshot for all the tests below
        if self.caching and mask is not None:
            if cache['tilecode'] == tilecode:
                if surface in cache:
                    return cache[surface][mask]
                else:
                    logger.debug(f'Surface {surface} not in cache for tile {tilecode}.')
        elif self.caching:
            logger.debug('Caching enabled but no mask provided.')
        if cache['tilecode'] != tilecode and points is None:
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
        Enable caching of ahn tiles and interpolation data.
    """

    NAME = 'npz'

    def __init__(self, data_folder, caching=True):
        super().__init__(data_folder, caching)

    def _load_tile(self, tilecode):
        return load_ahn_tile(os.path.join(self.path, 'ahn_' + tilecode + '.npz'))


class ArrayAHNReader(AHNReader):
    """In-memory AHN reader (synthetic tiles, tests): ``tiles`` maps a tile code to a dict
    with keys 'x', 'y', 'ground_surface', 'building_surface'."""

    NAME = 'npz'

    def __init__(self, tiles, caching=True):
        self.tiles = tiles
        self.path = Path('.')
        self._init_state()
        self.set_caching(caching)

    def _load_tile(self, tilecode):
        if tilecode not in self.tiles:
            raise AHNFileNotFoundError(f'No AHN data for tile {tilecode}.')
        return self.tiles[tilecode]


_FOREIGN_LOCK = threading.Lock()


def device_raster_for(ahn_reader, tilecode, surface):
    """DeviceRaster for ANY reader object honouring the reference reader protocol
    (``filter_tile(tilecode)`` returning the tile dict) -- including the reference's
    own ``upcp.utils.ahn_utils.NPZReader`` instances, whose single-slot tile cache is not
    thread-safe: their ``filter_tile`` is called under a lock, and what it returns is kept
    per tilecode on this side."""
    if hasattr(ahn_reader, 'device_raster'):
        return ahn_reader.device_raster(tilecode, surface)
    with _FOREIGN_LOCK:
        state = getattr(ahn_reader, '_upcp_b200_state', None)
        if state is None:
            state = (threading.Lock(), _LRU(_TILE_SLOTS), _LRU(_RASTER_SLOTS))
            try:
                ahn_reader._upcp_b200_state = state
            except AttributeError:
                pass
    lock, tiles, rasters = state
    key = (tilecode, surface, str(dev.current_device()), dev.stream_key())
    with lock:
        raster = rasters.fetch(key)
        if raster is not None:
            return raster
        ahn_tile = tiles.fetch(tilecode)
        if ahn_tile is None:
            ahn_tile = ahn_reader.filter_tile(tilecode)      # serialised: the foreign cache has ONE tile slot
            tiles.store(tilecode, ahn_tile)
    if surface not in ahn_tile:
        logger.error(f'Unknown surface: {surface}.')
        raise ValueError
    raster = dev.DeviceRaster(ahn_tile['x'], ahn_tile['y'], ahn_tile[surface])
    with lock:
        rasters.store(key, raster)
    return raster
