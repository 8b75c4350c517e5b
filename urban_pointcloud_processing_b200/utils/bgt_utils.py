"""BGT polygon reader for the labelling path (drop-in for the ★ part of
``upcp.utils.bgt_utils``: BGTReader :19-94, BGTPolyReader :129-166, get_polygons
:200-203).  Host side: CSV -> pandas -> per-tile bbox query -> rings.

The reference builds shapely polygons (``Polygon(poly).buffer(offset)``, optional
``unary_union``, bgt_utils.py:154-163).  shapely/GEOS is an optional dependency here:
when importable it is used exactly as the reference does; otherwise polygons are returned as
``RingPolygon`` objects exposing the same ``.exterior.coords`` / ``.interiors`` protocol that
``poly_clip`` reads (clip_utils.py:211-212), and a buffer is carried as the polygon's ``offset``
attribute: the point-in-polygon kernel evaluates it as a distance (``upcp_pip_plan_create_buffered``:
the exact Minkowski sum where GEOS cuts round joins into 8 segments per quarter circle; unpinned
against GEOS).  The OR of per-polygon masks equals the mask of the union up to
measure-zero boundary sets, so ``merge=True`` needs no polygon union."""
import ast
import logging
import os
from abc import ABC, abstractmethod
from pathlib import Path

import numpy as np
import pandas as pd

from .las_utils import get_bbox_from_tile_code

logger = logging.getLogger(__name__)

try:  # optional
    from shapely.geometry import Polygon as _ShapelyPolygon
    from shapely.ops import unary_union as _unary_union
    HAVE_SHAPELY = True
except Exception:  # pragma: no cover - shapely absent in this image
    HAVE_SHAPELY = False


class _Ring:
    def __init__(self, coords):
        self.xy = np.ascontiguousarray(np.asarray(coords, dtype=np.float64).reshape(-1, 2))   # what the device plan reads
        self.coords = [tuple(c) for c in self.xy]                                             # shapely's view of the ring


class RingPolygon:
    """Minimal polygon: closed exterior ring + optional interior rings."""

    def __init__(self, shell, holes=(), offset=0.0):
        self.offset = float(offset)                    # pending buffer distance (evaluated on the device)
        shell = np.asarray(shell, dtype=np.float64).reshape(-1, 2)
        if len(shell) >= 1 and not np.array_equal(shell[0], shell[-1]):
            shell = np.vstack([shell, shell[:1]])      # shapely closes rings implicitly
        self.exterior = _Ring(shell)
        self.interiors = []
        for h in holes:
            h = np.asarray(h, dtype=np.float64).reshape(-1, 2)
            if len(h) >= 1 and not np.array_equal(h[0], h[-1]):
                h = np.vstack([h, h[:1]])
            self.interiors.append(_Ring(h))

    def buffer(self, offset):
        """``shapely.Polygon.buffer``: the same rings with the distance added to the pending offset (the
        point-in-polygon kernel evaluates it; nothing is re-vertexed on the host)."""
        if offset == 0:
            return self
        out = RingPolygon.__new__(RingPolygon)
        out.exterior, out.interiors, out.offset = self.exterior, self.interiors, self.offset + float(offset)
        return out


class BGTReader(ABC):
    """Abstract class for reading BGT data from a CSV file or a folder of CSV files."""

    COLUMNS = None

    def __init__(self, bgt_file=None, bgt_folder=None, file_prefix=''):
        if (bgt_file is None) and (bgt_folder is None):
            print("Provide either a bgt_file or bgt_folder to load.")
            return None
        if (bgt_file is not None) and (bgt_folder is not None):
            print("Provide either a bgt_file or bgt_folder to load, not both")
            return None
        if (bgt_folder is not None) and (not os.path.isdir(bgt_folder)):
            print('The data folder specified does not exist')
            return None
        if (bgt_file is not None) and (not os.path.isfile(bgt_file)):
            print('The data file specified does not exist')
            return None
        super().__init__()
        self.file_prefix = file_prefix
        self.bgt_df = pd.DataFrame(columns=type(self).COLUMNS)
        if bgt_file is not None:
            self.bgt_df = pd.read_csv(Path(bgt_file), header=0, names=type(self).COLUMNS)
        else:
            frames = [pd.read_csv(file, header=0, names=type(self).COLUMNS)
                      for file in Path(bgt_folder).glob(self.file_prefix + '*.csv')]
            if len(frames) == 0:
                logger.error(f'No data files found in {Path(bgt_folder).as_posix()}.')
                return
            self.bgt_df = pd.concat(frames)

    @abstractmethod
    def filter_tile(self, tilecode):
        return NotImplementedError


class BGTPolyReader(BGTReader):
    """CSV columns: [bgt_type, polygon, x_min, y_max, x_max, y_min]."""

    COLUMNS = ['bgt_type', 'polygon', 'x_min', 'y_max', 'x_max', 'y_min']

    def filter_tile(self, tilecode, bgt_types=[], exclude_types=[], padding=0, offset=0, merge=False):
        """List of polygons overlapping the area of the given tile code."""
        ((bx_min, by_max), (bx_max, by_min)) = get_bbox_from_tile_code(tilecode, padding=padding)
        df = self.bgt_df
        sel = ((df.x_min < bx_max) & (df.x_max > bx_min) & (df.y_min < by_max) & (df.y_max > by_min))
        if len(bgt_types) >= 1:
            sel &= df.bgt_type.isin(list(bgt_types))
        if len(exclude_types) >= 1:
            sel &= ~df.bgt_type.isin(list(exclude_types))
        polygons = [ast.literal_eval(poly) for poly in df.polygon.values[sel.values]]
        if HAVE_SHAPELY:
            if len(polygons) > 1 and merge:
                union = _unary_union([_ShapelyPolygon(poly).buffer(offset) for poly in polygons])
                poly_offset = [union] if type(union) == _ShapelyPolygon else list(union.geoms)
            else:
                poly_offset = [_ShapelyPolygon(poly).buffer(offset) for poly in polygons]
        else:
            poly_offset = [RingPolygon(poly).buffer(offset) for poly in polygons]
        return [poly for poly in poly_offset if len(poly.exterior.coords) > 1]


class BGTPointReader(BGTReader):
    """CSV columns: [bgt_type, x, y] (drop-in for bgt_utils.py:97-126; read by
    BGTStreetFurnitureFuser)."""

    COLUMNS = ['bgt_type', 'x', 'y']

    def filter_tile(self, tilecode, bgt_types=[], exclude_types=[], padding=0, return_types=False):
        """Points of the objects inside the (padded, inclusive) area of the given tile code."""
        ((bx_min, by_max), (bx_max, by_min)) = get_bbox_from_tile_code(tilecode, padding=padding)
        df = self.bgt_df
        sel = (df.x <= bx_max) & (df.x >= bx_min) & (df.y <= by_max) & (df.y >= by_min)
        if len(bgt_types) >= 1:
            sel &= df.bgt_type.isin(list(bgt_types))
        if len(exclude_types) >= 1:
            sel &= ~df.bgt_type.isin(list(exclude_types))
        records = list(df[sel.values].to_records(index=False))
        if return_types:
            return records
        return [(x, y) for (_, x, y) in records]


class ArrayBGTPointReader:
    """In-memory point reader (synthetic tiles, tests): ``tiles`` maps a tile code to a list of
    (bgt_type, x, y)."""

    def __init__(self, tiles):
        self.tiles = tiles

    def filter_tile(self, tilecode, bgt_types=[], exclude_types=[], padding=0, return_types=False):
        out = []
        for t, x, y in self.tiles.get(tilecode, []):
            if len(bgt_types) >= 1 and t not in bgt_types:
                continue
            if len(exclude_types) >= 1 and t in exclude_types:
                continue
            out.append((t, x, y) if return_types else (x, y))
        return out


class ArrayBGTReader:
    """In-memory polygon reader (synthetic tiles, tests): ``tiles`` maps a tile code to a
    list of polygons (RingPolygon / shapely-like / raw rings); ``types`` optionally gives
    one bgt_type per polygon."""

    def __init__(self, tiles, types=None):
        self.tiles = tiles
        self.types = types or {}

    def filter_tile(self, tilecode, bgt_types=[], exclude_types=[], padding=0, offset=0, merge=False):
        polys = self.tiles.get(tilecode, [])
        types = self.types.get(tilecode)
        out = []
        for i, p in enumerate(polys):
            t = types[i] if types is not None else None
            if t is not None and len(bgt_types) >= 1 and t not in bgt_types:
                continue
            if t is not None and len(exclude_types) >= 1 and t in exclude_types:
                continue
            if not hasattr(p, 'exterior'):
                p = RingPolygon(p)
            out.append(p.buffer(offset) if offset != 0 else p)
        return [p for p in out if len(p.exterior.coords) > 1]


def get_polygons(bgt_file, tilecode, padding=0, offset=0, merge=False):
    """Polygons from a bgt_file for a specific tilecode."""
    return BGTPolyReader(bgt_file=bgt_file).filter_tile(tilecode, padding=padding, offset=offset, merge=merge)


def get_points(bgt_file, tilecode, padding=0, return_types=True):
    """Point objects from a bgt_file for a specific tilecode."""
    return BGTPointReader(bgt_file=bgt_file).filter_tile(tilecode, padding=padding, return_types=return_types)
