"""Import the REAL reference (``/root/reference/src/upcp``) in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (never on the
GPU box).  The reference's own Python is imported UNCHANGED; stand-in modules are put in
``sys.modules`` only for third-party natives that are not installed here:

    shapely            duck-typed Polygon (buffer(0) is the identity, unary_union keeps
                       the list) -- enough for offset=0 / merge=True call sites; area /
                       intersects / intersection(...).area for a CONVEX self (CarFuser's
                       bounding rectangle) by half-plane clipping
    pycc, cccorelib    ccPointCloud + labelConnectedComponents backed by oracle.natives
    open3d             PointCloud / KDTreeFlann / estimate_normals backed by oracle.natives
    laspy, tifffile, zarr, pyntcloud      empty stubs (never called on this path)

numba, numpy, scipy and pandas are the real packages, so ``clip_utils`` (numba PIP),
``FastGridInterpolator``, ``get_octree_level``, ``vector_angle`` and all processor control
flow run as the reference wrote them.  Used by ``oracle/make_golden.py`` to pin the
restatement in ``oracle/reference_py.py`` + ``oracle/natives.py``.
"""
import os
import sys
import types

import numpy as np

from . import natives

REFERENCE_SRC = '/root/reference/src'


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'upcp'))


# ----------------------------------------------------------------- shapely ---------
class _Ring:
    def __init__(self, coords):
        self.coords = [tuple(c) for c in coords]


class _Polygon:
    def __init__(self, shell=None, holes=None):
        shell = list(shell) if shell is not None else []
        if len(shell) and tuple(shell[0]) != tuple(shell[-1]):
            shell = shell + [shell[0]]
        self.exterior = _Ring(shell)
        self.interiors = [_Ring(h) for h in (holes or [])]

    def buffer(self, offset, *a, **k):
        if offset != 0:
            raise NotImplementedError('shapely stand-in: buffer(offset != 0)')
        return self

    # car_fuser.py:73-79 only: self is the (convex) minimum bounding rectangle of a cluster
    @property
    def area(self):
        from . import reference_py
        return abs(reference_py._shoelace([tuple(c) for c in self.exterior.coords[:-1]]))

    def intersects(self, other):
        return self.intersection(other).area > 0

    def intersection(self, other):
        from . import reference_py
        perc = reference_py.rect_overlap_percentage(np.asarray(self.exterior.coords), other)
        return types.SimpleNamespace(area=perc / 100.0 * self.area)


class _Multi:
    def __init__(self, geoms):
        self.geoms = list(geoms)


def _unary_union(polys):
    return _Multi(polys)


# --------------------------------------------------------------- pycc / cccorelib ---
class _ScalarField:
    def __init__(self, n):
        self.values = np.zeros(n, dtype=np.float32)

    def asArray(self):
        return self.values


class _ccPointCloud:
    def __init__(self, xs, ys, zs):
        self.xs = np.asarray(xs, dtype=np.float32)
        self.ys = np.asarray(ys, dtype=np.float32)
        self.zs = np.asarray(zs, dtype=np.float32)
        self.fields = {}
        self.names = []
        self.current = -1

    def getScalarFieldIndexByName(self, name):
        return self.names.index(name) if name in self.names else -1

    def addScalarField(self, name):
        self.names.append(name)
        self.fields[len(self.names) - 1] = _ScalarField(len(self.xs))
        return len(self.names) - 1

    def setCurrentScalarField(self, idx):
        self.current = idx

    def getScalarField(self, idx):
        return self.fields[idx]


class _AutoSegmentationTools:
    @staticmethod
    def labelConnectedComponents(cloud, level=0, sixConnexity=False):
        labels = natives.label_connected_components(cloud.xs, cloud.ys, cloud.zs, int(level))
        cloud.fields[cloud.current].values = labels
        return int(labels.max()) if len(labels) else 0


# ------------------------------------------------------------------ open3d ---------
class _IndexedPoint(np.ndarray):
    """A coordinate triple that remembers which cloud index it came from."""
    _idx = -1


class _VecList:
    def __init__(self, arr):
        self.arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float64))

    def __len__(self):
        return len(self.arr)

    def __getitem__(self, i):
        p = self.arr[i].view(_IndexedPoint)
        p._idx = int(i)
        return p

    def __array__(self, dtype=None, copy=None):
        return self.arr


class _HybridParam:
    def __init__(self, radius, max_nn):
        self.radius = radius
        self.max_nn = max_nn


class _PointCloud:
    def __init__(self, points=None):
        self._points = _VecList(points.arr if isinstance(points, _VecList) else np.zeros((0, 3))) \
            if points is not None else _VecList(np.zeros((0, 3)))
        self.normals = None

    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, v):
        self._points = v if isinstance(v, _VecList) else _VecList(v)

    def estimate_normals(self, search_param=None):
        self.normals = _VecList(natives.normals_hybrid(self._points.arr, search_param.radius,
                                                       search_param.max_nn))

    def compute_mean_and_covariance(self):
        pts = self._points.arr
        ids = np.arange(len(pts), dtype=np.int32)[None, :]
        # covariance of ALL points of this (small) cloud in their given order
        cov6 = natives.cov_knn(pts, np.repeat(ids, len(pts), axis=0))[0]
        mean = pts.mean(axis=0)
        return mean, natives.cov6_to_matrix(cov6)


class _KDTreeFlann:
    def __init__(self, pcd):
        self.pts = pcd.points.arr
        self._knn = {}

    def search_knn_vector_3d(self, query, k):
        idx = getattr(query, '_idx', -1)
        if idx < 0:
            raise NotImplementedError('open3d stand-in: query must be a cloud point')
        if k not in self._knn:
            self._knn[k] = natives.knn(self.pts, k, want_d2=True)
        rows, d2 = self._knn[k]
        row = rows[idx]
        n = int(np.count_nonzero(row >= 0))
        return n, list(row[:n]), list(d2[idx][:n])

    def search_radius_vector_3d(self, query, radius):
        """KDTreeFlann.search_radius_vector_3d: every point with squared distance < radius^2 (nanoflann
        RadiusResultSet), ascending; ties by index.  Candidates from a scipy cKDTree ball query with a little
        slack, the decision and the order from the exact squared distance ((dx*dx + dy*dy) + dz*dz)."""
        from scipy.spatial import cKDTree
        if not hasattr(self, '_tree'):
            self._tree = cKDTree(self.pts)
        q = np.asarray(query, dtype=np.float64).reshape(3)
        cand = np.asarray(self._tree.query_ball_point(q, radius * (1.0 + 1e-9) + 1e-12), dtype=np.int64)
        d = self.pts[cand] - q
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 < radius * radius
        cand, d2 = cand[keep], d2[keep]
        order = np.lexsort((cand, d2))
        return len(order), list(cand[order]), list(d2[order])


def _make_modules():
    mods = {}
    shapely = types.ModuleType('shapely')
    geometry = types.ModuleType('shapely.geometry')
    ops = types.ModuleType('shapely.ops')
    geometry.Polygon = _Polygon
    geometry.LineString = type('LineString', (), {'__init__': lambda self, *a, **k: None})
    geometry.Point = type('Point', (), {'__init__': lambda self, *a, **k: None})
    geometry.MultiPolygon = _Multi
    geometry.MultiPoint = type('MultiPoint', (), {})
    geometry.mapping = lambda *a, **k: None
    ops.unary_union = _unary_union
    ops.polygonize = lambda *a, **k: []
    ops.cascaded_union = _unary_union
    shapely.geometry = geometry
    shapely.ops = ops
    mods.update({'shapely': shapely, 'shapely.geometry': geometry, 'shapely.ops': ops})

    pycc = types.ModuleType('pycc')
    pycc.PointCoordinateType = np.float32
    pycc.ccPointCloud = _ccPointCloud
    cccorelib = types.ModuleType('cccorelib')
    cccorelib.AutoSegmentationTools = _AutoSegmentationTools
    mods.update({'pycc': pycc, 'cccorelib': cccorelib})

    o3d = types.ModuleType('open3d')
    o3d.geometry = types.SimpleNamespace(PointCloud=_PointCloud, KDTreeFlann=_KDTreeFlann,
                                         KDTreeSearchParamHybrid=_HybridParam)
    o3d.utility = types.SimpleNamespace(Vector3dVector=_VecList)
    mods['open3d'] = o3d

    for name in ('laspy', 'tifffile', 'zarr', 'pyntcloud'):
        m = types.ModuleType(name)
        mods[name] = m
    mods['tifffile'].TiffFile = object
    mods['tifffile'].imread = lambda *a, **k: None
    mods['pyntcloud'].PyntCloud = object
    return mods


_loaded = None


def load_reference(numba_cache_dir='/tmp/upcp_numba_cache'):
    """Import and return the reference package ``upcp`` (real code, stand-in natives)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('/root/reference is not available on this machine')
    os.makedirs(numba_cache_dir, exist_ok=True)
    os.environ.setdefault('NUMBA_CACHE_DIR', numba_cache_dir)
    for name, mod in _make_modules().items():
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mod
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import upcp  # noqa: F401
    import upcp.pipeline
    import upcp.utils.clip_utils
    import upcp.utils.interpolation
    import upcp.utils.math_utils
    import upcp.utils.ahn_utils
    import upcp.utils.bgt_utils
    import upcp.fusion
    import upcp.region_growing
    import upcp.region_growing.region_growing
    _loaded = upcp
    return upcp


Polygon = _Polygon
