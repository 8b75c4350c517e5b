"""ctypes wrappers around the plain-C oracle natives (TEST INFRASTRUCTURE ONLY)."""
import ctypes
from ctypes import c_double, c_int, c_int64, c_void_p

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.orc_label_cc.restype = c_int64
        _lib.orc_region_grow.restype = c_int64
        _lib.orc_knn.restype = c_int
        _lib.orc_normals_hybrid.restype = c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p)


def set_num_threads(n):
    """OpenMP threads of the natives (returns the number in force)."""
    lib().orc_set_num_threads.restype = c_int
    return int(lib().orc_set_num_threads(c_int(int(n))))


def is_inside(x, y, polygon):
    """numba clip_utils.is_inside (clip_utils.py:167-190)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    poly = np.ascontiguousarray(np.asarray(polygon, dtype=np.float64)[:, :2])
    out = np.zeros(len(x), dtype=np.uint8)
    lib().orc_is_inside(_p(x), _p(y), c_int64(len(x)), _p(poly), c_int(len(poly)), _p(out))
    return out.view(bool)


def octree_params(xs, ys, zs, max_level=21):
    xs = np.ascontiguousarray(xs, dtype=np.float32)
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    zs = np.ascontiguousarray(zs, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    lib().orc_octree_params(_p(xs), _p(ys), _p(zs), c_int64(len(xs)), c_int(max_level), _p(out))
    return out


def label_connected_components(xs, ys, zs, level, max_level=21):
    """cccorelib.AutoSegmentationTools.labelConnectedComponents (label_connected_comp.py:83-85):
    float32 coordinates in, component label (1..K) per point out, as float32 scalar field."""
    xs = np.ascontiguousarray(xs, dtype=np.float32)
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    zs = np.ascontiguousarray(zs, dtype=np.float32)
    out = np.zeros(len(xs), dtype=np.int32)
    k = lib().orc_label_cc(_p(xs), _p(ys), _p(zs), c_int64(len(xs)), c_int(int(level)), c_int(max_level),
                           _p(out))
    if k < 0:
        raise MemoryError('orc_label_cc')
    return out.astype(np.float32)


def knn(points, k, want_d2=False):
    """Exact kNN rows in ascending (squared distance, index) order; -1 padding."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m = len(pts)
    idx = np.full((m, k), -1, dtype=np.int32)
    d2 = np.full((m, k), np.inf, dtype=np.float64) if want_d2 else None
    if lib().orc_knn(_p(pts), c_int64(m), c_int(k), _p(idx), _p(d2) if want_d2 else None) != 0:
        raise MemoryError('orc_knn')
    return (idx, d2) if want_d2 else idx


def normals_hybrid(points, radius, max_nn):
    """open3d estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) (region_growing.py:89-92)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros((len(pts), 3), dtype=np.float64)
    if lib().orc_normals_hybrid(_p(pts), c_int64(len(pts)), c_double(radius), c_int(max_nn), _p(out)) != 0:
        raise MemoryError('orc_normals_hybrid')
    return out


def cov_knn(points, knn_idx):
    """compute_mean_and_covariance of every kNN row (region_growing.py:69-71) -> (m,6)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    out = np.zeros((len(pts), 6), dtype=np.float64)
    lib().orc_cov_knn(_p(pts), c_int64(len(pts)), _p(knn_idx), c_int(knn_idx.shape[1]), _p(out))
    return out


def cov6_to_matrix(cov6):
    c = np.asarray(cov6)
    m = np.empty(c.shape[:-1] + (3, 3))
    m[..., 0, 0] = c[..., 0]; m[..., 0, 1] = c[..., 1]; m[..., 0, 2] = c[..., 2]
    m[..., 1, 0] = c[..., 1]; m[..., 1, 1] = c[..., 3]; m[..., 1, 2] = c[..., 4]
    m[..., 2, 0] = c[..., 2]; m[..., 2, 1] = c[..., 4]; m[..., 2, 2] = c[..., 5]
    return m


def fast_eigen(cov6):
    cov6 = np.ascontiguousarray(cov6, dtype=np.float64).reshape(-1, 6)
    out = np.zeros((len(cov6), 3))
    lib().orc_fast_eigen(_p(cov6), c_int64(len(cov6)), _p(out))
    return out


def region_grow(knn_idx, normals, seed, can_seed, threshold_angle, angle_tol=0.0):
    """RegionGrowing._region_growing queue loop (region_growing.py:94-139).  Returns
    (region bool[m], near bool[m]) where near marks points with an angle test within
    angle_tol of the threshold."""
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    normals = np.ascontiguousarray(normals, dtype=np.float64)
    seed = np.ascontiguousarray(seed, dtype=np.uint8)
    can_seed = np.ascontiguousarray(can_seed, dtype=np.uint8)
    m = len(knn_idx)
    region = np.zeros(m, dtype=np.uint8)
    near = np.zeros(m, dtype=np.uint8)
    r = lib().orc_region_grow(_p(knn_idx), c_int(knn_idx.shape[1]), _p(normals), _p(seed), _p(can_seed),
                              c_int64(m), c_double(threshold_angle), c_double(angle_tol), _p(region),
                              _p(near))
    if r < 0:
        raise MemoryError('orc_region_grow')
    return region.view(bool), near.view(bool)
