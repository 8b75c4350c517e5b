"""Normal-based region growing (drop-in for
``upcp.region_growing.region_growing.RegionGrowing``, reference
src/upcp/region_growing/region_growing.py:15-170).

The reference builds an Open3D KD-tree, estimates normals with a hybrid search, and grows
the region with a Python ``while`` loop over a list of seeds.  Here: one exact
kNN + normals + curvature-covariance pass over a uniform-grid cell list
(``upcp_knn_normals``) and one persistent work-queue kernel that expands the frontier to
the least fixpoint (``upcp_region_grow``) -- which is what the Python loop computes,
independently of the order in which seeds are visited."""
import ctypes
import logging
import os

import numpy as np
import torch

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor, PerThread
from ..labels import Labels

logger = logging.getLogger(__name__)

CURVATURE_TOL = 1e-9


def default_cell_size(bbox_sel, m, radius):
    """Fine-cell edge of the search grid: aim at ~16 points per occupied cell of a surface-like
    cloud (urban LiDAR covers roughly 3x its footprint area), bounded by the search radius.  Near
    the upper bound the cell snaps to 0.525 x radius: a 5^3-cell block then always contains the
    whole normal-estimation ball, which lets the second block tier certify sparse queries."""
    ex = max(float(bbox_sel[3] - bbox_sel[0]), 1e-3)
    ey = max(float(bbox_sel[4] - bbox_sel[1]), 1e-3)
    c = float(np.sqrt(48.0 * ex * ey / max(int(m), 1)))
    r = max(float(radius), 1e-3)
    c = min(max(c, r / 8.0), 0.525 * r)
    return 0.525 * r if c > 0.45 * r else c


def knn_normals_device(tile, d_sel, m, bbox_sel, knn, max_nn, radius, cell_size=None,
                       want_d2=False, want_cov=True, want_curv=True, threshold_curve=None):
    """kNN rows (compact ids), normals, k-neighbourhood covariance and analytic eigenvalue
    ratios for the selected points of ``tile``."""
    lib = _lib.load()
    device = tile.device
    if cell_size is None:
        cell_size = default_cell_size(bbox_sel, m, radius)
    d_knn = torch.empty((m, knn), dtype=torch.int32, device=device)
    d_d2 = torch.empty((m, knn), dtype=torch.float64, device=device) if want_d2 else None
    d_normals = torch.empty((m, 3), dtype=torch.float64, device=device)
    d_cov = torch.empty((m, 6), dtype=torch.float64, device=device) if want_cov else None
    d_curv = torch.empty((m, 3), dtype=torch.float64, device=device) if want_curv else None
    d_flag = torch.empty(m, dtype=torch.uint8, device=device) if threshold_curve is not None else None
    ws = dev.workspace(device).get(lib.upcp_knn_ws_bytes(tile.n, m, max(knn, max_nn)))
    bbox_arr = (ctypes.c_double * 6)(*[float(v) for v in bbox_sel])
    _lib.check(lib.upcp_knn_normals(dev.ptr(tile.x), dev.ptr(tile.y), dev.ptr(tile.z), tile.n,
                                    dev.ptr(d_sel), m, bbox_arr, int(knn), int(max_nn), float(radius),
                                    float(cell_size), dev.ptr(d_knn), dev.ptr(d_d2), dev.ptr(d_normals),
                                    dev.ptr(d_cov), dev.ptr(d_curv),
                                    float(threshold_curve) if threshold_curve is not None else 0.0,
                                    dev.ptr(d_flag), dev.ptr(ws), ws.numel(), dev.stream_ptr()),
               'upcp_knn_normals')
    if threshold_curve is not None:
        return d_knn, d_d2, d_normals, d_cov, d_curv, d_flag
    return d_knn, d_d2, d_normals, d_cov, d_curv


def knn_normals_sorted_device(tile, d_sel, m, bbox_sel, knn, max_nn, radius, threshold_curve, cell_size=None):
    """The same search with all outputs in the library's spatial order (``upcp_knn_normals_sorted``):
    returns (d_knn, d_normals, d_cov, d_flag, d_order) where row s / neighbour ids are sorted
    positions and ``d_order[s]`` is the original point index of sorted position s."""
    lib = _lib.load()
    device = tile.device
    if cell_size is None:
        cell_size = default_cell_size(bbox_sel, m, radius)
    d_knn = torch.empty((m, knn), dtype=torch.int32, device=device)
    d_normals = torch.empty((m, 3), dtype=torch.float64, device=device)
    d_cov = torch.empty((m, 6), dtype=torch.float64, device=device)
    d_flag = torch.empty(m, dtype=torch.uint8, device=device)
    d_order = torch.empty(m, dtype=torch.int32, device=device)
    ws = dev.workspace(device).get(lib.upcp_knn_ws_bytes(tile.n, m, max(knn, max_nn)))
    bbox_arr = (ctypes.c_double * 6)(*[float(v) for v in bbox_sel])
    _lib.check(lib.upcp_knn_normals_sorted(dev.ptr(tile.x), dev.ptr(tile.y), dev.ptr(tile.z), tile.n,
                                           dev.ptr(d_sel), m, bbox_arr, int(knn), int(max_nn), float(radius),
                                           float(cell_size), dev.ptr(d_knn), dev.ptr(d_normals), dev.ptr(d_cov),
                                           float(threshold_curve), dev.ptr(d_flag), dev.ptr(d_order),
                                           dev.ptr(ws), ws.numel(), dev.stream_ptr()),
               'upcp_knn_normals_sorted')
    return d_knn, d_normals, d_cov, d_flag, d_order


def mask_take(d_full, d_order):
    """out[s] = full[order[s]] (full-tile mask -> sorted space)."""
    lib = _lib.load()
    m = d_order.numel()
    out = torch.empty(m, dtype=torch.uint8, device=d_full.device)
    _lib.check(lib.upcp_mask_take(dev.ptr(d_full), dev.ptr(d_order), m, dev.ptr(out), dev.stream_ptr()),
               'upcp_mask_take')
    return out


def mask_put(d_sorted, d_order, n):
    """full[order[s]] = 1 where sorted[s] (sorted space -> full-tile mask, zero elsewhere)."""
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.uint8, device=d_order.device)
    _lib.check(lib.upcp_mask_put(dev.ptr(d_sorted), dev.ptr(d_order), d_order.numel(), n, dev.ptr(out),
                                 dev.stream_ptr()), 'upcp_mask_put')
    return out


def can_seed_from_curvature(d_cov, d_curv, threshold_curve):
    """uint8 cuda tensor [m]: ``eig_val[0] / eig_val.sum() < threshold_curve``
    (region_growing.py:59-73,132).

    ``eig_val[0]`` is whatever LAPACK ``dgeev`` returns first (unsorted), which no closed
    form predicts.  The device supplies the three analytic ratios; when all three fall on
    the same side of the threshold (the normal case for the default threshold 1.0) the
    decision does not depend on the order.  The remaining points are resolved on the host
    with ``np.linalg.eig`` exactly as the reference does.  Returns (d_can_seed, n_host).
    """
    r = d_curv
    finite = torch.isfinite(r).all(dim=1)
    below = (r < threshold_curve - CURVATURE_TOL).all(dim=1) & finite
    above = (r >= threshold_curve + CURVATURE_TOL).all(dim=1) & finite
    ambiguous = ~(below | above)
    can_seed = below.to(torch.uint8)
    idx = torch.nonzero(ambiguous).flatten()
    n_host = int(idx.numel())
    if n_host:
        c = d_cov[idx].cpu().numpy()
        cov = np.empty((n_host, 3, 3))
        cov[:, 0, 0] = c[:, 0]; cov[:, 0, 1] = c[:, 1]; cov[:, 0, 2] = c[:, 2]
        cov[:, 1, 0] = c[:, 1]; cov[:, 1, 1] = c[:, 3]; cov[:, 1, 2] = c[:, 4]
        cov[:, 2, 0] = c[:, 2]; cov[:, 2, 1] = c[:, 4]; cov[:, 2, 2] = c[:, 5]
        with np.errstate(all='ignore'):
            eig_val = np.linalg.eig(cov)[0]
            curvature = eig_val[:, 0] / eig_val.sum(axis=1)
            res = (curvature < threshold_curve)
        can_seed[idx] = torch.from_numpy(np.ascontiguousarray(res.astype(np.uint8))).to(can_seed.device)
    return can_seed, n_host


def _host_curvature(cov6, threshold_curve):
    cov = np.empty((len(cov6), 3, 3))
    cov[:, 0, 0] = cov6[:, 0]; cov[:, 0, 1] = cov6[:, 1]; cov[:, 0, 2] = cov6[:, 2]
    cov[:, 1, 0] = cov6[:, 1]; cov[:, 1, 1] = cov6[:, 3]; cov[:, 1, 2] = cov6[:, 4]
    cov[:, 2, 0] = cov6[:, 2]; cov[:, 2, 1] = cov6[:, 4]; cov[:, 2, 2] = cov6[:, 5]
    with np.errstate(all='ignore'):
        eig_val = np.linalg.eig(cov)[0]
        return (eig_val[:, 0] / eig_val.sum(axis=1)) < threshold_curve


def resolve_seed_flags(d_flag, d_cov, threshold_curve):
    """Turn the device seed flags (1 / 0 / 2 = undecided) into a 0 / 1 mask: the undecided
    points are resolved with ``np.linalg.eig`` (LAPACK dgeev, unsorted eigenvalues) exactly as
    the reference does (region_growing.py:69-73), in parallel over the host cores.  Returns
    (d_can_seed, number of host-resolved points)."""
    idx = torch.nonzero(d_flag == 2).flatten()
    n_host = int(idx.numel())
    if n_host == 0:
        return d_flag, 0
    c = d_cov[idx].cpu().numpy()
    n_thr = max(1, min(_host_threads(), n_host // 256))
    if n_thr > 1:
        parts = list(_pool().map(lambda a: _host_curvature(a, threshold_curve), np.array_split(c, n_thr)))
        res = np.concatenate(parts)
    else:
        res = _host_curvature(c, threshold_curve)
    d_flag[idx] = torch.from_numpy(np.ascontiguousarray(res.astype(np.uint8))).to(d_flag.device)
    return d_flag, n_host


def region_grow_device(d_knn, d_normals, d_seed, d_can_seed, threshold_angle):
    lib = _lib.load()
    m, knn = d_knn.shape
    device = d_knn.device
    d_region = torch.empty(m, dtype=torch.uint8, device=device)
    d_info = torch.zeros(2, dtype=torch.int64, device=device)
    ws = dev.workspace(device).get(lib.upcp_region_grow_ws_bytes(m))
    _lib.check(lib.upcp_region_grow(dev.ptr(d_knn), int(knn), dev.ptr(d_normals), dev.ptr(d_seed),
                                    dev.ptr(d_can_seed), m, float(threshold_angle), dev.ptr(d_region),
                                    dev.ptr(d_info), dev.ptr(ws), ws.numel(), dev.stream_ptr()),
               'upcp_region_grow')
    return d_region, d_info


_POOL = None


def _host_threads():
    """Host threads for the LAPACK resolution: the cores of this box divided by the ranks sharing it."""
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    return max(1, min((os.cpu_count() or 1) // ranks, 16))


def _pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(_host_threads())
    return _POOL


def region_grow_overlapped(d_knn, d_normals, d_seed, d_flag, d_cov, threshold_curve, threshold_angle):
    """Growth to the least fixpoint with the undecided curvature flags (value 2) resolved on the
    host WHILE the device grows (``upcp_region_grow_begin / _resume / _end``).

    The undecided covariances are fetched first; the device then starts growing from the flags it
    could decide itself while the host puts the fetched ones through ``np.linalg.eig`` exactly as
    the reference does (region_growing.py:69-73); the flags are then rewritten and the growth
    resumed from the points that turned out to be seeds.  Growth is monotone, so the result equals the one-shot fixpoint.
    Returns (d_region, d_info, n_host)."""
    lib = _lib.load()
    m, knn = d_knn.shape
    device = d_knn.device
    # the undecided covariances leave the device BEFORE the growth is queued: the persistent
    # kernel occupies every SM, nothing else would get through while it runs
    idx = torch.nonzero(d_flag == 2).flatten()
    n_host = int(idx.numel())
    c = d_cov[idx].cpu().numpy() if n_host else None
    d_region = torch.empty(m, dtype=torch.uint8, device=device)
    d_info = torch.zeros(2, dtype=torch.int64, device=device)
    ws = dev.workspace(device, slot=2).get(lib.upcp_region_grow_ws_bytes(m))
    _lib.check(lib.upcp_region_grow_begin(dev.ptr(d_knn), int(knn), dev.ptr(d_normals), dev.ptr(d_seed),
                                          dev.ptr(d_flag), m, float(threshold_angle), dev.ptr(ws), ws.numel(),
                                          dev.stream_ptr()), 'upcp_region_grow_begin')
    # ---- host resolution of the undecided flags, overlapped with the kernels queued above
    if n_host:
        n_thr = max(1, min(_host_threads(), n_host // 256))
        if n_thr > 1:
            parts = list(_pool().map(lambda a: _host_curvature(a, threshold_curve), np.array_split(c, n_thr)))
            res = np.concatenate(parts)
        else:
            res = _host_curvature(c, threshold_curve)
        d_flag[idx] = torch.from_numpy(np.ascontiguousarray(res.astype(np.uint8))).to(device)
        _lib.check(lib.upcp_region_grow_resume(dev.ptr(d_knn), int(knn), dev.ptr(d_normals), dev.ptr(d_seed),
                                               dev.ptr(d_flag), m, dev.ptr(idx), n_host, float(threshold_angle),
                                               dev.ptr(ws), ws.numel(), dev.stream_ptr()),
                   'upcp_region_grow_resume')
    _lib.check(lib.upcp_region_grow_end(m, dev.ptr(d_region), dev.ptr(d_info), dev.ptr(ws), ws.numel(),
                                        dev.stream_ptr()), 'upcp_region_grow_end')
    return d_region, d_info, n_host


def _exact_radius_curvature(q, nb, threshold_curve):
    """The reference's curvature test for ONE point from its radius neighbourhood (region_growing.py:59-73):
    neighbours in ascending (squared distance, compact id) order, Open3D's single-pass cumulant covariance of
    their raw coordinates in that order, LAPACK eig (unsorted), eig_val[0] / sum < threshold."""
    d = nb[:, :3] - q
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    p = nb[np.lexsort((nb[:, 3], d2)), :3]
    n = float(len(p))
    s1 = np.cumsum(p, axis=0)[-1] / n
    prod = np.stack([p[:, 0] * p[:, 0], p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 1], p[:, 1] * p[:, 2],
                     p[:, 2] * p[:, 2]], axis=1)
    s2 = np.cumsum(prod, axis=0)[-1] / n
    cov = np.array([[s2[0] - s1[0] * s1[0], s2[1] - s1[0] * s1[1], s2[2] - s1[0] * s1[2]],
                    [s2[1] - s1[0] * s1[1], s2[3] - s1[1] * s1[1], s2[4] - s1[1] * s1[2]],
                    [s2[2] - s1[0] * s1[2], s2[4] - s1[1] * s1[2], s2[5] - s1[2] * s1[2]]])
    with np.errstate(all='ignore'):
        eig_val = np.linalg.eig(cov)[0]
        return bool((eig_val[0] / eig_val.sum()) < threshold_curve)


def region_grow_radius_device(tile, d_sel, m, bbox_sel, d_normals, d_seed, radius, threshold_angle, threshold_curve):
    """Least fixpoint of the growth over fixed-radius neighbourhoods (``_region_growing('radius')``,
    region_growing.py:104-106) on the device; all arrays in compact-id space.  Curvature flags the device
    cannot decide (see csrc/radius.cu) are resolved on the host from the exact ordered neighbour lists -- only
    for points the growth actually reaches -- and the growth resumed until nothing is left undecided.
    Returns (d_region uint8 [m], stats dict)."""
    lib = _lib.load()
    device = tile.device
    bbox_arr = (ctypes.c_double * 6)(*[float(v) for v in bbox_sel])
    ws = dev.workspace(device, slot=3).get(lib.upcp_radius_ws_bytes(tile.n, m))
    d_flag = torch.empty(m, dtype=torch.uint8, device=device)
    _lib.check(lib.upcp_radius_prepare(dev.ptr(tile.x), dev.ptr(tile.y), dev.ptr(tile.z), tile.n, dev.ptr(d_sel), m, bbox_arr,
                                       float(radius), dev.ptr(d_normals), dev.ptr(d_seed), float(threshold_curve),
                                       dev.ptr(d_flag), None, dev.ptr(ws), ws.numel(), dev.stream_ptr()), 'upcp_radius_prepare')
    d_region = torch.empty(m, dtype=torch.uint8, device=device)
    d_info = torch.zeros(2, dtype=torch.int64, device=device)
    d_extra, n_host, rounds = None, 0, 0
    # compact id -> coordinates of the few undecided points (queries of the host resolution)
    d_ids = torch.nonzero(d_sel).flatten() if d_sel is not None else None
    while True:
        rounds += 1
        _lib.check(lib.upcp_radius_grow(tile.n, m, bbox_arr, float(radius), dev.ptr(d_flag), float(threshold_angle),
                                        dev.ptr(d_extra), 0 if d_extra is None else int(d_extra.numel()),
                                        dev.ptr(ws), ws.numel(), dev.stream_ptr()), 'upcp_radius_grow')
        _lib.check(lib.upcp_radius_region(tile.n, m, dev.ptr(d_region), dev.ptr(d_info), dev.ptr(ws), ws.numel(),
                                          dev.stream_ptr()), 'upcp_radius_region')
        d_und = torch.nonzero((d_flag == 2) & (d_region != 0)).flatten()
        nq = int(d_und.numel())
        if nq == 0:
            break
        n_host += nq
        d_off = torch.empty((nq, 2), dtype=torch.int64, device=device)
        d_tot = torch.zeros(1, dtype=torch.int64, device=device)
        cap = max(4096, 512 * nq)
        while True:
            d_nb = torch.empty((cap, 4), dtype=torch.float64, device=device)
            _lib.check(lib.upcp_radius_neighbours(tile.n, m, bbox_arr, float(radius), dev.ptr(d_und), nq, dev.ptr(d_off),
                                                  dev.ptr(d_nb), cap, dev.ptr(d_tot), dev.ptr(ws), ws.numel(),
                                                  dev.stream_ptr()), 'upcp_radius_neighbours')
            total = int(d_tot.item())
            if total <= cap:
                break
            cap = total
        off, nb = d_off.cpu().numpy(), d_nb[:total].cpu().numpy()
        orig = d_ids[d_und] if d_ids is not None else d_und
        q = torch.stack((tile.x[orig], tile.y[orig], tile.z[orig]), dim=1).cpu().numpy()
        res = np.array([_exact_radius_curvature(q[i], nb[off[i, 0]:off[i, 0] + off[i, 1]], threshold_curve) for i in range(nq)])
        d_flag[d_und] = torch.from_numpy(res.astype(np.uint8)).to(device)
        d_extra = d_und
        if not res.any():
            break
    return d_region, {'n_curvature_on_host': n_host, 'rounds': rounds, '_d_info': d_info}


def mask_gather(d_sel, d_full, m):
    lib = _lib.load()
    n = d_sel.numel()
    out = torch.empty(m, dtype=torch.uint8, device=d_sel.device)
    ws = dev.workspace(d_sel.device, slot=1).get(lib.upcp_compact_ws_bytes(n))
    _lib.check(lib.upcp_mask_gather(dev.ptr(d_sel), n, dev.ptr(d_full), dev.ptr(out), None, dev.ptr(ws),
                                    ws.numel(), dev.stream_ptr()), 'upcp_mask_gather')
    return out


def mask_scatter(d_sel, d_compact):
    lib = _lib.load()
    n = d_sel.numel()
    out = torch.empty(n, dtype=torch.uint8, device=d_sel.device)
    ws = dev.workspace(d_sel.device, slot=1).get(lib.upcp_compact_ws_bytes(n))
    _lib.check(lib.upcp_mask_scatter(dev.ptr(d_sel), n, dev.ptr(d_compact), dev.ptr(out), dev.ptr(ws),
                                     ws.numel(), dev.stream_ptr()), 'upcp_mask_scatter')
    return out


class RegionGrowing(AbstractProcessor):
    """
    Region growing based on the smoothness constraint of
    https://pcl.readthedocs.io/projects/tutorials/en/latest/region_growing_segmentation.html

    Parameters (names and defaults as the reference, region_growing.py:20-22)
    ----------
    label, exclude_labels=[], threshold_angle=20, threshold_curve=1.0, max_nn=30,
    grow_region_knn=15, grow_region_radius=0.2
    """

    ignores_mask = True        # _set_mask (region_growing.py:34-49) replaces the incoming mask

    last_stats = PerThread(dict)        # counters of the calling thread's last call

    def __init__(self, label, exclude_labels=[], threshold_angle=20, threshold_curve=1.0, max_nn=30,
                 grow_region_knn=15, grow_region_radius=0.2):
        super().__init__(label)
        self.threshold_angle = threshold_angle
        self.threshold_curve = threshold_curve
        self.max_nn = max_nn
        self.grow_region_knn = grow_region_knn
        self.grow_region_radius = grow_region_radius
        self.exclude_labels = exclude_labels
        self.cell_size = None           # search-grid cell size; None = radius / 2
        self.method = 'knn'             # 'radius': what the reference computes through _region_growing('radius')
        self.last_stats = {}
        if max(int(grow_region_knn), int(max_nn)) > 32:
            # (the reference accepts any value; the per-query tier of the search keeps one neighbour per lane)
            raise ValueError('RegionGrowing on the B200 extension supports grow_region_knn and max_nn up to 32 '
                             f'(got grow_region_knn={grow_region_knn}, max_nn={max_nn})')

    # -- the reference's internal protocol (get_labels = _set_mask + _convert_input_cloud + _region_growing,
    # region_growing.py:164-168), kept because method='radius' is only reachable through it
    def _set_mask(self, las_labels):
        self._las_labels = np.asarray(las_labels)
        mask = np.ones((len(self._las_labels),), dtype=bool)
        for exclude_label in self.exclude_labels:
            mask = mask & (self._las_labels != exclude_label)
        self.list_of_seed_ids = np.where(self._las_labels[mask] == self.label)[0].tolist()
        self.mask_indices = np.where(mask)[0]
        self.label_mask = np.zeros(len(mask), dtype=bool)
        self.mask = mask

    def _convert_input_cloud(self, las):
        self._points = las

    def _region_growing(self, method='knn'):
        """Label mask (numpy bool[N]) of the grown region, for the cloud / labels handed to
        ``_convert_input_cloud`` / ``_set_mask``."""
        saved, self.method = self.method, method
        try:
            tile = dev.get_tile(self._points)
            d_labels = dev.upload_labels(self._las_labels, tile.device)
            self.label_mask = dev.download_mask(self._label_mask_device(tile, d_labels, None))
        finally:
            self.method = saved
        return self.label_mask

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode=None):
        logger.info('KDTree based Region Growing ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        if tile.ncols != 3:
            raise ValueError('RegionGrowing needs 3D points')
        # _set_mask (region_growing.py:34-49): the incoming mask is ignored
        d_sel = dev.mask_build(d_labels, None, base=True, ne=tuple(self.exclude_labels))
        _, bbox_sel, m = tile.bbox(d_sel)
        d_label_mask = torch.zeros(tile.n, dtype=torch.uint8, device=tile.device)
        if m == 0:
            return d_label_mask
        # seeds: selected points already carrying the label (region_growing.py:80-84)
        d_seed_full = dev.mask_build(d_labels, None, base=False, eq=(self.label,), ne=tuple(self.exclude_labels))
        n_seeds = dev.mask_count(d_seed_full)
        if n_seeds == 0:
            logger.debug('Input point cloud does not contain any seed points.')
            return d_label_mask
        if self.method == 'radius':
            # normals from the hybrid search as always (region_growing.py:89-92); neighbourhoods of the growth and of
            # the curvature test = everything within grow_region_radius (:61-63,104-106)
            _, _, d_normals, _, _ = knn_normals_device(tile, d_sel, m, bbox_sel, 1, self.max_nn, self.grow_region_radius,
                                                       cell_size=self.cell_size, want_cov=False, want_curv=False)
            d_seed = mask_gather(d_sel, d_seed_full, m)
            d_region, stats = region_grow_radius_device(tile, d_sel, m, bbox_sel, d_normals, d_seed, self.grow_region_radius,
                                                        self.threshold_angle, self.threshold_curve)
            self.last_stats = dict(stats, n_selected=m, n_seeds=n_seeds)
            return mask_scatter(d_sel, d_region)
        if self.method != 'knn':
            raise ValueError(f"method must be 'knn' or 'radius' (got {self.method!r})")
        # everything below lives in the search grid's spatial order (close points, close rows)
        d_knn, d_normals, d_cov, d_flag, d_order = knn_normals_sorted_device(
            tile, d_sel, m, bbox_sel, self.grow_region_knn, self.max_nn, self.grow_region_radius,
            self.threshold_curve, cell_size=self.cell_size)
        d_seed = mask_take(d_seed_full, d_order)
        d_region, d_info, n_host = region_grow_overlapped(d_knn, d_normals, d_seed, d_flag, d_cov,
                                                          self.threshold_curve, self.threshold_angle)
        self.last_stats = {'n_selected': m, 'n_seeds': n_seeds, 'n_curvature_on_host': n_host,
                           '_d_info': d_info}
        return mask_put(d_region, d_order, tile.n)
