"""Per-cluster ("object") helpers shared by the clients of LabelConnectedComp.get_components:
CarFuser (reference fusion/car_fuser.py:45-89) and BGTStreetFurnitureFuser
(fusion/street_furniture_fuser.py:46-89).

The reference loops over the clusters in Python and rescans the masked cloud per cluster.
Here the device groups the points by cluster and reduces all clusters in one pass
(``upcp_component_stats``); the host only sees one small record per cluster, runs the
reference's per-object geometry (convex hull + minimum bounding rectangle, on the few clusters
whose height qualifies) and hands the accepted objects back to the device
(``upcp_prism_clip`` / ``upcp_component_mark``)."""
import numpy as np
import torch

from .. import _lib
from .. import device as dev
from . import bgt_utils

FIXED_SCALE = 16777216.0          # 2^24, see csrc/objects.cu


class ComponentStats:
    """Host view of the per-cluster reductions + device handles to the grouping."""

    def __init__(self, tile, d_comp, d_gz=None, cap=1 << 16):
        lib = _lib.load()
        n = tile.n
        device = tile.device
        self.tile = tile
        while True:
            self.d_dense = torch.empty(n, dtype=torch.int32, device=device)
            self.d_order = torch.empty(max(n, 1), dtype=torch.int32, device=device)
            d_seg = torch.zeros(cap + 1, dtype=torch.int64, device=device)
            d_cnt = torch.zeros(cap, dtype=torch.int64, device=device)
            d_sum = torch.zeros(cap, dtype=torch.int64, device=device)
            d_max = torch.zeros(cap, dtype=torch.float64, device=device)
            d_box = torch.zeros((cap, 4), dtype=torch.float64, device=device)
            d_info = torch.zeros(4, dtype=torch.int64, device=device)
            ws = dev.workspace(device).get(lib.upcp_component_stats_ws_bytes(n))
            status = lib.upcp_component_stats(dev.ptr(d_comp), n, dev.ptr(tile.x), dev.ptr(tile.y),
                                              dev.ptr(tile.z) if tile.ncols == 3 else None, dev.ptr(d_gz), cap,
                                              dev.ptr(self.d_dense), dev.ptr(self.d_order), dev.ptr(d_seg),
                                              dev.ptr(d_cnt), dev.ptr(d_sum), dev.ptr(d_max), dev.ptr(d_box),
                                              dev.ptr(d_info), dev.ptr(ws), ws.numel(), dev.stream_ptr())
            info = d_info.cpu().numpy()
            if status == -4 and int(info[0]) > cap:           # UPCP_E_CAPACITY: retry with room for all clusters
                cap = int(info[0])
                continue
            _lib.check(status, 'upcp_component_stats')
            break
        k = self.n_components = int(info[0])
        self.n_points = int(info[1])
        self.n_gz_out_of_range = int(info[3])
        self.seg_first = d_seg[:k + 1].cpu().numpy()
        self.size = np.diff(self.seg_first)
        self.count_finite = d_cnt[:k].cpu().numpy()
        sum_fixed = d_sum[:k].cpu().numpy()
        # mean of the finite ground elevations (np.mean(valid_values), car_fuser.py:65): the integer
        # sum is exact, one rounding in the division; NaN where no finite value exists
        with np.errstate(invalid='ignore', divide='ignore'):
            self.mean_gz = np.where(self.count_finite > 0,
                                    (sum_fixed.astype(np.float64) / FIXED_SCALE) / self.count_finite, np.nan)
        self.max_z = d_max[:k].cpu().numpy()
        self.bbox = d_box[:k].cpu().numpy()

    def xy_of(self, k):
        """(size, 2) xy of cluster k in original point order (== points[mask][cc_mask][:, :2])."""
        lib = _lib.load()
        a, b = int(self.seg_first[k]), int(self.seg_first[k + 1])
        out = torch.empty((b - a, 2), dtype=torch.float64, device=self.tile.device)
        _lib.check(lib.upcp_gather_xy(dev.ptr(self.tile.x), dev.ptr(self.tile.y),
                                      dev.ptr(self.d_order[a:b]) if b > a else None, b - a, dev.ptr(out),
                                      dev.stream_ptr()), 'upcp_gather_xy')
        return out.cpu().numpy()

    def mark(self, accepted):
        """uint8 cuda mask [N] of the points whose cluster is flagged in ``accepted`` (bool[K])."""
        lib = _lib.load()
        out = torch.zeros(self.tile.n, dtype=torch.uint8, device=self.tile.device)
        acc = np.ascontiguousarray(np.asarray(accepted, dtype=np.uint8))
        if self.tile.n == 0 or acc.size == 0:
            return out
        d_acc = torch.from_numpy(acc).to(self.tile.device)
        _lib.check(lib.upcp_component_mark(dev.ptr(self.d_dense), self.tile.n, dev.ptr(d_acc), int(acc.size),
                                           dev.ptr(out), dev.stream_ptr()), 'upcp_component_mark')
        return out


def height_candidates(stats, min_height, max_height):
    """Clusters with at least one finite ground value whose top lies in
    [cc_z + min_height, cc_z + max_height] (car_fuser.py:64-69).  Returns (ids, cc_z, max_z)."""
    with np.errstate(invalid='ignore'):
        cc_z = stats.mean_gz
        lo, hi = cc_z + min_height, cc_z + max_height
        ok = (stats.count_finite > 0) & (lo <= stats.max_z) & (stats.max_z <= hi)
    return np.where(ok)[0], cc_z, hi


def prism_clip_device(tile, d_sel, rings, zranges):
    """uint8 cuda mask [N]: OR over i of poly_box_clip(points, rings[i], bottom, top)
    (clip_utils.py:240-262) restricted to ``d_sel``; rings are closed (k, 2) arrays."""
    lib = _lib.load()
    out = torch.zeros(tile.n, dtype=torch.uint8, device=tile.device)
    if len(rings) == 0 or tile.n == 0:
        return out
    xy = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 2) for r in rings]))
    off = np.zeros(len(rings) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(r) for r in rings])
    zr = np.ascontiguousarray(np.asarray(zranges, dtype=np.float64).reshape(-1, 2))
    d_xy = torch.from_numpy(xy).to(tile.device)
    d_off = torch.from_numpy(off).to(tile.device)
    d_zr = torch.from_numpy(zr).to(tile.device)
    _lib.check(lib.upcp_prism_clip(dev.ptr(tile.x), dev.ptr(tile.y), dev.ptr(tile.z), tile.n, dev.ptr(d_sel),
                                   dev.ptr(d_xy), dev.ptr(d_off), dev.ptr(d_zr), len(rings), dev.ptr(out),
                                   dev.stream_ptr()), 'upcp_prism_clip')
    return out


# ---------------------------------------------------------------- rectangle / polygon overlap ---
def _ring_area(ring):
    x, y = ring[:, 0], ring[:, 1]
    return 0.5 * float(np.sum(x[:-1] * y[1:] - x[1:] * y[:-1])) if len(ring) >= 3 else 0.0


def _clip_ring_by_convex(ring, clip):
    """Sutherland-Hodgman: the part of the (possibly concave) closed ``ring`` inside the convex,
    counter-clockwise closed ring ``clip``; returns a closed ring (possibly with zero-area spikes,
    which do not change the shoelace area)."""
    out = [tuple(p) for p in ring[:-1]]
    for a, b in zip(clip[:-1], clip[1:]):
        if not out:
            break
        ex, ey = b[0] - a[0], b[1] - a[1]
        side = [ex * (p[1] - a[1]) - ey * (p[0] - a[0]) for p in out]
        nxt = []
        for i, p in enumerate(out):
            q, sp, sq = out[i - 1], side[i], side[i - 1]
            if (sp >= 0) != (sq >= 0):
                t = sq / (sq - sp)
                nxt.append((q[0] + t * (p[0] - q[0]), q[1] + t * (p[1] - q[1])))
            if sp >= 0:
                nxt.append(p)
        out = nxt
    if len(out) < 3:
        return np.zeros((0, 2))
    return np.array(out + [out[0]])


def overlap_percentage(rect_ring, polygon):
    """100 * area(rect ∩ polygon) / area(rect) (car_fuser.py:76-79).  shapely when installed (as the
    reference); otherwise the rectangle, being convex, clips the polygon's rings directly."""
    if bgt_utils.HAVE_SHAPELY and hasattr(polygon, 'intersection'):
        from shapely.geometry import Polygon
        p1 = Polygon(rect_ring)
        if not p1.intersects(polygon):
            return 0.0
        return (p1.intersection(polygon).area / p1.area) * 100
    rect = np.asarray(rect_ring, dtype=np.float64)
    area = _ring_area(rect)
    if area == 0.0:
        return 0.0
    if area < 0:
        rect, area = rect[::-1], -area
    ext = np.asarray(polygon.exterior.coords, dtype=np.float64)[:, :2]
    inter = abs(_ring_area(_clip_ring_by_convex(ext, rect)))
    for hole in polygon.interiors:
        inter -= abs(_ring_area(_clip_ring_by_convex(np.asarray(hole.coords, dtype=np.float64)[:, :2], rect)))
    return (inter / area) * 100
