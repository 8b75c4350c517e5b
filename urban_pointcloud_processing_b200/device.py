"""Device-resident tile state shared by the drop-in processors.

PyTorch is used here only as plumbing: device memory (``torch.empty``), the
current CUDA stream and host<->device copies.  All arithmetic is done by the
hand-written kernels behind the C ABI (``_lib``).

HBM layout of one tile (N points):
    x, y, z    fp64 [N] each (structure of arrays; 24 B/pt)
    labels     int32 [N]
    masks      uint8 [N] (0 / 1), allocated per processor call
    rasters    fp64 [ny, nx] + bin edges (2 MB per 500x500 AHN surface, L2-resident)
    PIP plans  one blob per polygon set (KBs .. a few MB, L2-resident)
"""
import ctypes
import warnings
import weakref

import numpy as np
import torch

from . import _lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _from_numpy(a):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # read-only numpy arrays
        return torch.from_numpy(a)


class Workspace:
    """Grow-only scratch buffer handed to the C ABI (caller-owned workspace)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self.buf


_workspaces = {}


def stream_key(device=None):
    """Handle of the current CUDA stream: scratch memory and cached device objects are kept per
    stream, so that tiles in flight on different streams (Pipeline.process_clouds workers) never
    share a buffer and freed memory is only ever recycled in stream order."""
    return int(torch.cuda.current_stream(device).cuda_stream)


def workspace(device, slot=0):
    key = (str(device), slot, stream_key(device))
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = Workspace(device)
    return ws


_worker_streams = {}


def worker_stream(device, name):
    """Persistent side stream ``name`` of ``device``.  The streams are created once and re-used by
    every batch call: the caching allocator keeps freed blocks per stream, so fresh streams per call
    would strand the previous call's memory in pools nobody allocates from any more."""
    key = (str(device), name)
    st = _worker_streams.get(key)
    if st is None:
        st = _worker_streams[key] = torch.cuda.Stream(device)
    return st


def release_workspaces(device, stream_handle):
    """Drop the scratch buffers of one stream (a finished worker)."""
    for key in [k for k in list(_workspaces) if k[0] == str(device) and k[2] == int(stream_handle)]:
        _workspaces.pop(key, None)


def current_device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


# --------------------------------------------------------------------- masks --
def upload_mask(mask, n, device):
    """numpy bool[n] (or None) -> uint8 cuda tensor (or None)."""
    if mask is None:
        return None
    if isinstance(mask, torch.Tensor):
        return mask
    m = np.ascontiguousarray(np.asarray(mask, dtype=bool))
    if m.shape != (n,):
        raise ValueError(f'mask has shape {m.shape}, expected ({n},)')
    return _from_numpy(m.view(np.uint8)).to(device, non_blocking=True)


def download_mask(d_mask):
    return d_mask.cpu().numpy().view(bool)


def mask_build(d_labels, d_mask_in=None, base=False, eq=(), ne=()):
    """out = ((mask_in or base) | labels in eq) & labels not in ne  (see upcp_mask_build)."""
    lib = _lib.load()
    n = d_labels.numel()
    out = torch.empty(n, dtype=torch.uint8, device=d_labels.device)
    eq = [int(v) for v in eq]
    ne = [int(v) for v in ne]
    if len(eq) > 8 or len(ne) > 8:
        raise ValueError('at most 8 labels per set')
    eq_arr = (ctypes.c_int32 * 8)(*eq)
    ne_arr = (ctypes.c_int32 * 8)(*ne)
    _lib.check(lib.upcp_mask_build(ptr(d_labels), n, ptr(d_mask_in), int(bool(base)), eq_arr, len(eq),
                                   ne_arr, len(ne), ptr(out), stream_ptr()), 'upcp_mask_build')
    return out


def mask_combine(d_a, d_b, op, out=None):
    """op: 'and' | 'or' | 'andnot'."""
    lib = _lib.load()
    code = {'and': 0, 'or': 1, 'andnot': 2}[op]
    if out is None:
        out = torch.empty_like(d_a)
    _lib.check(lib.upcp_mask_combine(ptr(d_a), ptr(d_b), d_a.numel(), code, ptr(out), stream_ptr()),
               'upcp_mask_combine')
    return out


def mask_count(d_mask):
    """Number of set entries (synchronises)."""
    lib = _lib.load()
    cnt = torch.empty(1, dtype=torch.int64, device=d_mask.device)
    _lib.check(lib.upcp_mask_count(ptr(d_mask), d_mask.numel(), ptr(cnt), stream_ptr()), 'upcp_mask_count')
    return int(cnt.item())


def labels_assign(d_labels, d_mask, value):
    lib = _lib.load()
    _lib.check(lib.upcp_labels_assign(ptr(d_labels), d_labels.numel(), ptr(d_mask), int(value), stream_ptr()),
               'upcp_labels_assign')


def label_hist(d_labels):
    """int64[256] histogram of label values (device tensor)."""
    lib = _lib.load()
    hist = torch.zeros(256, dtype=torch.int64, device=d_labels.device)
    _lib.check(lib.upcp_label_hist(ptr(d_labels), d_labels.numel(), ptr(hist), stream_ptr()), 'upcp_label_hist')
    return hist


# -------------------------------------------------------------------- labels --
def upload_labels(labels, device):
    """Host integer labels -> int32 cuda tensor (widened on the device for u8/u16)."""
    lib = _lib.load()
    if isinstance(labels, torch.Tensor):
        return labels
    a = np.asarray(labels)
    n = a.shape[0]
    out = torch.empty(n, dtype=torch.int32, device=device)
    if n == 0:
        return out
    if a.dtype == np.uint8 and a.flags.c_contiguous:
        raw = _from_numpy(a).to(device, non_blocking=True)
        _lib.check(lib.upcp_labels_widen_u8(ptr(raw), n, ptr(out), stream_ptr()), 'upcp_labels_widen_u8')
    elif a.dtype == np.uint16 and a.flags.c_contiguous:
        raw = _from_numpy(a.view(np.int16)).to(device, non_blocking=True)
        _lib.check(lib.upcp_labels_widen_u16(ptr(raw), n, ptr(out), stream_ptr()), 'upcp_labels_widen_u16')
    else:
        out.copy_(_from_numpy(np.ascontiguousarray(a.astype(np.int32))), non_blocking=True)
    return out


def download_labels(d_labels, labels):
    """Write the device labels back INTO the host array `labels` (in-place contract)."""
    lib = _lib.load()
    n = d_labels.numel()
    if n == 0:
        return labels
    if labels.dtype == np.uint8 and labels.flags.c_contiguous and labels.flags.writeable:
        nar = torch.empty(n, dtype=torch.uint8, device=d_labels.device)
        _lib.check(lib.upcp_labels_narrow_u8(ptr(d_labels), n, ptr(nar), stream_ptr()), 'upcp_labels_narrow_u8')
        _from_numpy(labels).copy_(nar)
    elif labels.dtype == np.uint16 and labels.flags.c_contiguous and labels.flags.writeable:
        nar = torch.empty(n, dtype=torch.int16, device=d_labels.device)
        _lib.check(lib.upcp_labels_narrow_u16(ptr(d_labels), n, ptr(nar), stream_ptr()), 'upcp_labels_narrow_u16')
        _from_numpy(labels.view(np.int16)).copy_(nar)
    else:
        labels[...] = d_labels.cpu().numpy().astype(labels.dtype)
    return labels


# ---------------------------------------------------------------------- tile --
class DeviceTile:
    """One point-cloud tile resident in HBM as fp64 structure-of-arrays.

    Accepts the ``points`` argument of the reference plugin contract
    (abstract_processor.py:23-43): float64 ``(N, 3)`` (or ``(N, 2)``), C- or
    F-ordered (pipeline.py:124 delivers F-order).  Never mutates it.
    """

    def __init__(self, points, device=None):
        lib = _lib.require_cuda()
        self.device = device if device is not None else current_device()
        pts = np.asarray(points)
        if pts.ndim != 2 or pts.shape[1] not in (2, 3):
            raise ValueError('points must have shape (n_points, 3) or (n_points, 2)')
        if pts.dtype != np.float64:
            pts = pts.astype(np.float64)
        self.n = int(pts.shape[0])
        self.ncols = int(pts.shape[1])
        n = self.n
        self.x = torch.empty(n, dtype=torch.float64, device=self.device)
        self.y = torch.empty(n, dtype=torch.float64, device=self.device)
        self.z = torch.empty(n, dtype=torch.float64, device=self.device)
        self.h2d_bytes = n * self.ncols * 8
        if n:
            if pts.strides[0] == 8:            # columns contiguous (F-order): 2-3 plain H2D copies
                self.x.copy_(_from_numpy(pts[:, 0]), non_blocking=True)
                self.y.copy_(_from_numpy(pts[:, 1]), non_blocking=True)
                if self.ncols == 3:
                    self.z.copy_(_from_numpy(pts[:, 2]), non_blocking=True)
                else:
                    self.z.zero_()
            else:                              # C-order AoS: one H2D copy + transposition kernel
                aos = _from_numpy(np.ascontiguousarray(pts)).to(self.device, non_blocking=True)
                _lib.check(lib.upcp_points_aos_to_soa(ptr(aos), n, self.ncols, ptr(self.x), ptr(self.y),
                                                      ptr(self.z), stream_ptr()), 'upcp_points_aos_to_soa')
        self._bbox_all = None

    @classmethod
    def from_device(cls, x, y, z):
        """Wrap coordinates that already live in HBM (bench / on-device generator)."""
        self = cls.__new__(cls)
        _lib.require_cuda()
        self.device = x.device
        self.n = int(x.numel())
        self.ncols = 3
        self.x, self.y, self.z = x, y, z
        self.h2d_bytes = 0
        self._bbox_all = None
        return self

    @classmethod
    def from_las(cls, las, device=None):
        """Tile from a LAS object (``laspy.LasData`` or ``utils.las_io.LasData``): the raw int32
        ``X, Y, Z`` are uploaded (12 B per point) and ``x = X * scale + offset`` is evaluated in
        fp64 on the device -- bit for bit the host-side scaled view the reference stacks into
        ``points`` (pipeline.py:123-124)."""
        lib = _lib.require_cuda()
        self = cls.__new__(cls)
        self.device = device if device is not None else current_device()
        ints = [np.ascontiguousarray(np.asarray(getattr(las, k)), dtype=np.int32) for k in ('X', 'Y', 'Z')]
        n = int(ints[0].shape[0])
        self.n, self.ncols = n, 3
        self.x = torch.empty(n, dtype=torch.float64, device=self.device)
        self.y = torch.empty(n, dtype=torch.float64, device=self.device)
        self.z = torch.empty(n, dtype=torch.float64, device=self.device)
        self.h2d_bytes = n * 12
        self._bbox_all = None
        if n:
            d_ints = [_from_numpy(a).to(self.device, non_blocking=True) for a in ints]
            scale = (ctypes.c_double * 3)(*[float(v) for v in las.header.scales])
            offset = (ctypes.c_double * 3)(*[float(v) for v in las.header.offsets])
            _lib.check(lib.upcp_points_from_scaled_int32(ptr(d_ints[0]), ptr(d_ints[1]), ptr(d_ints[2]), n, scale,
                                                         offset, ptr(self.x), ptr(self.y), ptr(self.z), stream_ptr()),
                       'upcp_points_from_scaled_int32')
        return self

    def bbox(self, d_sel=None):
        """(bbox_all[6], bbox_sel[6], n_sel) in fp64 -- one reduction + one small D2H."""
        lib = _lib.load()
        out = torch.empty(13, dtype=torch.float64, device=self.device)
        ws = workspace(self.device).get(lib.upcp_bbox_ws_bytes(self.n))
        _lib.check(lib.upcp_bbox(ptr(self.x), ptr(self.y), ptr(self.z), self.n, ptr(d_sel), ptr(out),
                                 ptr(ws), ws.numel(), stream_ptr()), 'upcp_bbox')
        h = out.cpu().numpy()
        return h[0:6].copy(), h[6:12].copy(), int(h[12])

    def bbox_all(self):
        if self._bbox_all is None:
            self._bbox_all = self.bbox(None)[0]
        return self._bbox_all


# HBM copies are re-used only where the host array cannot have changed behind our back: READ-ONLY numpy
# arrays (``points.flags.writeable == False``), keyed on the identity of the live array.  A writable array
# -- a staging buffer that is refilled in place, a cloud the caller jitters between two calls -- is uploaded
# on every call, exactly as the reference reads the live array on every call.  Callers that want to keep a
# writable cloud resident pass the ``DeviceTile`` itself wherever ``points`` is expected.
_TILE_CACHE = []
_TILE_CACHE_SIZE = 2


def get_tile(points, device=None):
    """DeviceTile for ``points`` (a DeviceTile is returned as it is)."""
    if isinstance(points, DeviceTile):
        return points
    device = device if device is not None else current_device()
    pts = np.asarray(points)
    if pts.flags.writeable:
        return DeviceTile(pts, device)
    for ref, dev_, tile in _TILE_CACHE:
        if ref() is pts and dev_ == device:
            return tile
    tile = DeviceTile(pts, device)
    try:
        ref = weakref.ref(pts)
    except TypeError:
        return tile
    _TILE_CACHE.append((ref, device, tile))
    while len(_TILE_CACHE) > _TILE_CACHE_SIZE:
        _TILE_CACHE.pop(0)
    return tile


def clear_tile_cache():
    _TILE_CACHE.clear()


# -------------------------------------------------------------------- raster --
class DeviceRaster:
    """FastGridInterpolator state in HBM (interpolation.py:329-334)."""

    def __init__(self, grid_x, grid_y, values, device=None):
        _lib.require_cuda()
        self.device = device if device is not None else current_device()
        grid_x = np.asarray(grid_x, dtype=np.float64)
        grid_y = np.asarray(grid_y, dtype=np.float64)
        values = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        # Same host arithmetic as the reference constructor.
        step_x = grid_x[1] - grid_x[0]
        step_y = grid_y[0] - grid_y[1]
        bin_x = grid_x - (step_x / 2)
        bin_y = grid_y + (step_y / 2)
        self.nx = int(bin_x.shape[0])
        self.ny = int(bin_y.shape[0])
        if values.shape != (self.ny, self.nx):
            raise ValueError(f'values has shape {values.shape}, expected ({self.ny}, {self.nx})')
        self.bin_x = _from_numpy(np.ascontiguousarray(bin_x)).to(self.device)
        self.bin_y = _from_numpy(np.ascontiguousarray(bin_y)).to(self.device)
        self.values = _from_numpy(values).to(self.device)

    def lookup(self, tile, d_sel=None, out=None):
        """values[y_idx, x_idx] per point -> fp64 cuda tensor [N]."""
        lib = _lib.load()
        if out is None:
            out = torch.full((tile.n,), float('nan'), dtype=torch.float64, device=self.device)
        _lib.check(lib.upcp_raster_lookup(ptr(tile.x), ptr(tile.y), tile.n, ptr(d_sel), ptr(self.bin_x),
                                          self.nx, ptr(self.bin_y), self.ny, ptr(self.values), ptr(out),
                                          stream_ptr()), 'upcp_raster_lookup')
        return out

    def classify(self, tile, mode, a=0.0, b=0.0, d_sel=None, d_labels=None, assign=0, want_mask=True):
        """Fused lookup + height-threshold test -> uint8 mask [N] (and optional label write)."""
        lib = _lib.load()
        out = torch.empty(tile.n, dtype=torch.uint8, device=self.device) if want_mask else None
        _lib.check(lib.upcp_raster_classify(ptr(tile.x), ptr(tile.y), ptr(tile.z), tile.n, ptr(d_sel),
                                            ptr(self.bin_x), self.nx, ptr(self.bin_y), self.ny,
                                            ptr(self.values), int(mode), float(a), float(b), ptr(out),
                                            ptr(d_labels), int(assign), stream_ptr()),
                   'upcp_raster_classify')
        return out


# ----------------------------------------------------------------------- PIP --
def rings_from_polygons(polygons):
    """Flatten shapely-like polygons (``.exterior.coords`` / ``.interiors``) or raw
    rings into the ring arrays of the C ABI.  Matches what clip_utils.poly_clip reads
    (clip_utils.py:211-212)."""
    xy, off, poly_id, hole = [], [0], [], []
    if all(type(poly).__name__ == 'RingPolygon' for poly in polygons):
        # the readers' own polygon type holds its rings as closed (n, 2) float64 arrays: no per-ring conversion
        # (a BGT tile with thousands of footprints spent 20 ms of host time per fuser call in the generic loop)
        counts = []
        for p, poly in enumerate(polygons):
            for is_hole, ring in enumerate([poly.exterior] + list(poly.interiors)):
                a = getattr(ring, 'xy', None)
                if a is None:                              # (a ring object that only offers .coords)
                    a = np.asarray(ring.coords, dtype=np.float64).reshape(-1, 2)
                xy.append(a); counts.append(len(a)); poly_id.append(p); hole.append(1 if is_hole else 0)
        ring_xy = np.ascontiguousarray(np.concatenate(xy, axis=0) if xy else np.zeros((0, 2)), dtype=np.float64)
        off_arr = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=off_arr[1:])
        return ring_xy, off_arr, np.asarray(poly_id, dtype=np.int32), np.asarray(hole, dtype=np.uint8)
    for p, poly in enumerate(polygons):
        if hasattr(poly, 'loops'):                     # even-odd region (alpha_shape_utils.LoopPolygon): ring kind 2
            for loop in poly.loops:
                ring = np.asarray(loop.coords, dtype=np.float64).reshape(-1, 2)
                xy.append(ring)
                off.append(off[-1] + ring.shape[0])
                poly_id.append(p)
                hole.append(2)
            continue
        if hasattr(poly, 'exterior'):
            ext = np.asarray(poly.exterior.coords, dtype=np.float64)
            ints = [np.asarray(i.coords, dtype=np.float64) for i in poly.interiors]
        else:
            ext = np.asarray(poly, dtype=np.float64)
            ints = []
        rings = [(ext, 0)] + [(i, 1) for i in ints]
        for ring, is_hole in rings:
            ring = ring.reshape(-1, ring.shape[-1] if ring.ndim == 2 else 2)[:, :2] if ring.size else ring.reshape(0, 2)
            xy.append(ring)
            off.append(off[-1] + ring.shape[0])
            poly_id.append(p)
            hole.append(is_hole)
    ring_xy = np.ascontiguousarray(np.concatenate(xy, axis=0) if xy else np.zeros((0, 2)), dtype=np.float64)
    return (ring_xy, np.asarray(off, dtype=np.int64), np.asarray(poly_id, dtype=np.int32),
            np.asarray(hole, dtype=np.uint8))


class PipPlan:
    """Device-resident point-in-polygon plan for a set of polygons.  Polygons that carry a pending buffer
    (``RingPolygon.offset``, i.e. ``filter_tile(offset=...)`` / ``poly.buffer(d)`` without shapely) are
    grouped by offset into buffered sub-plans; the query ORs the groups."""

    def __init__(self, polygons, device=None, grid_dim=0):
        _lib.require_cuda()
        self.device = device if device is not None else current_device()
        groups = {}
        for poly in polygons:
            groups.setdefault(float(getattr(poly, 'offset', 0.0) or 0.0), []).append(poly)
        self.parts = [self._build(polys, off, grid_dim) for off, polys in sorted(groups.items())]
        self.n_rings = sum(p[2] for p in self.parts)
        self.n_vertices = sum(p[3] for p in self.parts)
        self.nbytes = sum(int(p[0].numel()) for p in self.parts)
        self.blob = self.parts[0][0] if len(self.parts) == 1 else None

    def _build(self, polygons, offset, grid_dim):
        lib = _lib.load()
        ring_xy, off, poly_id, hole = rings_from_polygons(polygons)
        handle = ctypes.c_void_p()
        _lib.check(lib.upcp_pip_plan_create_buffered(
            ring_xy.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
            poly_id.ctypes.data_as(ctypes.c_void_p), hole.ctypes.data_as(ctypes.c_void_p),
            int(poly_id.shape[0]), int(grid_dim), float(offset), ctypes.byref(handle)), 'upcp_pip_plan_create_buffered')
        try:
            nbytes = lib.upcp_pip_plan_bytes(handle)
            blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(lib.upcp_pip_plan_upload(handle, ptr(blob), nbytes, stream_ptr()), 'upcp_pip_plan_upload')
        finally:
            lib.upcp_pip_plan_destroy(handle)
        return blob, float(offset), int(poly_id.shape[0]), int(ring_xy.shape[0])

    def query(self, tile, d_sel=None):
        """uint8 mask [N]: point inside the polygon set (restricted to d_sel)."""
        lib = _lib.load()
        out = None
        for blob, offset, _, _ in self.parts:
            part = torch.empty(tile.n, dtype=torch.uint8, device=self.device)
            fn, name = (lib.upcp_pip_query, 'upcp_pip_query') if offset == 0.0 else \
                (lib.upcp_pip_query_buffered, 'upcp_pip_query_buffered')
            _lib.check(fn(ptr(blob), ptr(tile.x), ptr(tile.y), tile.n, ptr(d_sel), ptr(part), stream_ptr()), name)
            out = part if out is None else mask_combine(out, part, 'or', out=out)
        if out is None:
            out = torch.zeros(tile.n, dtype=torch.uint8, device=self.device)
        return out
