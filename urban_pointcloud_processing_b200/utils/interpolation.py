"""Raster lookup drop-in for ``upcp.utils.interpolation.FastGridInterpolator``
(reference src/upcp/utils/interpolation.py:311-348), evaluated on the GPU."""
import numpy as np

from .. import device as dev


class FastGridInterpolator:
    """Nearest-cell lookup in gridded data; grid coordinates are cell centroids.

    Parameters
    ----------
    grid_x : array-like, ascending x-coordinates of the grid
    grid_y : array-like, descending y-coordinates of the grid
    values : array of shape (Ny, Nx)
    """

    def __init__(self, grid_x, grid_y, values):
        grid_x = np.asarray(grid_x, dtype=np.float64)
        grid_y = np.asarray(grid_y, dtype=np.float64)
        step_x = grid_x[1] - grid_x[0]
        step_y = grid_y[0] - grid_y[1]
        self.bin_x = grid_x - (step_x / 2)
        self.bin_y = grid_y + (step_y / 2)
        self.values = values
        self._grid = (grid_x, grid_y)
        self._raster = None

    def device_raster(self):
        if self._raster is None:
            self._raster = dev.DeviceRaster(self._grid[0], self._grid[1], self.values)
        return self._raster

    def __call__(self, positions):
        """Evaluate at ``positions`` (Np, >=2): column 0 = x, column 1 = y."""
        positions = np.asarray(positions, dtype=np.float64)
        tile = dev.DeviceTile(positions[:, :2]) if positions.shape[1] != 2 else dev.DeviceTile(positions)
        out = self.device_raster().lookup(tile)
        return out.cpu().numpy()


class SpatialInterpolator:
    """Inverse-distance-weighting / maximum interpolator over scattered 2-D data, evaluated on the GPU
    (drop-in for ``upcp.utils.interpolation.SpatialInterpolator``, reference utils/interpolation.py:17-308,
    as AHN preprocessing uses it: preprocessing/ahn_preprocessing.py:155-162).

    Parameters (names and defaults of the reference)
    ----------
    coordinates : array (N, 2);  values : array (N,);  method : 'idw' | 'max';  weights : None;  leafsize : ignored

    The reference queries a scipy cKDTree and loops over the positions in Python; here the data points are
    binned into a uniform grid with cells of ``max_dist`` and every position is one GPU thread
    (``upcp_interpolate``).  ``max_dist`` must be finite (the preprocessing passes 1.0 / 0.5).
    """

    METHODS = {'idw': 0, 'max': 1}

    def __init__(self, coordinates, values, method='idw', weights=None, leafsize=10):
        coordinates = np.asarray(coordinates, dtype=np.float64)
        if coordinates.ndim != 2 or coordinates.shape[1] != 2:
            raise ValueError('the B200 SpatialInterpolator handles 2-D coordinates (N, 2)')
        values = np.asarray(values, dtype=np.float64)
        if values.ndim != 1 or coordinates.shape[0] != values.size:
            raise ValueError('The number of values must match the number of coordinates.')
        if weights is not None:
            raise NotImplementedError('per-point weights are not used by the AHN preprocessing and not supported')
        if method not in self.METHODS:
            raise ValueError(f'method must be one of {tuple(self.METHODS)}')
        self.coords_ndim = 2
        self.coordinates, self.values, self.method, self.weights = coordinates, values, method, None

    def __call__(self, positions, n_neighbors=8, max_dist=np.inf, eps=0.0, power=1.0, reg=0.0, conf_dist=1e-12,
                 fill_value=np.nan, dtype=float, workers=1):
        import ctypes

        import torch
        from .. import _lib
        lib = _lib.require_cuda()
        n_neighbors = int(n_neighbors)
        if n_neighbors < 1:
            raise ValueError('n_neighbors must be a positive integer')
        if not np.isfinite(max_dist):
            raise NotImplementedError('the B200 SpatialInterpolator needs a finite max_dist')
        if eps != 0.0:
            raise NotImplementedError('approximate neighbour search (eps > 0) is not supported')
        if conf_dist is None or conf_dist <= 0.0:
            conf_dist = 0.0
        positions = np.reshape(np.asanyarray(positions, dtype=np.float64), (-1, 2))
        device = dev.current_device()
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)      # noqa: E731
        px, py, pv = up(self.coordinates[:, 0]), up(self.coordinates[:, 1]), up(self.values)
        qx, qy = up(positions[:, 0]), up(positions[:, 1])
        out = torch.empty(len(positions), dtype=torch.float64, device=device)
        n = len(self.values)
        bbox = (ctypes.c_double * 4)(float(self.coordinates[:, 0].min()), float(self.coordinates[:, 1].min()),
                                     float(self.coordinates[:, 0].max()), float(self.coordinates[:, 1].max()))
        ws = dev.workspace(device).get(lib.upcp_interpolate_ws_bytes(n))
        _lib.check(lib.upcp_interpolate(dev.ptr(px), dev.ptr(py), dev.ptr(pv), n, bbox, dev.ptr(qx), dev.ptr(qy),
                                        len(positions), self.METHODS[self.method], n_neighbors, float(max_dist),
                                        float(power), float(reg), float(conf_dist), float(fill_value), dev.ptr(out),
                                        dev.ptr(ws), ws.numel(), dev.stream_ptr()), 'upcp_interpolate')
        res = out.cpu().numpy()
        if dtype is not None and dtype is not float:
            res = res.astype(dtype)
        return res[0] if len(res) == 1 else res
