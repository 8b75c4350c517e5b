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
