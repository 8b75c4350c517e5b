"""Clipping tools evaluated on the GPU (drop-ins for the ★ part of
``upcp.utils.clip_utils``: rectangle_clip :22-40, is_inside :167-190, poly_clip
:193-238, poly_box_clip :241-262)."""
import numpy as np

from .. import device as dev


def rectangle_clip(points, rect):
    """Mask of points within (x_min, y_min, x_max, y_max), inclusive."""
    points = np.asarray(points)
    return ((points[:, 0] >= rect[0]) & (points[:, 0] <= rect[2])
            & (points[:, 1] >= rect[1]) & (points[:, 1] <= rect[3]))


def poly_clip_device(tile, polygons, d_sel=None):
    """uint8 cuda mask of the points of ``tile`` inside any of ``polygons``."""
    plan = dev.PipPlan(polygons, device=tile.device)
    return plan.query(tile, d_sel)


def poly_clip(points, poly):
    """Mask of all points within a polygon (boundary counts as inside; interiors are
    excluded).  ``poly`` exposes ``.exterior.coords`` and ``.interiors``."""
    points = np.asarray(points, dtype=np.float64)
    if len(points) == 0:
        return np.zeros((0,), dtype=bool)
    tile = dev.get_tile(points)
    return dev.download_mask(poly_clip_device(tile, [poly]))


def is_inside(x, y, polygon):
    """For each (x[i], y[i]): inside / on the boundary of the linear ring ``polygon``."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if len(x) == 0:
        return np.zeros((0,), dtype=bool)
    ring = np.asarray(polygon, dtype=np.float64)
    # No bbox prefilter in is_inside: the exact edge test alone decides, which is what the
    # plan computes for points inside the ring's bbox; outside the bbox the crossing
    # number is 0 and no boundary case can trigger, so the result is identical.
    tile = dev.DeviceTile(np.column_stack((x, y)))
    return dev.download_mask(dev.PipPlan([ring], device=tile.device).query(tile))


def poly_box_clip(points, poly, bottom=-np.inf, top=np.inf):
    """Mask of all points within a 3D polygon with fixed height."""
    points = np.asarray(points)
    clip_mask = poly_clip(points, poly)
    return clip_mask & ((points[:, 2] <= top) & (points[:, 2] >= bottom))
