"""Host-side math helpers that sit on the labelling path (drop-ins for
``upcp.utils.math_utils``: vector_angle :9-18, get_octree_level :21-33,
compute_bounding_box :36-57).  These are O(1)/O(n) host conveniences; the device
path computes the same quantities inside its kernels (bbox reduction, angle test)."""
import numpy as np


def vector_angle(u, v=np.array([0., 0., 1.])):
    """Angle in degrees between vectors 'u' and 'v' (default: the vertical axis)."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    c = np.dot(u / np.linalg.norm(u), v / np.linalg.norm(v))
    clip = np.minimum(1, np.maximum(c, -1))
    return np.rad2deg(np.arccos(clip))


def octree_level_from_extent(max_dim, grid_size):
    """Octree level for a cloud whose largest extent is ``max_dim``."""
    if max_dim < 0.001:
        return 0
    octree_level = np.rint(-np.log(grid_size / max_dim) / (np.log(2)))
    if octree_level > 0:
        return int(np.int64(octree_level))
    return 1


def get_octree_level(points, grid_size):
    """Nearest octree level for a desired grid_size (largest extent over all columns)."""
    points = np.asarray(points)
    dims = np.zeros((points.shape[1],))
    for d in range(points.shape[1]):
        dims[d] = np.max(points[:, d]) - np.min(points[:, d])
    return octree_level_from_extent(np.max(dims), grid_size)


def compute_bounding_box(points):
    """(x_min, y_min, x_max, y_max) of a point list; further columns are ignored."""
    points = np.asarray(points)
    return (np.min(points[:, 0]), np.min(points[:, 1]), np.max(points[:, 0]), np.max(points[:, 1]))
