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


def minimum_bounding_rectangle(points):
    """Smallest-area enclosing rectangle of an (n, 2) point set (drop-in for
    math_utils.py:62-125; per-object host work of CarFuser / BGTStreetFurnitureFuser).

    Returns (corners (4, 2), hull_points, width, length, center_point) with width <= length.
    The candidate orientations are the hull edge directions folded into [0, pi/2); as in the
    reference the hull is walked as an OPEN chain (the edge closing the qhull vertex cycle is not
    a candidate), and every product goes through the same numpy calls on the same operands so
    that the corners come out bit for bit.

    Said plainly: this is the reference's function restated, not a redesign.  It is host code
    for a handful of clusters per tile; the corners feed a point-in-prism test whose membership must be
    bit-exact, and they only are when cos / arctan2 / dot are numpy's (glibc, BLAS) on the same operands -- a
    device hull + rotating calipers would differ in the last ulp of a corner.  The device side of these
    fusers is objects.cu (cluster statistics, gathers, prism clip)."""
    from scipy.spatial import ConvexHull
    points = np.asarray(points)
    quarter = np.pi / 2.
    hull_points = points[ConvexHull(points).vertices]
    steps = hull_points[1:] - hull_points[:-1]
    theta = np.unique(np.abs(np.mod(np.arctan2(steps[:, 1], steps[:, 0]), quarter)))
    # one 2 x 2 rotation per candidate angle: [[c, cos(t - pi/2)], [cos(t + pi/2), c]]
    frames = np.vstack([np.cos(theta), np.cos(theta - quarter), np.cos(theta + quarter), np.cos(theta)]).T
    frames = frames.reshape((-1, 2, 2))
    turned = np.dot(frames, hull_points.T)
    lo_x, hi_x = np.nanmin(turned[:, 0], axis=1), np.nanmax(turned[:, 0], axis=1)
    lo_y, hi_y = np.nanmin(turned[:, 1], axis=1), np.nanmax(turned[:, 1], axis=1)
    best = np.argmin((hi_x - lo_x) * (hi_y - lo_y))
    x1, x2, y1, y2 = hi_x[best], lo_x[best], hi_y[best], lo_y[best]
    frame = frames[best]
    center_point = np.dot([(x1 + x2) / 2, (y1 + y2) / 2], frame)
    corners = np.zeros((4, 2))
    for row, corner in enumerate(([x1, y2], [x2, y2], [x2, y1], [x1, y1])):
        corners[row] = np.dot(corner, frame)
    extent = [(x1 - x2), (y1 - y2)]
    return corners, hull_points, min(extent), max(extent), center_point
