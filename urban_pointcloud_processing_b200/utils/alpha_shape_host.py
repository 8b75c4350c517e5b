"""Concave hull (alpha shape) of a small 2-D point set on the host -- scipy Delaunay +
shapely, used only by the shapely-gated ground refinement (reference
utils/alpha_shape_utils.py:11-35: triangles whose circumradius is below 1/alpha are kept,
their union is the hull)."""
import numpy as np


def alpha_shape(points, alpha=1.0):
    from scipy.spatial import Delaunay
    from shapely.geometry import MultiPoint, Polygon
    from shapely.ops import unary_union
    points = np.asarray(points, dtype=np.float64)
    if len(points) < 4:
        return [MultiPoint(points).convex_hull]
    tri = Delaunay(points)
    keep = []
    for ia, ib, ic in tri.simplices:
        pa, pb, pc = points[ia], points[ib], points[ic]
        a = np.linalg.norm(pa - pb); b = np.linalg.norm(pb - pc); c = np.linalg.norm(pc - pa)
        s = (a + b + c) / 2.0
        area = np.sqrt(max(s * (s - a) * (s - b) * (s - c), 0.0))
        if area > 0 and a * b * c / (4.0 * area) < 1.0 / alpha:
            keep.append(Polygon([pa, pb, pc]))
    if not keep:
        return [MultiPoint(points).convex_hull]
    union = unary_union(keep)
    return [union] if union.geom_type == 'Polygon' else list(union.geoms)
