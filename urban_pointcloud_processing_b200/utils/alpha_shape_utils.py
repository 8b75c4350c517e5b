"""Concave hull (alpha shape) of a 2-D point set -- the geometric half of
``upcp.utils.alpha_shape_utils`` (reference utils/alpha_shape_utils.py:11-92), without shapely.

The reference triangulates the points (scipy Delaunay), keeps the triangles whose circumradius is below
``1 / alpha``, collects the edges that belong to exactly ONE kept triangle (``get_alpha_shape_edges(...,
only_outer=True)``) and lets shapely stitch them into polygons with holes (``generate_poly_from_edges``).
As a point set those polygons are the union of the kept triangles, whose interior is where a ray crosses
the outer edges an odd number of times -- whatever loop the edges are stitched into.  So here the edges are
walked into closed trails (any order) and handed to the point-in-polygon kernel as the loops of an even-odd
``LoopPolygon``; ``buffer(d)`` is carried as the polygon's pending offset (see ``bgt_utils.RingPolygon``).
"""
import numpy as np


def alpha_shape_edges(points, alpha):
    """(E, 2) int array of the outer edges -- ``get_alpha_shape_edges(points, alpha, only_outer=True)``
    (alpha_shape_utils.py:40-92) as an array of undirected vertex pairs (i < j), same arithmetic."""
    from scipy.spatial import Delaunay
    points = np.asarray(points, dtype=np.float64)
    assert points.shape[0] > 3, 'Need at least four points'
    tri = Delaunay(points).simplices
    pa, pb, pc = points[tri[:, 0]], points[tri[:, 1]], points[tri[:, 2]]
    a = np.sqrt((pa[:, 0] - pb[:, 0]) ** 2 + (pa[:, 1] - pb[:, 1]) ** 2)
    b = np.sqrt((pb[:, 0] - pc[:, 0]) ** 2 + (pb[:, 1] - pc[:, 1]) ** 2)
    c = np.sqrt((pc[:, 0] - pa[:, 0]) ** 2 + (pc[:, 1] - pa[:, 1]) ** 2)
    s = (a + b + c) / 2.0
    int_var = s * (s - a) * (s - b) * (s - c)
    ok = int_var > 0
    with np.errstate(invalid='ignore', divide='ignore'):
        circum_r = a * b * c / (4.0 * np.sqrt(np.where(ok, int_var, 1.0)))
    keep = ok & (circum_r < (1 / alpha))
    t = tri[keep]
    e = np.concatenate((t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]), axis=0)
    e = np.sort(e, axis=1)
    if len(e) == 0:
        return np.zeros((0, 2), dtype=np.int64)
    uniq, counts = np.unique(e, axis=0, return_counts=True)
    return uniq[counts == 1].astype(np.int64)


def closed_trails(edges):
    """Walk undirected edges (every vertex of even degree, as the outer edges of a triangle complex are)
    into closed trails; returns lists of vertex ids with the first repeated at the end."""
    adj = {}
    for k, (i, j) in enumerate(edges):
        adj.setdefault(int(i), []).append((int(j), k))
        adj.setdefault(int(j), []).append((int(i), k))
    used = np.zeros(len(edges), dtype=bool)
    trails = []
    for k0, (i0, j0) in enumerate(edges):
        if used[k0]:
            continue
        used[k0] = True
        trail = [int(i0), int(j0)]
        cur = int(j0)
        while cur != trail[0]:
            nxt = None
            lst = adj[cur]
            while lst:
                v, k = lst.pop()
                if not used[k]:
                    nxt = (v, k)
                    break
            if nxt is None:          # odd-degree vertex: cannot happen for a triangle complex; close the trail
                trail.append(trail[0])
                break
            used[nxt[1]] = True
            trail.append(nxt[0])
            cur = nxt[0]
        trails.append(trail)
    return trails


class _Loop:
    def __init__(self, coords):
        self.coords = [tuple(c) for c in np.asarray(coords, dtype=np.float64).reshape(-1, 2)]


class LoopPolygon:
    """A region bounded by closed loops under the EVEN-ODD rule (outer boundaries, holes, islands in holes:
    no nesting analysis needed), with an optional pending buffer distance.  Quacks like the polygons
    ``poly_clip`` reads where it can: ``exterior`` is the first loop, ``interiors`` is empty; the device plan
    reads ``loops`` and ``offset``."""

    def __init__(self, loops, offset=0.0):
        self.loops = [_Loop(l) for l in loops]
        self.offset = float(offset)
        self.exterior = self.loops[0] if self.loops else _Loop(np.zeros((0, 2)))
        self.interiors = []

    def buffer(self, offset):
        if offset == 0:
            return self
        out = LoopPolygon.__new__(LoopPolygon)
        out.loops, out.exterior, out.interiors = self.loops, self.exterior, self.interiors
        out.offset = self.offset + float(offset)
        return out


def alpha_shape(points, alpha=1.0):
    """Concave hull of ``points`` as a list with ONE even-odd ``LoopPolygon`` (the reference returns the same
    region as a list of shapely polygons with holes, alpha_shape_utils.py:11-35)."""
    points = np.asarray(points, dtype=np.float64)
    edges = alpha_shape_edges(points, alpha)
    loops = [points[t] for t in closed_trails(edges) if len(t) >= 4]
    return [LoopPolygon(loops)] if loops else []


def convex_hull_poly(points):
    """math_utils.convex_hull_poly (math_utils.py:60-62): hull ring of scipy ConvexHull with option QJ."""
    from scipy.spatial import ConvexHull
    points = np.asarray(points, dtype=np.float64)
    ring = points[ConvexHull(points, qhull_options='QJ').vertices]
    return LoopPolygon([np.vstack([ring, ring[:1]])])
