"""numpy restatement of the reference's Python-level control flow on the labelling path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference
lines it follows (paths relative to /root/reference/src/upcp).  Natives the reference
gets from third-party wheels are taken from ``oracle.natives`` (plain-C restatement).
All functions are pure: inputs are never mutated unless stated.
"""
import numpy as np

from . import natives

GROUND, BUILDING, NOISE, UNKNOWN = 9, 10, 99, 0     # labels.py:9-24


# ------------------------------------------------------------------ raster --------
def fast_grid_lookup(grid_x, grid_y, values, positions):
    """FastGridInterpolator.__init__/__call__ (utils/interpolation.py:329-348)."""
    grid_x = np.asarray(grid_x); grid_y = np.asarray(grid_y)
    step_x = grid_x[1] - grid_x[0]
    step_y = grid_y[0] - grid_y[1]
    bin_x = grid_x - (step_x / 2)
    bin_y = grid_y + (step_y / 2)
    x_idx = np.digitize(positions[:, 0], bin_x) - 1
    y_idx = np.digitize(positions[:, 1], bin_y, right=True) - 1
    return values[y_idx, x_idx]


def interpolate(ahn_tile, points, surface='ground_surface'):
    """AHNReader.interpolate without cache (utils/ahn_utils.py:74-102)."""
    return fast_grid_lookup(ahn_tile['x'], ahn_tile['y'], ahn_tile[surface], points)


# --------------------------------------------------------------------- PIP --------
def compute_bounding_box(points):
    """math_utils.compute_bounding_box (utils/math_utils.py:36-57)."""
    return (np.min(points[:, 0]), np.min(points[:, 1]), np.max(points[:, 0]), np.max(points[:, 1]))


def rectangle_clip(points, rect):
    """clip_utils.rectangle_clip (utils/clip_utils.py:22-40)."""
    return ((points[:, 0] >= rect[0]) & (points[:, 0] <= rect[2])
            & (points[:, 1] >= rect[1]) & (points[:, 1] <= rect[3]))


def seg_dist2(ax, ay, bx, by, px, py):
    """Squared distance from the points (px, py) to the segment a-b; the expressions of the device code."""
    ex, ey = bx - ax, by - ay
    wx, wy = px - ax, py - ay
    len2 = ex * ex + ey * ey
    t = (wx * ex + wy * ey) / len2 if len2 > 0.0 else np.zeros_like(px)
    t = np.minimum(np.maximum(t, 0.0), 1.0)
    dx, dy = wx - t * ex, wy - t * ey
    return dx * dx + dy * dy


def poly_clip_buffered(points, poly, d, return_exposure=False):
    """poly_clip(points, poly.buffer(d)) with the buffer taken as a DISTANCE (what the device evaluates where
    the reference calls GEOS: bgt_utils.py:154-163, ahn_fuser.py:104-107): inside the polygon or within d of
    its boundary (d > 0); inside and farther than |d| from it (d < 0).  GEOS cuts the round joins at convex
    corners into 8 segments per quarter circle, i.e. its buffer is smaller by up to d (1 - cos(pi/32)) there:
    ``return_exposure`` also returns the number of points that lie in that sliver of a vertex' arc."""
    exterior = np.array(poly.exterior.coords)
    interiors = [np.array(interior.coords) for interior in poly.interiors if len(interior.coords) >= 3]
    n = len(points)
    if len(exterior) < 3:
        return (np.zeros(n, dtype=bool), 0) if return_exposure else np.zeros(n, dtype=bool)
    px, py = np.ascontiguousarray(points[:, 0]), np.ascontiguousarray(points[:, 1])
    inside = natives.is_inside(px, py, exterior)
    for interior in interiors:
        inside &= ~natives.is_inside(px, py, interior)
    dmin2 = np.full(n, np.inf)
    vmin2 = np.full(n, np.inf)                   # distance to the nearest VERTEX (round joins)
    for ring in [exterior] + interiors:
        for k in range(len(ring) - 1):
            dmin2 = np.minimum(dmin2, seg_dist2(ring[k, 0], ring[k, 1], ring[k + 1, 0], ring[k + 1, 1], px, py))
            vmin2 = np.minimum(vmin2, (px - ring[k, 0]) ** 2 + (py - ring[k, 1]) ** 2)
    out = (inside | (dmin2 <= d * d)) if d > 0 else (inside & (dmin2 > d * d))
    if return_exposure:
        lo = abs(d) * np.cos(np.pi / 32.0)
        sliver = (dmin2 == vmin2) & (dmin2 <= d * d) & (dmin2 > lo * lo) & ~inside
        return out, int(sliver.sum())
    return out


def poly_clip(points, poly):
    """clip_utils.poly_clip (utils/clip_utils.py:193-238); poly has .exterior.coords/.interiors.  A polygon
    carrying a pending buffer (``offset`` attribute, no shapely) goes through ``poly_clip_buffered``."""
    if getattr(poly, 'offset', 0.0):
        return poly_clip_buffered(points, poly, float(poly.offset))
    clip_mask = np.zeros((len(points),), dtype=bool)
    exterior = np.array(poly.exterior.coords)
    interiors = [np.array(interior.coords) for interior in poly.interiors]
    if len(exterior) < 3:
        return clip_mask
    bbox_mask = rectangle_clip(points, compute_bounding_box(exterior))
    exterior_mask = natives.is_inside(points[bbox_mask, 0], points[bbox_mask, 1], exterior)
    bbox_inds = np.where(bbox_mask)[0]
    clip_mask[bbox_inds[exterior_mask]] = True
    for interior in interiors:
        if len(interior) < 3:
            continue
        bbox_mask = rectangle_clip(points, compute_bounding_box(interior))
        interior_mask = natives.is_inside(points[bbox_mask, 0], points[bbox_mask, 1], interior)
        bbox_inds = np.where(bbox_mask)[0]
        clip_mask[bbox_inds[interior_mask]] = False
    return clip_mask


# ------------------------------------------------------------------ fusers --------
REFINE_DEFAULTS = {'bottom': 0.02, 'top': 0.5, 'grid_size': 0.4, 'min_comp_size': 50, 'buffer': 0.05,
                   'use_concave': True, 'concave_alpha': 2.0}


def alpha_shape_outer_edges(points, alpha):
    """alpha_shape_utils.get_alpha_shape_edges(points, alpha, only_outer=True) (alpha_shape_utils.py:40-92),
    the literal loop; returns a set of directed (i, j) pairs."""
    from scipy.spatial import Delaunay
    assert points.shape[0] > 3, 'Need at least four points'
    edges = set()

    def add_edge(i, j):
        if (i, j) in edges or (j, i) in edges:
            assert (j, i) in edges, "Can't go twice over same directed edge right?"
            edges.remove((j, i))
            return
        edges.add((i, j))
    for ia, ib, ic in Delaunay(points).simplices:
        pa, pb, pc = points[ia], points[ib], points[ic]
        a = np.sqrt((pa[0] - pb[0]) ** 2 + (pa[1] - pb[1]) ** 2)
        b = np.sqrt((pb[0] - pc[0]) ** 2 + (pb[1] - pc[1]) ** 2)
        c = np.sqrt((pc[0] - pa[0]) ** 2 + (pc[1] - pa[1]) ** 2)
        s = (a + b + c) / 2.0
        int_var = s * (s - a) * (s - b) * (s - c)
        if int_var <= 0:
            continue
        circum_r = a * b * c / (4.0 * np.sqrt(int_var))
        if circum_r < (1 / alpha):
            add_edge(int(ia), int(ib)); add_edge(int(ib), int(ic)); add_edge(int(ic), int(ia))
    return edges


def edges_clip_buffered(points, verts, edges, d):
    """Points inside the region bounded by ``edges`` (pairs into ``verts``; even-odd rule: for the outer edges
    of a triangle complex that is the union of the triangles, i.e. the polygons generate_poly_from_edges
    builds, alpha_shape_utils.py:177-203) or within d of them -- ``poly_clip(points, poly.buffer(d))`` with
    the buffer as a distance.  The crossing rule per edge is clip_utils._point_inside_poly's."""
    px, py = np.ascontiguousarray(points[:, 0]), np.ascontiguousarray(points[:, 1])
    parity = np.zeros(len(px), dtype=bool)
    boundary = np.zeros(len(px), dtype=bool)
    dmin2 = np.full(len(px), np.inf)
    for i, j in edges:
        seg = np.array([verts[i], verts[j], verts[i]])              # a closed two-edge ring crosses twice: use one edge
        xi, yi, xj, yj = verts[i][0], verts[i][1], verts[j][0], verts[j][1]
        dy, dy2 = py - yi, py - yj
        cand = (dy * dy2 <= 0.0) & ((px >= xi) | (px >= xj))
        nonh = cand & ((dy < 0) | (dy2 < 0))
        with np.errstate(invalid='ignore', divide='ignore'):
            F = dy * (xj - xi) / (dy - dy2) + xi
        parity ^= nonh & (px > F)
        boundary |= nonh & (px == F)
        hor = cand & ~nonh & (dy2 == 0) & ((px == xj) | ((dy == 0) & ((px - xi) * (px - xj) <= 0)))
        boundary |= hor
        dmin2 = np.minimum(dmin2, seg_dist2(xi, yi, xj, yj, px, py))
        del seg
    inside = parity | boundary
    return (inside | (dmin2 <= d * d)) if d > 0 else (inside & (dmin2 > d * d)) if d < 0 else inside


def refine_ground_mask(points, points_z, ground_mask, labels, target_label, params):
    """AHNFuser._refine_ground + _refine_layer (fusion/ahn_fuser.py:76-126) on the masked points; the hull
    polygons as regions (outer edges of the alpha complex / convex hull ring), buffer as a distance."""
    from scipy.spatial import ConvexHull
    with np.errstate(invalid='ignore'):
        mask = ((points[:, 2] <= points_z + params['top']) & (points[:, 2] >= points_z - params['bottom']))
    ref_mask = np.zeros((len(points),), dtype=bool)
    if np.count_nonzero(labels[mask] == target_label) == 0:
        return ref_mask
    pts, lab, gmask = points[mask, :], labels[mask], ground_mask[mask]
    add = np.zeros((len(pts),), dtype=bool)
    label_ids = np.where(lab == target_label)[0]
    ground_ids = np.where(gmask)[0]
    point_components = lcc_get_components(pts[label_ids], params['grid_size'], params['min_comp_size'])
    for cc in sorted(set(np.unique(point_components)).difference((-1,))):
        cc_mask = (point_components == cc)
        xy = np.ascontiguousarray(pts[label_ids[cc_mask], 0:2])
        if params['use_concave']:
            edges = alpha_shape_outer_edges(xy, params['concave_alpha'])
        else:
            hull = ConvexHull(xy, qhull_options='QJ').vertices
            edges = {(int(hull[k]), int(hull[(k + 1) % len(hull)])) for k in range(len(hull))}
        if len(edges) < 3:
            continue
        # bounding-box prefilter (only an optimisation here)
        used = np.array(sorted({i for e in edges for i in e}))
        lo, hi = xy[used].min(0) - abs(params['buffer']), xy[used].max(0) + abs(params['buffer'])
        gp = pts[ground_ids]
        near = np.where((gp[:, 0] >= lo[0]) & (gp[:, 0] <= hi[0]) & (gp[:, 1] >= lo[1]) & (gp[:, 1] <= hi[1]))[0]
        if len(near):
            hit = edges_clip_buffered(gp[near], xy, edges, params['buffer'])
            add[ground_ids[near[hit]]] = True
    ref_mask[mask] = add
    return ref_mask


def ahn_fuser_labels(points, labels, mask, ahn_tile, label, target='ground', epsilon=0.2, refine_ground=False,
                     params=None):
    """AHNFuser.get_labels (fusion/ahn_fuser.py:128-173)."""
    surface = 'ground_surface' if target == 'ground' else 'building_surface'
    target_z = interpolate(ahn_tile, points[mask, :], surface)
    label_mask = np.zeros((len(points),), dtype=bool)
    with np.errstate(invalid='ignore'):
        if target == 'ground':
            ground_mask = (np.abs(points[mask, 2] - target_z) < epsilon)
            if refine_ground and np.count_nonzero(ground_mask) > 0:
                p = dict(REFINE_DEFAULTS, **(params or {}))
                tmp_labels = labels[mask].copy()
                tmp_labels[ground_mask] = label
                ref_mask = refine_ground_mask(points[mask], target_z, ground_mask, tmp_labels, UNKNOWN, p)
                ground_mask = ground_mask & ~ref_mask
            label_mask[mask] = ground_mask
        else:
            label_mask[mask] = points[mask, 2] < target_z + epsilon
    labels[label_mask] = label
    return labels


def building_fuser_labels(points, labels, mask, polygons, label, ahn_tile=None, ahn_eps=0.2):
    """BGTBuildingFuser.get_labels (fusion/building_fuser.py:46-102)."""
    label_mask = np.zeros((len(points),), dtype=bool)
    if len(polygons) == 0:
        return labels
    if mask is None:
        mask = np.ones((len(points),), dtype=bool)
    mask_ids = np.where(mask)[0]
    building_mask = np.zeros((len(mask_ids),), dtype=bool)
    for polygon in polygons:
        building_mask = building_mask | poly_clip(points[mask, :], polygon)
    if ahn_tile is not None:
        bld_z = interpolate(ahn_tile, points[mask, :], 'building_surface')
        bld_z_valid = np.isfinite(bld_z)
        ahn_mask = (points[mask_ids[bld_z_valid], 2] <= bld_z[bld_z_valid] + ahn_eps)
        building_mask[bld_z_valid] = building_mask[bld_z_valid] & ahn_mask
    label_mask[mask_ids[building_mask]] = True
    labels[label_mask] = label
    return labels


def road_fuser_labels(points, labels, polygons, label):
    """BGTRoadFuser.get_labels (fusion/road_fuser.py:50-95): domain is labels == GROUND."""
    label_mask = np.zeros(len(points), dtype=bool)
    if len(polygons) == 0:
        return labels
    mask = labels == GROUND
    mask_ids = np.where(mask)[0]
    road_mask = np.zeros((len(mask_ids),), dtype=bool)
    for polygon in polygons:
        road_mask = road_mask | poly_clip(points[mask, :], polygon)
    label_mask[mask_ids[road_mask]] = True
    labels[label_mask] = label
    return labels


# --------------------------------------------------------------------- LCC --------
def get_octree_level(points, grid_size):
    """math_utils.get_octree_level (utils/math_utils.py:21-33)."""
    dims = np.zeros((points.shape[1],))
    for d in range(points.shape[1]):
        dims[d] = np.max(points[:, d]) - np.min(points[:, d])
    max_dim = np.max(dims)
    if max_dim < 0.001:
        return 0
    octree_level = np.rint(-np.log(grid_size / max_dim) / (np.log(2)))
    if octree_level > 0:
        return int(np.int64(octree_level))
    return 1


def lcc_point_components(points, mask, grid_size, min_component_size):
    """_convert_input_cloud + _label_connected_comp
    (region_growing/label_connected_comp.py:53-97).  Returns float ids over masked points,
    -1 for components below min_component_size."""
    xs = (points[mask, 0]).astype(np.float32)
    ys = (points[mask, 1]).astype(np.float32)
    if points.shape[1] == 2:
        zs = np.zeros(xs.shape).astype(np.float32)
    else:
        zs = (points[mask, 2]).astype(np.float32)
    octree_level = get_octree_level(points, grid_size)
    if len(xs) == 0:
        return np.zeros((0,), dtype=np.float32), octree_level
    point_components = np.around(natives.label_connected_components(xs, ys, zs, octree_level))
    if min_component_size > 1:
        cc_labels, counts = np.unique(point_components, return_counts=True)
        cc_labels_filtered = cc_labels[counts < min_component_size]
        point_components[np.isin(point_components, cc_labels_filtered)] = -1
    return point_components, octree_level


def lcc_fill_components(point_components, mask, point_labels, label, min_component_size, threshold,
                        literal=False):
    """_fill_components (label_connected_comp.py:99-135).  `literal` runs the reference's
    per-component loop; the default computes the same thing with bincount."""
    mask_indices = np.where(mask)[0]
    label_mask = np.zeros(len(mask), dtype=bool)
    if literal:
        cc_labels = set(np.unique(point_components)).difference((-1,))
        for cc in cc_labels:
            cc_mask = (point_components == cc)
            cc_size = np.count_nonzero(cc_mask)
            if cc_size < min_component_size:
                continue
            seed_count = np.count_nonzero(point_labels[mask_indices[cc_mask]] == label)
            if (float(seed_count) / cc_size) > threshold:
                label_mask[mask_indices[cc_mask]] = True
        return label_mask
    valid = point_components >= 0
    ids = point_components[valid].astype(np.int64)
    if len(ids) == 0:
        return label_mask
    sizes = np.bincount(ids)
    seeds = np.bincount(ids, weights=(point_labels[mask_indices[valid]] == label).astype(np.float64))
    with np.errstate(invalid='ignore', divide='ignore'):
        accept = (sizes >= min_component_size) & ((seeds.astype(np.float64) / sizes) > threshold)
    label_mask[mask_indices[valid]] = accept[ids]
    return label_mask


def lcc_get_label_mask(points, labels, mask, label, exclude_labels=(), grid_size=0.1,
                       min_component_size=100, threshold=0.1, literal=False):
    """LabelConnectedComp.get_label_mask (label_connected_comp.py:137-181)."""
    if exclude_labels:
        m = np.ones((len(points),), dtype=bool)
        for exclude_label in exclude_labels:
            m = m & (labels != exclude_label)
    elif mask is None:
        m = np.ones((len(points),), dtype=bool)
    else:
        m = mask.copy()
        m[labels == label] = True
    comps, _ = lcc_point_components(points, m, grid_size, min_component_size)
    return lcc_fill_components(comps, m, labels, label, min_component_size, threshold, literal)


def lcc_get_components(points, grid_size, min_component_size):
    """LabelConnectedComp.get_components with no exclusions (label_connected_comp.py:207-231)."""
    m = np.ones((len(points),), dtype=bool)
    return lcc_point_components(points, m, grid_size, min_component_size)[0]


def layer_lcc_labels(points, labels, mask, ahn_tile, label, params, reset_noise=False):
    """LayerLCC.get_labels (region_growing/layer_lcc.py:95-145)."""
    mask_copy = mask.copy()
    mask_copy[labels == label] = True
    if reset_noise:
        noise_mask = labels == NOISE
        mask_copy[noise_mask] = True
    points_z = interpolate(ahn_tile, points[mask_copy], 'ground_surface')
    label_mask = np.zeros((len(points),), dtype=bool)
    layer_mask = np.zeros((np.count_nonzero(mask_copy),), dtype=bool)
    labels_copy = labels[mask_copy].copy()
    pts = points[mask_copy, :]
    for layer in params:
        p = {'bottom': -np.inf, 'top': np.inf, 'grid_size': 0.1, 'min_comp_size': 100, 'threshold': 0.5}
        p.update(layer)
        # _filter_layer (layer_lcc.py:70-93)
        with np.errstate(invalid='ignore'):
            ids = np.where((pts[:, 2] > points_z + p['bottom']) & (pts[:, 2] <= points_z + p['top']))[0]
        m = np.zeros((len(pts),), dtype=bool)
        if len(ids) > 0 and np.count_nonzero(labels_copy[ids] == label) > 0:
            lcc_mask = lcc_get_label_mask(pts[ids], labels_copy[ids].copy(), None, label,
                                          grid_size=p['grid_size'], min_component_size=p['min_comp_size'],
                                          threshold=p['threshold'])
            m[ids[lcc_mask]] = True
        layer_mask = layer_mask | m
        labels_copy[layer_mask] = label
    label_mask[mask_copy] = layer_mask
    if reset_noise:
        label_mask = label_mask & (mask | noise_mask)
    else:
        label_mask = label_mask & mask
    labels[label_mask] = label
    return labels


def noise_filter_labels(points, labels, mask, ahn_tile, label, epsilon=0.2, grid_size=0.1,
                        min_component_size=100):
    """NoiseFilter.get_labels (fusion/noise_filter.py:43-84)."""
    point_components = lcc_get_components(points[mask], grid_size, min_component_size)
    cc_mask = point_components == -1
    target_z = interpolate(ahn_tile, points[mask], 'ground_surface')
    with np.errstate(invalid='ignore'):
        ground_mask = (points[mask, 2] - target_z) < -epsilon
    label_mask = np.zeros((len(points),), dtype=bool)
    label_mask[mask] = cc_mask | ground_mask
    labels[label_mask] = label
    return labels


# ------------------------------------------------ object fusers (clients of LCC) ----
def minimum_bounding_rectangle(points):
    """math_utils.minimum_bounding_rectangle (utils/math_utils.py:62-125): qhull convex hull,
    candidate orientations = directions of the OPEN chain of hull edges folded into [0, pi/2),
    smallest-area box.  Returns (corners, width, length, centre)."""
    from scipy.spatial import ConvexHull
    pi2 = np.pi / 2.
    hull_points = points[ConvexHull(points).vertices]
    edges = hull_points[1:] - hull_points[:-1]
    angles = np.unique(np.abs(np.mod(np.arctan2(edges[:, 1], edges[:, 0]), pi2)))
    rotations = np.vstack([np.cos(angles), np.cos(angles - pi2), np.cos(angles + pi2), np.cos(angles)]).T
    rotations = rotations.reshape((-1, 2, 2))
    rot_points = np.dot(rotations, hull_points.T)
    min_x = np.nanmin(rot_points[:, 0], axis=1); max_x = np.nanmax(rot_points[:, 0], axis=1)
    min_y = np.nanmin(rot_points[:, 1], axis=1); max_y = np.nanmax(rot_points[:, 1], axis=1)
    best_idx = np.argmin((max_x - min_x) * (max_y - min_y))
    x1, x2, y1, y2 = max_x[best_idx], min_x[best_idx], max_y[best_idx], min_y[best_idx]
    r = rotations[best_idx]
    center_point = np.dot([(x1 + x2) / 2, (y1 + y2) / 2], r)
    rect = np.zeros((4, 2))
    rect[0] = np.dot([x1, y2], r); rect[1] = np.dot([x2, y2], r)
    rect[2] = np.dot([x2, y1], r); rect[3] = np.dot([x1, y1], r)
    dims = [(x1 - x2), (y1 - y2)]
    return rect, min(dims), max(dims), center_point


def _half_plane_cut(poly, a, b):
    """Vertices of the closed polygon ``poly`` (list of (x, y), not repeated) on the left of a->b."""
    def side(p):
        return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
    out = []
    for i in range(len(poly)):
        p, q = poly[i], poly[(i + 1) % len(poly)]
        sp, sq = side(p), side(q)
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def _shoelace(poly):
    return 0.5 * sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
                     for i in range(len(poly))) if len(poly) >= 3 else 0.0


def rect_overlap_percentage(rect_ring, poly):
    """(p1.intersection(p2).area / p1.area) * 100 for a convex p1 (car_fuser.py:74-79); stands in for
    GEOS: successive half-plane cuts of p2's rings by the rectangle's edges, shoelace areas."""
    rect = [tuple(p) for p in np.asarray(rect_ring)[:-1]]
    if _shoelace(rect) < 0:
        rect = rect[::-1]
    area = _shoelace(rect)
    if area == 0:
        return 0.0

    def clipped_area(coords):
        ring = [tuple(c[:2]) for c in list(coords)[:-1]]
        for i in range(len(rect)):
            if len(ring) < 3:
                return 0.0
            ring = _half_plane_cut(ring, rect[i], rect[(i + 1) % len(rect)])
        return abs(_shoelace(ring))
    inter = clipped_area(poly.exterior.coords) - sum(clipped_area(h.coords) for h in poly.interiors)
    return inter / area * 100


def _object_candidates(points_m, ground_z, point_components, min_height, max_height):
    """Shared head of the two per-cluster loops (car_fuser.py:53-69)."""
    for cc in sorted(set(np.unique(point_components)).difference((-1,))):
        cc_mask = (point_components == cc)
        target_z = ground_z[cc_mask]
        valid_values = target_z[np.isfinite(target_z)]
        if valid_values.size != 0:
            cc_z = np.mean(valid_values)
            min_z = cc_z + min_height
            max_z = cc_z + max_height
            cluster_height = np.amax(points_m[cc_mask][:, 2])
            if min_z <= cluster_height <= max_z:
                yield cc_mask, cc_z, max_z


def car_fuser_labels(points, labels, mask, ahn_tile, road_polygons, label, grid_size=0.05,
                     min_component_size=5000, overlap_perc=20, params=None):
    """CarFuser.get_labels + _label_car_like_components (fusion/car_fuser.py:45-137)."""
    if len(road_polygons) == 0:
        return labels
    p = params
    pm = points[mask]
    ground_z = interpolate(ahn_tile, pm, 'ground_surface')
    point_components = lcc_get_components(pm, grid_size, min_component_size)
    car_mask = np.zeros(len(pm), dtype=bool)
    for cc_mask, cc_z, max_z in _object_candidates(pm, ground_z, point_components, p['min_height'], p['max_height']):
        mbrect, mbr_width, mbr_length, _ = minimum_bounding_rectangle(pm[cc_mask][:, :2])
        poly = np.vstack((mbrect, mbrect[0]))
        if p['min_width'] < mbr_width < p['max_width'] and p['min_length'] < mbr_length < p['max_length']:
            for p2 in road_polygons:
                if rect_overlap_percentage(poly, p2) > overlap_perc:
                    box = type('P', (), {'exterior': type('R', (), {'coords': poly})(), 'interiors': []})()
                    car_mask = car_mask | (poly_clip(pm, box) & ((pm[:, 2] <= max_z) & (pm[:, 2] >= cc_z)))
                    break
    label_mask = np.zeros((len(points),), dtype=bool)
    label_mask[mask] = car_mask
    labels[label_mask] = label
    return labels


def street_furniture_labels(points, labels, mask, ahn_tile, bgt_points, label, grid_size=0.05,
                            min_component_size=1500, max_dist=1., params=None):
    """BGTStreetFurnitureFuser.get_labels + _label_street_furniture_like_components
    (fusion/street_furniture_fuser.py:46-137)."""
    if len(bgt_points) == 0:
        return labels
    p = params
    pm = points[mask]
    ground_z = interpolate(ahn_tile, pm, 'ground_surface')
    point_components = lcc_get_components(pm, grid_size, min_component_size)
    sf_mask = np.zeros(len(pm), dtype=bool)
    for cc_mask, cc_z, max_z in _object_candidates(pm, ground_z, point_components, p['min_height'], p['max_height']):
        _, mbr_width, mbr_length, center_point = minimum_bounding_rectangle(pm[cc_mask][:, :2])
        if p['min_width'] < mbr_width < p['max_width'] and p['min_length'] < mbr_length < p['max_length']:
            for bgt_point in bgt_points:
                if np.linalg.norm(np.asarray(bgt_point) - center_point) <= max_dist:
                    sf_mask[cc_mask] = True
                    break
    label_mask = np.zeros((len(points),), dtype=bool)
    label_mask[mask] = sf_mask
    labels[label_mask] = label
    return labels


# ---------------------------------------------------------------------- RG --------
def vector_angle(u, v):
    """math_utils.vector_angle (utils/math_utils.py:9-18)."""
    c = np.dot(u / np.linalg.norm(u), v / np.linalg.norm(v))
    clip = np.minimum(1, np.maximum(c, -1))
    return np.rad2deg(np.arccos(clip))


def rg_setup(coords, max_nn=30, grow_region_knn=15, grow_region_radius=0.2, threshold_curve=1.0):
    """Everything _region_growing derives from the cloud alone: kNN rows (KDTreeFlann),
    normals (estimate_normals, region_growing.py:89-92) and the curvature seed test of every
    point (region_growing.py:59-73,128-134)."""
    knn = natives.knn(coords, grow_region_knn)
    normals = natives.normals_hybrid(coords, grow_region_radius, max_nn)
    cov = natives.cov6_to_matrix(natives.cov_knn(coords, knn))
    with np.errstate(all='ignore'):
        eig_val = np.linalg.eig(cov)[0]
        curvature = eig_val[:, 0] / eig_val.sum(axis=1)
        can_seed = curvature < threshold_curve
    return knn, normals, can_seed


def radius_rows(coords, radius):
    """search_radius_vector_3d of every point (region_growing.py:104-106): lists of indices with squared
    distance < radius^2 in ascending (squared distance, index) order."""
    from scipy.spatial import cKDTree
    tree = cKDTree(coords)
    rows = []
    for i, cand in enumerate(tree.query_ball_point(coords, radius * (1.0 + 1e-9) + 1e-12)):
        cand = np.asarray(cand, dtype=np.int64)
        d = coords[cand] - coords[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 < radius * radius
        cand, d2 = cand[keep], d2[keep]
        rows.append(cand[np.lexsort((cand, d2))])
    return rows


def curvature_of_rows(coords, rows, threshold_curve):
    """_compute_point_curvature per point (region_growing.py:59-73): Open3D's single-pass cumulant covariance of
    the neighbours IN THEIR ORDER, LAPACK eig (unsorted), eig_val[0] / sum < threshold."""
    out = np.zeros(len(rows), dtype=bool)
    with np.errstate(all='ignore'):
        for i, row in enumerate(rows):
            p = coords[row]
            n = float(len(row))
            s1 = np.cumsum(p, axis=0)[-1] / n
            prod = np.stack([p[:, 0] * p[:, 0], p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 1], p[:, 1] * p[:, 2],
                             p[:, 2] * p[:, 2]], axis=1)
            s2 = np.cumsum(prod, axis=0)[-1] / n
            cov = np.array([[s2[0] - s1[0] * s1[0], s2[1] - s1[0] * s1[1], s2[2] - s1[0] * s1[2]],
                            [s2[1] - s1[0] * s1[1], s2[3] - s1[1] * s1[1], s2[4] - s1[1] * s1[2]],
                            [s2[2] - s1[0] * s1[2], s2[4] - s1[1] * s1[2], s2[5] - s1[2] * s1[2]]])
            eig_val = np.linalg.eig(cov)[0]
            out[i] = (eig_val[0] / eig_val.sum()) < threshold_curve
    return out


def rg_literal_loop_rows(rows, normals, seed_ids, can_seed, threshold_angle):
    """The `while` loop of _region_growing with variable-length neighbour lists (method='radius')."""
    list_of_seed_ids = list(seed_ids)
    region = list(seed_ids)
    processed = np.full(len(rows), False)
    processed[list_of_seed_ids] = True
    idx = 0
    while idx < len(list_of_seed_ids):
        seed = list_of_seed_ids[idx]
        seed_normal = normals[seed]
        for neighbor_id in rows[seed][1:]:
            if processed[neighbor_id]:
                continue
            if vector_angle(seed_normal, normals[neighbor_id]) < threshold_angle:
                region.append(neighbor_id)
                processed[neighbor_id] = True
                if can_seed[neighbor_id]:
                    list_of_seed_ids.append(neighbor_id)
        idx = idx + 1
    out = np.zeros(len(rows), dtype=bool)
    out[region] = True
    return out


def rg_literal_loop(knn, normals, seed_ids, can_seed, threshold_angle):
    """The Python `while` loop of _region_growing, literally (region_growing.py:94-139)."""
    list_of_seed_ids = list(seed_ids)
    region = list(seed_ids)
    processed = np.full(len(knn), False)
    processed[list_of_seed_ids] = True
    idx = 0
    while idx < len(list_of_seed_ids):
        seed = list_of_seed_ids[idx]
        seed_normal = normals[seed]
        row = knn[seed]
        k = int(np.count_nonzero(row >= 0))
        neighbor_idx = row[1:k]
        for neighbor_id in neighbor_idx:
            if processed[neighbor_id]:
                continue
            current_angle = vector_angle(seed_normal, normals[neighbor_id])
            if current_angle < threshold_angle:
                region.append(neighbor_id)
                processed[neighbor_id] = True
                if can_seed[neighbor_id]:
                    list_of_seed_ids.append(neighbor_id)
        idx = idx + 1
    out = np.zeros(len(knn), dtype=bool)
    out[region] = True
    return out


def region_growing_labels(points, labels, label, exclude_labels=(), threshold_angle=20,
                          threshold_curve=1.0, max_nn=30, grow_region_knn=15, grow_region_radius=0.2,
                          literal=False, angle_tol=0.0, return_near=False, method='knn'):
    """RegionGrowing.get_labels (region_growing/region_growing.py:143-170); method='radius' is what
    _region_growing('radius') computes (:104-106, only reachable by calling it directly)."""
    mask = np.ones((len(labels),), dtype=bool)
    for exclude_label in exclude_labels:
        mask = mask & (labels != exclude_label)
    seed_ids = np.where(labels[mask] == label)[0]
    mask_indices = np.where(mask)[0]
    label_mask = np.zeros(len(mask), dtype=bool)
    near_full = np.zeros(len(mask), dtype=bool)
    if len(mask_indices) > 0 and len(seed_ids) > 0:
        coords = np.ascontiguousarray(points[mask, :3])
        if method == 'radius':
            rows = radius_rows(coords, grow_region_radius)
            normals = natives.normals_hybrid(coords, grow_region_radius, max_nn)
            can_seed = curvature_of_rows(coords, rows, threshold_curve)
            region = rg_literal_loop_rows(rows, normals, seed_ids.tolist(), can_seed, threshold_angle)
            label_mask[mask_indices[region]] = True
            labels[label_mask] = label
            return (labels, near_full) if return_near else labels
        knn, normals, can_seed = rg_setup(coords, max_nn, grow_region_knn, grow_region_radius,
                                          threshold_curve)
        if literal:
            region = rg_literal_loop(knn, normals, seed_ids.tolist(), can_seed, threshold_angle)
            near = np.zeros(len(coords), dtype=bool)
        else:
            seed = np.zeros(len(coords), dtype=np.uint8)
            seed[seed_ids] = 1
            region, near = natives.region_grow(knn, normals, seed, can_seed, threshold_angle, angle_tol)
        label_mask[mask_indices[region]] = True
        near_full[mask_indices[near]] = True
    labels[label_mask] = label
    if return_near:
        return labels, near_full
    return labels


# ---------------------------------------------------------------- pipeline --------
def process_cloud(processors, points, labels):
    """Pipeline.process_cloud (pipeline.py:63-97): each processor is a callable
    (points, labels, mask) -> labels and sees mask = (labels == 0)."""
    for proc in processors:
        mask = (labels == 0)
        labels = proc(points, labels, mask)
    return labels


def label_histogram(labels):
    """analysis_tools.get_label_stats counts (analysis/analysis_tools.py:8-18) as int64[256]."""
    vals, counts = np.unique(labels, return_counts=True)
    hist = np.zeros(256, dtype=np.int64)
    for v, c in zip(vals, counts):
        hist[int(v) if 0 <= int(v) <= 255 else 255] += c
    return hist


def canonical_components(comp):
    """Relabel component ids by the minimum point index of each component; -1 stays -1."""
    comp = np.asarray(comp)
    out = np.full(comp.shape, -1, dtype=np.int64)
    valid = comp >= 0
    if not valid.any():
        return out
    ids = comp[valid].astype(np.int64)
    idx = np.where(valid)[0]
    first = np.full(ids.max() + 1, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, ids, idx)
    out[valid] = first[ids]
    return out
