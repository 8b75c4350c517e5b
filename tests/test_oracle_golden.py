"""CPU: the oracle restatement against the golden vectors generated from the REAL reference
code (oracle/make_golden.py).  This is what pins the oracle on machines without
/root/reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, from_mm, unpack
from oracle import natives, reference_py as R
from urban_pointcloud_processing_b200 import synthetic
from urban_pointcloud_processing_b200.utils.bgt_utils import RingPolygon

PIP_CASES = ['bgt_buildings', 'bgt_roads', 'synthetic_buildings', 'synthetic_roads', 'handmade', 'lattice']


def polygons_from_golden(g, name):
    xy, off, poly, hole = g[f'{name}__ring_xy'], g[f'{name}__ring_off'], g[f'{name}__ring_poly'], g[f'{name}__ring_hole']
    polys = {}
    for r in range(len(poly)):
        ring = xy[off[r]:off[r + 1]]
        polys.setdefault(int(poly[r]), [None, []])
        if hole[r]:
            polys[int(poly[r])][1].append(ring)
        else:
            polys[int(poly[r])][0] = ring
    out = []
    for p in sorted(polys):
        e, h = polys[p]
        rp = RingPolygon.__new__(RingPolygon)
        rp.exterior = type('R', (), {'coords': [tuple(c) for c in e]})()
        rp.interiors = [type('R', (), {'coords': [tuple(c) for c in hh]})() for hh in h]
        out.append(rp)
    return out


@pytest.mark.parametrize('name', PIP_CASES)
def test_pip_oracle_matches_reference_golden(golden, name):
    g = golden('pip_golden.npz')
    pts = g[f'{name}__points']
    expected = unpack(g[f'{name}__expected'], int(g[f'{name}__n']))
    got = np.zeros(len(pts), dtype=bool)
    for poly in polygons_from_golden(g, name):
        got |= R.poly_clip(pts, poly)
    assert np.array_equal(got, expected)


def test_pip_known_answers():
    # unit square: boundary and vertices count as inside (clip_utils.py:150-160)
    sq = RingPolygon([(0, 0), (1, 0), (1, 1), (0, 1), (0, 0)])
    pts = np.array([[0.5, 0.5], [0, 0], [1, 0.5], [0.5, 0], [0.5, 1], [1, 1], [0, 0.5], [2, 0.5], [-1, 0.5], [0.5, 2]])
    assert R.poly_clip(pts, sq).tolist() == [True] * 7 + [False] * 3
    # hole boundary is excluded
    ph = RingPolygon([(0, 0), (4, 0), (4, 4), (0, 4), (0, 0)], [[(1, 1), (1, 3), (3, 3), (3, 1), (1, 1)]])
    pts = np.array([[2, 2], [1, 2], [0.5, 2], [3, 3], [3.5, 3.5]], dtype=float)
    assert R.poly_clip(pts, ph).tolist() == [False, False, True, False, True]


def test_raster_oracle_matches_reference_golden(golden):
    g = golden('raster_golden.npz')
    ahn = np.load(os.path.join(GOLDEN, 'demo', 'ahn_2386_9702.npz'))
    for surface, key in (('ground_surface', 'ground'), ('building_surface', 'building')):
        got = R.fast_grid_lookup(ahn['x'], ahn['y'], ahn[key].astype(float), g['points'])
        assert np.array_equal(got, g[surface], equal_nan=True)


def test_scalars_oracle_matches_reference_golden(golden):
    g = golden('scalars_golden.npz')
    for e, gs, lvl in zip(g['extent'], g['grid_size'], g['level']):
        p = np.zeros((2, 3)); p[1, 0] = e; p[1, 1] = e * 0.3
        assert R.get_octree_level(p, gs) == lvl
    ang = np.array([R.vector_angle(a, b) for a, b in zip(g['u'], g['v'])])
    np.testing.assert_allclose(ang, g['angle'], rtol=0, atol=1e-9)


def _pipeline_inputs(g):
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']))
    points = np.asfortranarray(from_mm(g['points_mm']))
    assert np.array_equal(points, tile['points']), 'synthetic generator drifted from the golden fixture'
    return tile, points


def test_pipeline_oracle_matches_reference_golden(golden):
    """Every stage of the reference control flow (run over the restated natives when the
    fixture was made) is reproduced by oracle.reference_py on this machine."""
    g = golden('pipeline_golden.npz')
    tile, points = _pipeline_inputs(g)
    ahn = tile['ahn_tile']
    # the csv round trip of the fixture drops interior rings
    blds = [RingPolygon(np.array(p.exterior.coords)) for p in tile['building_polygons']]
    roads = tile['road_polygons']
    labels = np.zeros(len(points), dtype=np.uint16)
    labels = R.ahn_fuser_labels(points, labels, labels == 0, ahn, 9, 'ground', 0.2)
    assert np.array_equal(labels, g['labels_after_ahn_ground'])
    labels = R.road_fuser_labels(points, labels, roads, 1)
    assert np.array_equal(labels, g['labels_after_bgt_road'])
    labels = R.building_fuser_labels(points, labels, labels == 0, blds, 10, ahn, 0.2)
    assert np.array_equal(labels, g['labels_after_bgt_building'])
    labels = R.layer_lcc_labels(points, labels, labels == 0, ahn, 10,
                                [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                                 {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}])
    assert np.array_equal(labels, g['labels_after_layer_lcc'])
    labels[g['car_seed_ids']] = 40
    lm = R.lcc_get_label_mask(points, labels, labels == 0, 40, grid_size=0.1, min_component_size=30, threshold=0.02)
    labels[lm] = 40
    assert np.array_equal(labels, g['labels_after_lcc_car'])
    labels = R.noise_filter_labels(points, labels, labels == 0, ahn, 99, 0.2, 0.1, 10)
    assert np.array_equal(labels, g['labels_after_noise'])
    labels = R.region_growing_labels(points, labels, 10, exclude_labels=(9, 1))
    assert np.array_equal(labels, g['labels_after_region_growing'])
    comps = R.canonical_components(R.lcc_get_components(points[tile['kind'] >= 2], 0.2, 20))
    assert np.array_equal(comps, g['components_kind_ge2'])


def test_fill_components_literal_equals_vectorised():
    rng = np.random.default_rng(5)
    tile = synthetic.make_tile(20000, seed=3)
    pts = tile['points']
    labels = np.where(rng.random(len(pts)) < 0.05, 40, 0).astype(np.uint16)
    mask = rng.random(len(pts)) < 0.8
    a = R.lcc_get_label_mask(pts, labels, mask, 40, grid_size=0.4, min_component_size=5, threshold=0.04, literal=True)
    b = R.lcc_get_label_mask(pts, labels, mask, 40, grid_size=0.4, min_component_size=5, threshold=0.04, literal=False)
    assert np.array_equal(a, b) and a.any()


def test_region_growing_literal_loop_equals_queue_native():
    """Order independence: the literal Python while-loop and the C queue
    loop reach the same fixpoint."""
    tile = synthetic.make_tile(12000, seed=11, facade_fraction=0.7)
    pts = tile['points']
    labels = np.zeros(len(pts), dtype=np.uint16)
    labels = R.ahn_fuser_labels(pts, labels, labels == 0, tile['ahn_tile'], 9, 'ground', 0.2)
    labels = R.building_fuser_labels(pts, labels, labels == 0, tile['building_polygons'], 10, tile['ahn_tile'], 0.2)
    a = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), literal=True)
    b = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), literal=False)
    assert np.array_equal(a, b)
    assert np.count_nonzero(a == 10) > np.count_nonzero(labels == 10)


def test_knn_native_against_bruteforce():
    rng = np.random.default_rng(0)
    pts = np.round(rng.random((1500, 3)) * 4, 2)          # coarse lattice => many exact ties
    pts = np.vstack([pts, pts[:40]])                       # and exact duplicates
    idx, d2 = natives.knn(pts, 9, want_d2=True)
    d = pts[:, None, :] - pts[None, :, :]
    dd = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    order = np.lexsort((np.broadcast_to(np.arange(len(pts)), dd.shape), dd), axis=1)[:, :9]
    assert np.array_equal(idx, order)
    assert np.array_equal(d2, np.take_along_axis(dd, order, axis=1))


# ------------------------------------------------- object fusers ("next" row 2) ---
OBJ_CAR_PARAMS = {'min_height': 1.2, 'max_height': 2.4, 'min_width': 1.4, 'max_width': 2.4, 'min_length': 3.0,
                  'max_length': 6.0}
OBJ_POLE_PARAMS = {'min_height': 1.5, 'max_height': 9.0, 'min_width': 0.1, 'max_width': 1.2, 'min_length': 0.1,
                   'max_length': 1.2}


def objects_tile(g):
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']), n_cars=12)
    assert np.array_equal(tile['points'], from_mm(g['points_mm']))
    poles = [('lichtmast', x, y) for x, y in tile['pole_xy']][::2] + [('boom', x, y) for x, y in tile['pole_xy']][1::2]
    return tile, poles


def test_object_fusers_oracle_matches_reference_golden(golden):
    """CarFuser / BGTStreetFurnitureFuser of the REAL reference (objects_golden.npz) vs the restatement."""
    g = golden('objects_golden.npz')
    tile, poles = objects_tile(g)
    points, ahn = tile['points'], tile['ahn_tile']
    labels = np.zeros(len(points), dtype=np.uint16)
    labels = R.ahn_fuser_labels(points, labels, labels == 0, ahn, 9, 'ground', 0.2)
    assert np.array_equal(labels, g['labels_after_ground'])
    labels = R.car_fuser_labels(points, labels, labels == 0, ahn, tile['road_polygons'][:1], 40, 0.1, 300, 20,
                                OBJ_CAR_PARAMS)
    assert np.array_equal(labels, g['labels_after_car']) and (labels == 40).sum() > 1000
    lights = [(x, y) for t, x, y in poles if t == 'lichtmast']
    labels = R.street_furniture_labels(points, labels, labels == 0, ahn, lights, 60, 0.1, 100, 1., OBJ_POLE_PARAMS)
    assert np.array_equal(labels, g['labels_after_street_light']) and (labels == 60).sum() > 300


def test_minimum_bounding_rectangle_and_overlap_host_logic(golden):
    """The product's host-side object geometry: bit-identical to the reference's
    minimum_bounding_rectangle (golden vectors), overlap percentage equal to the oracle's
    independent clipping and to closed-form cases."""
    from urban_pointcloud_processing_b200.utils import math_utils, object_utils
    from urban_pointcloud_processing_b200.utils.bgt_utils import RingPolygon
    g = golden('objects_golden.npz')
    for k in range(6):
        rect, hull, w, l, ctr = math_utils.minimum_bounding_rectangle(g[f'mbr_in_{k}'])
        assert np.array_equal(np.concatenate([rect.ravel(), [w, l], ctr]), g[f'mbr_out_{k}'])
        r2, w2, l2, c2 = R.minimum_bounding_rectangle(g[f'mbr_in_{k}'])
        assert np.array_equal(rect, r2) and (w, l) == (w2, l2)
    unit = np.array([[0., 0.], [2., 0.], [2., 1.], [0., 1.], [0., 0.]])
    half = RingPolygon([(1., -5.), (9., -5.), (9., 5.), (1., 5.)])
    assert object_utils.overlap_percentage(unit, half) == 50.0 == R.rect_overlap_percentage(unit, half)
    assert object_utils.overlap_percentage(unit[::-1], half) == 50.0
    holed = RingPolygon([(-1., -1.), (3., -1.), (3., 2.), (-1., 2.)], [[(0.5, 0.25), (1.5, 0.25), (1.5, 0.75), (0.5, 0.75)]])
    assert abs(object_utils.overlap_percentage(unit, holed) - 75.0) < 1e-12
    away = RingPolygon([(5., 5.), (6., 5.), (6., 6.)])
    assert object_utils.overlap_percentage(unit, away) == 0.0 == R.rect_overlap_percentage(unit, away)
    rng = np.random.default_rng(3)
    tile = synthetic.make_tile(1000, seed=1)
    for _ in range(50):                                      # rotated rectangles against the concave road strips
        c, a = rng.uniform(10, 40, 2) + np.array(tile['origin']), rng.uniform(0, np.pi)
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        box = (np.array([[-2.2, -0.9], [2.2, -0.9], [2.2, 0.9], [-2.2, 0.9], [-2.2, -0.9]]) @ rot.T) + c
        for road in tile['road_polygons']:
            assert abs(object_utils.overlap_percentage(box, road) - R.rect_overlap_percentage(box, road)) < 1e-9


# ---------- natives with no reference golden ("parity unpinned"): independent cross-checks ----------
def test_cc_native_connectivity_against_scipy_ndimage():
    """The graph half of the cccorelib restatement: components of occupied cells under 26-connectivity must
    be what scipy.ndimage.label finds on the same voxel grid (the cell arithmetic itself is DgmOctree's and
    stays unpinned; here it is only re-used to rasterise)."""
    from scipy import ndimage
    rng = np.random.default_rng(12)
    blobs = [rng.normal(c, s, (n, 3)) for c, s, n in (((2, 2, 2), 0.3, 1500), ((6, 2, 2), 0.5, 2500),
                                                      ((4, 7, 3), 0.2, 800), ((8, 8, 8), 0.05, 300))]
    pts = np.vstack(blobs + [rng.random((400, 3)) * 10]).astype(np.float32)
    for level in (4, 6, 8):
        got = natives.label_connected_components(pts[:, 0], pts[:, 1], pts[:, 2], level).astype(np.int64)
        mn = natives.octree_params(pts[:, 0], pts[:, 1], pts[:, 2])
        cs = mn[3]
        pos = np.empty((len(pts), 3), dtype=np.int64)
        for d in range(3):                                          # float32 arithmetic, clamp, then >> (21 - level)
            q = ((pts[:, d] - mn[d]) / cs).astype(np.float32)
            pos[:, d] = np.clip(q.astype(np.int64), 0, (1 << 21) - 1) >> (21 - level)
        grid = np.zeros((1 << level,) * 3, dtype=bool) if level <= 6 else None
        if grid is not None:
            grid[pos[:, 0], pos[:, 1], pos[:, 2]] = True
            lab, k = ndimage.label(grid, structure=np.ones((3, 3, 3)))
            want = lab[pos[:, 0], pos[:, 1], pos[:, 2]]
        else:                                                       # level 8: sparse graph instead of a dense volume
            from scipy.sparse import coo_matrix
            from scipy.sparse.csgraph import connected_components
            cells, inv = np.unique(pos, axis=0, return_inverse=True)
            index = {tuple(c): i for i, c in enumerate(cells)}
            rows, cols = [], []
            for i, c in enumerate(cells):
                for off in np.ndindex(3, 3, 3):
                    j = index.get((c[0] + off[0] - 1, c[1] + off[1] - 1, c[2] + off[2] - 1))
                    if j is not None and j != i:
                        rows.append(i); cols.append(j)
            k, comp = connected_components(coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(len(cells),) * 2))
            want = comp[inv.ravel()] + 1
        assert got.min() == 1 and got.max() == k
        assert np.array_equal(R.canonical_components(got.astype(np.float64)), R.canonical_components(want.astype(np.float64)))


def test_normals_and_covariance_natives_against_lapack():
    """The Open3D restatement (hybrid search, cumulant covariance, FastEigen3x3 normal) against an independent
    formulation: cKDTree neighbourhoods, two-pass covariance, LAPACK eigh.  Normals agree up to sign within the
    perturbation bound of the single-pass covariance wherever the two smallest eigenvalues are separated; points with fewer than 3 neighbours inside the
    radius get (0, 0, 1)."""
    from scipy.spatial import cKDTree
    tile = synthetic.make_tile(30000, seed=17, facade_fraction=0.6)
    pts = tile['points'][tile['kind'] != 0]
    radius, max_nn, knn = 0.2, 30, 15
    nrm = natives.normals_hybrid(pts, radius, max_nn)
    tree = cKDTree(pts)
    dist, idx = tree.query(pts, k=max_nn)
    inside = dist < radius                                           # strict, as nanoflann's radius test on d2 < r2
    checked = defaults = 0
    for i in range(0, len(pts), 2):
        nb = pts[idx[i][inside[i]]]
        if len(nb) < 3:
            assert np.array_equal(nrm[i], [0.0, 0.0, 1.0])
            defaults += 1
            continue
        c = np.cov((nb - nb.mean(axis=0)).T, bias=True)
        w, v = np.linalg.eigh(c)
        gap = w[1] - w[0]
        if gap < 1e-3 * max(w[2], 1e-30):                            # smallest eigenvector ill-defined: skip
            continue
        # Open3D accumulates the cumulants on the unshifted RD coordinates (y ~ 4.9e5): its covariance carries an
        # absolute error ~ eps * y^2 ~ 1e-4, which turns the eigenvector by at most ~ error / gap
        sin_bound = 2e-4 / gap
        assert 1.0 - abs(np.dot(v[:, 0], nrm[i])) < max(1e-9, sin_bound * sin_bound), (i, gap)
        checked += 1
    assert checked > 800 and defaults > 5
    rows = natives.knn(pts, knn)
    cov6 = natives.cov_knn(pts, rows)
    for i in range(0, len(pts), 11):
        nb = pts[rows[i][rows[i] >= 0]]
        want = np.cov((nb - nb.mean(axis=0)).T, bias=True)
        got = natives.cov6_to_matrix(cov6[i])
        # single-pass cumulants on unshifted RD coordinates (x ~ 1.2e5, y ~ 4.9e5): absolute error ~ eps * y^2
        assert np.abs(got - want).max() < 5e-4, i
    ev = natives.fast_eigen(cov6[::11])
    assert ev.shape == (len(cov6[::11]), 3)


# ------------------------------------------- non-default constructor arguments ---
VAR_RG = (dict(threshold_angle=5), dict(threshold_angle=60, grow_region_knn=8, max_nn=12, grow_region_radius=0.5),
          dict(threshold_curve=0.2))
VAR_LCC = dict(exclude_labels=[9], grid_size=0.2, min_component_size=10, threshold=0.03)
VAR_LAYER = [{'bottom': -1.0, 'top': 3.0, 'grid_size': 0.2, 'min_comp_size': 20, 'threshold': 0.2}]


def test_parameter_variants_oracle_matches_reference_golden(golden):
    """variants_golden.npz: the REAL reference classes with non-default arguments (building surface, no AHN cap,
    road type list, LCC exclusion list, LayerLCC(reset_noise=True), RegionGrowing angles / neighbourhoods /
    curvature threshold) vs the restatement."""
    g = golden('variants_golden.npz')
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']), facade_fraction=0.5)
    points, ahn = tile['points'], tile['ahn_tile']
    assert np.array_equal(points, from_mm(g['points_mm']))
    blds = [type(p)(np.array(p.exterior.coords)) for p in tile['building_polygons']]      # the CSV format has no holes
    zero = np.zeros(len(points), dtype=np.uint16)
    assert np.array_equal(R.ahn_fuser_labels(points, zero.copy(), zero == 0, ahn, 10, 'building', 0.1), g['ahn_building'])
    ground = R.ahn_fuser_labels(points, zero.copy(), zero == 0, ahn, 9, 'ground', 0.3)
    assert np.array_equal(ground, g['ahn_ground'])
    nocap = R.building_fuser_labels(points, ground.copy(), ground == 0, blds, 10, None)
    assert np.array_equal(nocap, g['building_no_ahn'])
    assert np.array_equal(R.road_fuser_labels(points, ground.copy(), tile['road_polygons'][1:], 1), g['road_fietspad'])
    seeded = ground.copy()
    seeded[g['car_seed_ids']] = 40
    lm = R.lcc_get_label_mask(points, seeded, seeded == 0, 40, **VAR_LCC)
    seeded[lm] = 40
    assert np.array_equal(seeded, g['lcc_exclude'])
    noisy = nocap.copy()
    noisy[g['noise_ids']] = 99
    got = R.layer_lcc_labels(points, noisy, noisy == 0, ahn, 10, [dict(p) for p in VAR_LAYER], reset_noise=True)
    assert np.array_equal(got, g['layer_lcc_reset_noise'])
    for k, kw in enumerate(VAR_RG):
        got = R.region_growing_labels(points, nocap.copy(), 10, exclude_labels=(9,), **kw)
        assert np.array_equal(got, g[f'region_growing_{k}']), kw
        assert (got != nocap).sum() > 500


def test_config1_oracle_matches_reference_golden(golden):
    """BASELINE config 1 at its full size (4 M points over the real demo AHN / BGT data): the oracle's
    restatement against the labels the reference's own Pipeline + RegionGrowing classes produced."""
    from conftest import CONFIG1_LAYERS, config1_cloud
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
    g = golden('config1_golden.npz')
    points, _ = config1_cloud(g)
    demo = os.path.join(GOLDEN, 'demo')
    ahn_tile = ahn_utils.load_ahn_tile(os.path.join(demo, 'ahn_2386_9702.npz'))
    bld = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_buildings_demo.csv')).filter_tile(
        '2386_9702', bgt_types=['pand'], merge=True)
    road = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_roads_demo.csv')).filter_tile(
        '2386_9702', bgt_types=['rijbaan_lokale_weg', 'parkeervlak', 'fietspad'], merge=True)
    o = np.zeros(len(points), dtype=np.uint16)
    o = R.ahn_fuser_labels(points, o, o == 0, ahn_tile, 9, 'ground', 0.2)
    o = R.road_fuser_labels(points, o, road, 1)
    o = R.building_fuser_labels(points, o, o == 0, bld, 10, ahn_tile, 0.2)
    o = R.layer_lcc_labels(points, o, o == 0, ahn_tile, 10, [dict(p) for p in CONFIG1_LAYERS])
    assert np.array_equal(o, g['labels_pipeline'])
    grown = R.region_growing_labels(points, o.copy(), 10, exclude_labels=(9,))
    assert np.array_equal(grown, g['labels_region_growing'])
    assert np.count_nonzero(grown == 10) > np.count_nonzero(o == 10) + 100_000


def test_alpha_shape_edges_match_reference_golden(golden):
    """Outer edges of the alpha complex: the product's vectorised alpha_shape_edges and the oracle's literal loop
    against the REAL get_alpha_shape_edges (alpha_golden.npz); the edges stitch into closed trails."""
    from urban_pointcloud_processing_b200.utils import alpha_shape_utils as A
    g = golden('alpha_golden.npz')
    for k in range(int(g['n_cases'])):
        pts = from_mm(g[f'points_mm_{k}'])
        for alpha in (2.0, 0.7):
            want = g[f'edges_{k}_{alpha}']
            got = A.alpha_shape_edges(pts, alpha)
            assert np.array_equal(got, want.astype(np.int64)), (k, alpha)
            ora = R.alpha_shape_outer_edges(pts, alpha)
            assert sorted(tuple(sorted(e)) for e in ora) == [tuple(e) for e in want.tolist()]
            trails = A.closed_trails(got)
            assert sum(len(t) - 1 for t in trails) == len(got) and all(t[0] == t[-1] for t in trails)
            # even-odd over the trails == union of the kept triangles: probe with the oracle's edge-soup test
            if len(got):
                rng = np.random.default_rng(k)
                probes = pts[rng.integers(0, len(pts), 300)] + rng.normal(0, 0.05, (300, 2))
                soup = R.edges_clip_buffered(probes, pts, [tuple(e) for e in got.tolist()], 0.0)
                loops = np.zeros(len(probes), dtype=bool)
                for t in trails:
                    loops ^= natives.is_inside(probes[:, 0], probes[:, 1], pts[t])
                assert np.array_equal(soup, loops)
