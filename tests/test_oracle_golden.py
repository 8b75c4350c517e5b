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
