"""Generate tests/golden/* from the REAL reference code (run in the build container).

    python -m oracle.make_golden

Imports /root/reference/src/upcp unchanged through oracle.ref_harness (stand-ins only for
the third-party natives that are not installed), runs the reference's own functions /
classes on seeded inputs and stores inputs + outputs as small .npz fixtures.  While doing
so it asserts that the restatement in oracle/reference_py.py + oracle/natives.py gives
identical results -- this is what pins the oracle.  The fixtures travel to the GPU box;
/root/reference does not.
"""
import ast
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

from oracle import natives, ref_harness, reference_py as R          # noqa: E402
from urban_pointcloud_processing_b200 import synthetic              # noqa: E402

REF_DATA = '/root/reference/datasets'
TILECODE = '2386_9702'


def mm(a):
    """Coordinates on the LAS grid as integer millimetres (fixtures store these)."""
    return np.round(np.asarray(a) * 1000.0).astype(np.int64)


def from_mm(a):
    return a.astype(np.float64) * 0.001


def pack(mask):
    return np.packbits(np.asarray(mask, dtype=bool))


def copy_demo_data():
    dst = os.path.join(GOLDEN, 'demo')
    os.makedirs(dst, exist_ok=True)
    for rel in ('ahn/ahn_2386_9702.npz', 'ahn/ahn_2386_9702.laz', 'bgt/bgt_buildings_demo.csv', 'bgt/bgt_roads_demo.csv'):
        shutil.copyfile(os.path.join(REF_DATA, rel), os.path.join(dst, os.path.basename(rel)))
    with open(os.path.join(dst, 'README.md'), 'w') as f:
        f.write('Demo DATA fixtures copied verbatim from the reference checkout '
                '(`datasets/ahn/ahn_2386_9702.npz`, `datasets/ahn/ahn_2386_9702.laz`, `datasets/bgt/bgt_buildings_demo.csv`, '
                '`datasets/bgt/bgt_roads_demo.csv`; Amsterdam-AI-Team/Urban_PointCloud_Processing, '
                'GPL-3.0; sources: AHN / BGT open data).  Data only, no source code.\n')


def probe_points(rng, rings, n_random):
    """Random mm-grid points over the rings' bbox + every vertex, edge midpoint and points
    sharing a vertex' y (rays through vertices) -- the boundary cases of the crossing test."""
    allv = np.vstack(rings)
    lo = allv.min(0) - 1.0
    hi = allv.max(0) + 1.0
    pts = [from_mm(mm(lo + rng.random((n_random, 2)) * (hi - lo)))]
    for r in rings:
        pts.append(r.copy())
        pts.append(0.5 * (r[:-1] + r[1:]))
        k = min(len(r), 40)
        sel = rng.choice(len(r), size=k, replace=False)
        xs = from_mm(mm(lo[0] + rng.random(k) * (hi[0] - lo[0])))
        pts.append(np.column_stack((xs, r[sel, 1])))
        pts.append(np.column_stack((r[sel, 0], from_mm(mm(lo[1] + rng.random(k) * (hi[1] - lo[1]))))))
    return np.ascontiguousarray(np.vstack(pts))


def golden_pip(upcp):
    from upcp.utils import clip_utils
    rng = np.random.default_rng(1234)
    cases = {}
    # real BGT demo rings
    for name, csv in (('bgt_buildings', 'bgt/bgt_buildings_demo.csv'), ('bgt_roads', 'bgt/bgt_roads_demo.csv')):
        df = pd.read_csv(os.path.join(REF_DATA, csv))
        rings = [np.array(ast.literal_eval(p), dtype=np.float64) for p in df.polygon.values]
        cases[name] = [(r, []) for r in rings]
    # synthetic footprints with courtyards + road strips
    tile = synthetic.make_tile(1000, seed=7)
    cases['synthetic_buildings'] = [(np.array(p.exterior.coords), [np.array(i.coords) for i in p.interiors])
                                    for p in tile['building_polygons']]
    cases['synthetic_roads'] = [(np.array(p.exterior.coords), []) for p in tile['road_polygons']]
    # hand-made boundary cases: square with hole, horizontal edges, collinear vertices, overlap
    sq = np.array([(0, 0), (2, 0), (4, 0), (4, 4), (2, 4), (0, 4), (0, 2), (0, 0)], dtype=float)
    hole = np.array([(1, 1), (1, 3), (3, 3), (3, 1), (1, 1)], dtype=float)
    comb = np.array([(5, 0), (9, 0), (9, 3), (8, 3), (8, 1), (7, 1), (7, 3), (6, 3), (6, 1), (5, 1), (5, 0)], dtype=float)
    over = np.array([(3, 3), (6, 3), (6, 6), (3, 6), (3, 3)], dtype=float)
    degenerate = np.array([(0, 0), (1, 1)], dtype=float)
    cases['handmade'] = [(sq, [hole]), (comb, []), (over, []), (degenerate, [])]
    # a slice of the config-4 lattice
    lat = synthetic.make_polygon_lattice(n_side=9, seed=11)
    cases['lattice'] = [(np.array(p.exterior.coords), [np.array(i.coords) for i in p.interiors]) for p in lat]

    out = {}
    for name, polys in cases.items():
        rings = [np.asarray(e) for e, _ in polys if len(e) >= 2]
        n_random = 6000 if name != 'handmade' else 3000
        pts = probe_points(rng, rings, n_random)
        if name == 'handmade':      # integer-lattice probes hit every boundary configuration
            gx, gy = np.meshgrid(np.arange(-1, 11, 0.5), np.arange(-1, 8, 0.5))
            pts = np.vstack([pts, np.column_stack((gx.ravel(), gy.ravel()))])
        expected = np.zeros(len(pts), dtype=bool)
        per_poly = []
        for ext, ints in polys:
            poly = ref_harness.Polygon(ext, ints)
            m = clip_utils.poly_clip(pts, poly)                # REAL reference (numba)
            assert np.array_equal(m, R.poly_clip(pts, poly)), f'oracle PIP mismatch in {name}'
            per_poly.append(m)
            expected |= m
        ring_xy, ring_off, ring_poly, ring_hole = [], [0], [], []
        for p, (ext, ints) in enumerate(polys):
            # store the ring exactly as the stand-in polygon closes it
            poly = ref_harness.Polygon(ext, ints)
            for is_hole, ring in [(0, np.array(poly.exterior.coords).reshape(-1, 2))] + \
                                 [(1, np.array(i.coords).reshape(-1, 2)) for i in poly.interiors]:
                ring_xy.append(ring); ring_off.append(ring_off[-1] + len(ring))
                ring_poly.append(p); ring_hole.append(is_hole)
        out[f'{name}__points'] = pts
        out[f'{name}__ring_xy'] = np.vstack(ring_xy)
        out[f'{name}__ring_off'] = np.array(ring_off, dtype=np.int64)
        out[f'{name}__ring_poly'] = np.array(ring_poly, dtype=np.int32)
        out[f'{name}__ring_hole'] = np.array(ring_hole, dtype=np.uint8)
        out[f'{name}__expected'] = pack(expected)
        out[f'{name}__n'] = np.array(len(pts))
        print(f'  pip/{name}: {len(polys)} polygons, {len(pts)} probes, {expected.sum()} inside')
    np.savez_compressed(os.path.join(GOLDEN, 'pip_golden.npz'), **out)


def golden_raster(upcp):
    from upcp.utils.interpolation import FastGridInterpolator
    from upcp.utils.ahn_utils import load_ahn_tile
    rng = np.random.default_rng(4321)
    ahn = load_ahn_tile(os.path.join(REF_DATA, 'ahn', f'ahn_{TILECODE}.npz'))
    gx, gy = ahn['x'], ahn['y']
    bin_x = gx - (gx[1] - gx[0]) / 2
    bin_y = gy + (gy[0] - gy[1]) / 2
    pts = [np.column_stack((gx[0] - 1 + rng.random(6000) * 52, gy[-1] - 1 + rng.random(6000) * 52))]
    # exact bin edges and their fp64 neighbours, on both axes
    for e in (bin_x, np.nextafter(bin_x, -np.inf), np.nextafter(bin_x, np.inf)):
        pts.append(np.column_stack((e, np.full(len(e), gy[250]))))
    for e in (bin_y, np.nextafter(bin_y, -np.inf), np.nextafter(bin_y, np.inf)):
        pts.append(np.column_stack((np.full(len(e), gx[250]), e)))
    pts.append(np.array([[gx[0] - 10, gy[0] + 10], [gx[-1] + 10, gy[-1] - 10], [gx[0] - 10, gy[100]],
                         [gx[100], gy[0] + 10]]))
    pts.append(from_mm(mm(np.column_stack((gx[0] + rng.random(3000) * 50, gy[-1] + rng.random(3000) * 50)))))
    pts = np.ascontiguousarray(np.vstack(pts))
    out = {'points': pts}
    for surface in ('ground_surface', 'building_surface'):
        ref = FastGridInterpolator(gx, gy, ahn[surface])(pts)      # REAL reference
        mine = R.fast_grid_lookup(gx, gy, ahn[surface], pts)
        assert np.array_equal(ref, mine, equal_nan=True), 'oracle raster mismatch'
        out[surface] = ref
    np.savez_compressed(os.path.join(GOLDEN, 'raster_golden.npz'), **out)
    print(f'  raster: {len(pts)} probes')


def golden_scalars(upcp):
    from upcp.utils import math_utils
    rng = np.random.default_rng(99)
    ext, gs, lvl = [], [], []
    for e in list(rng.uniform(0.0005, 200, 60)) + [0.0009, 0.001, 50.0, 50.001, 49.999, 12.8, 25.6, 51.2]:
        for g in (0.05, 0.1, 0.2, 0.4, 1.0):
            p = np.zeros((2, 3)); p[1, 0] = e; p[1, 1] = e * 0.3
            ext.append(e); gs.append(g); lvl.append(int(math_utils.get_octree_level(p, g)))
            assert lvl[-1] == R.get_octree_level(p, g)
    u = rng.normal(size=(200, 3)); v = rng.normal(size=(200, 3))
    ang = np.array([math_utils.vector_angle(a, b) for a, b in zip(u, v)])
    np.savez_compressed(os.path.join(GOLDEN, 'scalars_golden.npz'), extent=np.array(ext), grid_size=np.array(gs),
                        level=np.array(lvl), u=u, v=v, angle=ang)
    print(f'  scalars: {len(lvl)} octree levels, {len(ang)} angles')


def write_readers(tmp, tile):
    """npz + csv files in the reference's own formats for the synthetic tile."""
    ahn = tile['ahn_tile']
    np.savez(os.path.join(tmp, f"ahn_{tile['tilecode']}.npz"), x=ahn['x'], y=ahn['y'],
             ground=ahn['ground_surface'].astype(np.float16), building=ahn['building_surface'].astype(np.float16))
    def csv(polys, types, path):
        rows = []
        for p, t in zip(polys, types):
            ring = [[float(c[0]), float(c[1])] for c in p.exterior.coords]
            a = np.array(ring)
            rows.append((t, str(ring), a[:, 0].min(), a[:, 1].max(), a[:, 0].max(), a[:, 1].min()))
        pd.DataFrame(rows, columns=['bgt_type', 'polygon', 'x_min', 'y_max', 'x_max', 'y_min']).to_csv(path, index=False)
    csv(tile['building_polygons'], ['pand'] * len(tile['building_polygons']), os.path.join(tmp, 'buildings.csv'))
    csv(tile['road_polygons'], ['rijbaan_lokale_weg', 'fietspad'], os.path.join(tmp, 'roads.csv'))


def golden_pipeline(upcp):
    """Reference Pipeline + processors (real Python, restated natives) on a synthetic tile."""
    from upcp.pipeline import Pipeline
    from upcp.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser, NoiseFilter
    from upcp.region_growing import LabelConnectedComp, LayerLCC
    from upcp.region_growing.region_growing import RegionGrowing
    from upcp.utils.ahn_utils import NPZReader
    from upcp.utils.bgt_utils import BGTPolyReader
    from upcp.labels import Labels

    n = 60000
    tile = synthetic.make_tile(n, seed=20240001)
    # NOTE: the csv round trip drops interior rings (the reference's CSV format has none)
    points = tile['points']
    tc = tile['tilecode']
    tmp = tempfile.mkdtemp(prefix='upcp_golden_')
    write_readers(tmp, tile)
    ahn_reader = NPZReader(tmp)
    bld_reader = BGTPolyReader(bgt_file=os.path.join(tmp, 'buildings.csv'))
    road_reader = BGTPolyReader(bgt_file=os.path.join(tmp, 'roads.csv'))
    ahn_tile = ahn_reader.filter_tile(tc)
    bld_polys = bld_reader.filter_tile(tc, bgt_types=['pand'], merge=True)
    road_polys = road_reader.filter_tile(tc, bgt_types=BGTRoadFuser.ALL_TYPES, merge=True)

    building_top = {'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5}
    building_bottom = {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}
    stages = [
        ('ahn_ground', AHNFuser(Labels.GROUND, ahn_reader, target='ground', epsilon=0.2, refine_ground=False)),
        ('bgt_road', BGTRoadFuser(Labels.ROAD, bgt_reader=road_reader)),
        ('bgt_building', BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld_reader, offset=0, ahn_reader=ahn_reader)),
        ('layer_lcc', LayerLCC(Labels.BUILDING, ahn_reader, params=[dict(building_top), dict(building_bottom)])),
        ('lcc_car', LabelConnectedComp(Labels.CAR, grid_size=0.1, min_component_size=30, threshold=0.02)),
        ('noise', NoiseFilter(Labels.NOISE, ahn_reader, epsilon=0.2, min_component_size=10)),
        ('region_growing', RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD))),
    ]
    labels = np.zeros(n, dtype=np.uint16)
    # seed a few car points so that the LCC stage has something to grow from
    car_ids = np.where(tile['kind'] == 2)[0]
    out = {'points_mm': mm(points).astype(np.int32), 'seed': np.array(20240001), 'n': np.array(n)}
    oracle_labels = labels.copy()
    ahn_reader.cache_interpolator(tc, points, surface='ground_surface')
    for name, proc in stages:
        if name == 'lcc_car':
            labels[car_ids[::50]] = Labels.CAR
            oracle_labels[car_ids[::50]] = Labels.CAR
            out['car_seed_ids'] = car_ids[::50].astype(np.int32)
        mask = labels == 0
        labels = proc.get_labels(points, labels, mask, tc)          # REAL reference control flow
        # the same stage through the restatement
        omask = oracle_labels == 0
        if name == 'ahn_ground':
            oracle_labels = R.ahn_fuser_labels(points, oracle_labels, omask, ahn_tile, Labels.GROUND, 'ground', 0.2)
        elif name == 'bgt_road':
            oracle_labels = R.road_fuser_labels(points, oracle_labels, road_polys, Labels.ROAD)
        elif name == 'bgt_building':
            oracle_labels = R.building_fuser_labels(points, oracle_labels, omask, bld_polys, Labels.BUILDING,
                                                    ahn_tile, 0.2)
        elif name == 'layer_lcc':
            oracle_labels = R.layer_lcc_labels(points, oracle_labels, omask, ahn_tile, Labels.BUILDING,
                                               [building_top, building_bottom])
        elif name == 'lcc_car':
            lm = R.lcc_get_label_mask(points, oracle_labels, omask, Labels.CAR, grid_size=0.1,
                                      min_component_size=30, threshold=0.02)
            oracle_labels[lm] = Labels.CAR
        elif name == 'noise':
            oracle_labels = R.noise_filter_labels(points, oracle_labels, omask, ahn_tile, Labels.NOISE, 0.2, 0.1, 10)
        elif name == 'region_growing':
            oracle_labels = R.region_growing_labels(points, oracle_labels, Labels.BUILDING,
                                                    exclude_labels=(Labels.GROUND, Labels.ROAD))
        assert np.array_equal(labels, oracle_labels), f'oracle mismatch after stage {name}'
        out[f'labels_after_{name}'] = labels.astype(np.uint8).copy()
        print(f'  pipeline/{name}: {np.count_nonzero(labels)} labelled, hist '
              f'{dict(zip(*np.unique(labels, return_counts=True)))}')
    # get_components through the real class
    comps = LabelConnectedComp(grid_size=0.2, min_component_size=20).get_components(points[tile['kind'] >= 2])
    assert np.array_equal(R.canonical_components(comps),
                          R.canonical_components(R.lcc_get_components(points[tile['kind'] >= 2], 0.2, 20)))
    out['components_kind_ge2'] = R.canonical_components(comps).astype(np.int32)
    np.savez_compressed(os.path.join(GOLDEN, 'pipeline_golden.npz'), **out)
    shutil.rmtree(tmp, ignore_errors=True)


def golden_demo(upcp):
    """Real demo AHN/BGT data + the reference fusers on a synthetic cloud over tile 2386_9702."""
    from upcp.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from upcp.utils.ahn_utils import NPZReader
    from upcp.utils.bgt_utils import BGTPolyReader
    from upcp.labels import Labels
    from upcp.pipeline import Pipeline
    rng = np.random.default_rng(23869702)
    n = 80000
    ahn_reader = NPZReader(os.path.join(REF_DATA, 'ahn'))
    ahn = ahn_reader.filter_tile(TILECODE)
    x = 119300 + rng.random(n) * 50
    y = 485100 + rng.random(n) * 50
    pts2 = from_mm(mm(np.column_stack((x, y))))
    gz = R.fast_grid_lookup(ahn['x'], ahn['y'], ahn['ground_surface'], pts2)
    bz = R.fast_grid_lookup(ahn['x'], ahn['y'], ahn['building_surface'], pts2)
    z = np.where(rng.random(n) < 0.5, np.nan_to_num(gz, nan=0.5) + rng.normal(0, 0.15, n),
                 np.nan_to_num(bz, nan=6.0) * rng.random(n) + rng.normal(0, 0.2, n))
    points = np.asfortranarray(np.column_stack((pts2, from_mm(mm(z)))))
    bld = BGTPolyReader(bgt_file=os.path.join(REF_DATA, 'bgt', 'bgt_buildings_demo.csv'))
    road = BGTPolyReader(bgt_file=os.path.join(REF_DATA, 'bgt', 'bgt_roads_demo.csv'))
    procs = (AHNFuser(Labels.GROUND, ahn_reader, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn_reader))
    labels = np.zeros(n, dtype=np.uint16)
    pipeline = Pipeline(processors=procs, ahn_reader=ahn_reader, caching=True)
    labels = pipeline.process_cloud(TILECODE, points, labels)          # REAL reference pipeline
    np.savez_compressed(os.path.join(GOLDEN, 'demo_golden.npz'), points_mm=mm(points).astype(np.int32),
                        labels=labels.astype(np.uint8))
    print(f'  demo: hist {dict(zip(*np.unique(labels, return_counts=True)))}')


OBJ_N, OBJ_SEED, OBJ_CARS = 600000, 20240007, 12
CAR_PARAMS = {'min_height': 1.2, 'max_height': 2.4, 'min_width': 1.4, 'max_width': 2.4, 'min_length': 3.0,
              'max_length': 6.0}
POLE_PARAMS = {'min_height': 1.5, 'max_height': 9.0, 'min_width': 0.1, 'max_width': 1.2, 'min_length': 0.1,
               'max_length': 1.2}


def golden_objects(upcp):
    """The REAL CarFuser / BGTStreetFurnitureFuser classes (clients of get_components) on a synthetic tile;
    stand-ins: clustering (oracle natives) and the rectangle / road overlap (half-plane clipping)."""
    from upcp.fusion import AHNFuser, CarFuser, BGTStreetFurnitureFuser
    from upcp.utils.ahn_utils import NPZReader
    from upcp.utils.bgt_utils import BGTPolyReader, BGTPointReader
    from upcp.labels import Labels
    from urban_pointcloud_processing_b200 import synthetic
    tile = synthetic.make_tile(OBJ_N, seed=OBJ_SEED, n_cars=OBJ_CARS)
    points, tc = tile['points'], tile['tilecode']
    tmp = tempfile.mkdtemp()
    write_readers(tmp, tile)
    pd.DataFrame([('lichtmast', x, y) for x, y in tile['pole_xy']][::2] + [('boom', x, y) for x, y in tile['pole_xy']][1::2],
                 columns=['bgt_type', 'x', 'y']).to_csv(os.path.join(tmp, 'points.csv'), index=False)
    ahn_reader = NPZReader(tmp)
    road = BGTPolyReader(bgt_file=os.path.join(tmp, 'roads.csv'))
    poles = BGTPointReader(bgt_file=os.path.join(tmp, 'points.csv'))
    ahn_tile = tile['ahn_tile']
    out = {'n': np.array(OBJ_N), 'seed': np.array(OBJ_SEED), 'points_mm': mm(points).astype(np.int32)}
    labels = np.zeros(OBJ_N, dtype=np.uint16)
    labels = AHNFuser(Labels.GROUND, ahn_reader, target='ground', epsilon=0.2, refine_ground=False).get_labels(
        points, labels, labels == 0, tc)
    want = labels.copy()
    out['labels_after_ground'] = labels.astype(np.uint8).copy()
    car = CarFuser(Labels.CAR, ahn_reader, road, grid_size=0.1, min_component_size=300, params=dict(CAR_PARAMS),
                   exclude_types=['fietspad'])
    labels = car.get_labels(points, labels, labels == 0, tc)
    roads_kept = road.filter_tile(tc, exclude_types=['fietspad'], merge=False)
    want = R.car_fuser_labels(points, want, want == 0, ahn_tile, roads_kept, Labels.CAR, 0.1, 300, 20, CAR_PARAMS)
    assert np.array_equal(labels, want), 'oracle mismatch: car fuser'
    out['labels_after_car'] = labels.astype(np.uint8).copy()
    n_car = int((labels == Labels.CAR).sum())
    sf = BGTStreetFurnitureFuser(Labels.STREET_LIGHT, 'lichtmast', poles, ahn_reader, grid_size=0.1,
                                 min_component_size=100, params=dict(POLE_PARAMS))
    labels = sf.get_labels(points, labels, labels == 0, tc)
    pts_kept = poles.filter_tile(tc, bgt_types=['lichtmast'])
    want = R.street_furniture_labels(points, want, want == 0, ahn_tile, pts_kept, Labels.STREET_LIGHT, 0.1, 100, 1.,
                                     POLE_PARAMS)
    assert np.array_equal(labels, want), 'oracle mismatch: street furniture fuser'
    out['labels_after_street_light'] = labels.astype(np.uint8).copy()
    n_sf = int((labels == Labels.STREET_LIGHT).sum())
    # minimum bounding rectangles of a few clusters, straight from the reference function
    from upcp.utils.math_utils import minimum_bounding_rectangle
    rng = np.random.default_rng(5)
    rects = []
    for k in range(6):
        c = rng.random((200 + 50 * k, 2)) * np.array([4.5, 1.8])
        a = rng.uniform(0, np.pi)
        c = c @ np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) + np.array([119320.0, 485120.0])
        rect, _, w, l, ctr = minimum_bounding_rectangle(c)
        r2, w2, l2, c2 = R.minimum_bounding_rectangle(c)
        assert np.array_equal(rect, r2) and w == w2 and l == l2 and np.array_equal(ctr, c2)
        out[f'mbr_in_{k}'] = c
        out[f'mbr_out_{k}'] = np.concatenate([rect.ravel(), [w, l], ctr])
    np.savez_compressed(os.path.join(GOLDEN, 'objects_golden.npz'), **out)
    shutil.rmtree(tmp, ignore_errors=True)
    print(f'  objects: {n_car} car points, {n_sf} street-light points')
    assert n_car > 1000 and n_sf > 300


VAR_N, VAR_SEED = 40000, 20240011
VAR_RG = (dict(threshold_angle=5), dict(threshold_angle=60, grow_region_knn=8, max_nn=12, grow_region_radius=0.5),
          dict(threshold_curve=0.2))
VAR_LCC = dict(exclude_labels=[9], grid_size=0.2, min_component_size=10, threshold=0.03)
VAR_LAYER = [{'bottom': -1.0, 'top': 3.0, 'grid_size': 0.2, 'min_comp_size': 20, 'threshold': 0.2}]


def golden_variants(upcp):
    """Non-default constructor arguments through the REAL reference classes: AHNFuser(target='building'),
    BGTBuildingFuser without the AHN cap, BGTRoadFuser(bgt_types=...), LabelConnectedComp(exclude_labels=...),
    LayerLCC(reset_noise=True), RegionGrowing with other angles / neighbourhoods / a curvature threshold
    inside the range of the ratios (the LAPACK eigenvalue-order path)."""
    from upcp.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from upcp.region_growing import LabelConnectedComp, LayerLCC
    from upcp.region_growing.region_growing import RegionGrowing
    from upcp.utils.ahn_utils import NPZReader
    from upcp.utils.bgt_utils import BGTPolyReader
    from upcp.labels import Labels
    tile = synthetic.make_tile(VAR_N, seed=VAR_SEED, facade_fraction=0.5)
    points, tc = tile['points'], tile['tilecode']
    tmp = tempfile.mkdtemp(prefix='upcp_golden_')
    write_readers(tmp, tile)
    ahn_reader = NPZReader(tmp)
    bld_reader = BGTPolyReader(bgt_file=os.path.join(tmp, 'buildings.csv'))
    road_reader = BGTPolyReader(bgt_file=os.path.join(tmp, 'roads.csv'))
    ahn_tile = ahn_reader.filter_tile(tc)
    out = {'points_mm': mm(points).astype(np.int32), 'seed': np.array(VAR_SEED), 'n': np.array(VAR_N)}

    def both(name, ref_call, oracle_call, labels):
        got = ref_call(labels.copy())
        want = oracle_call(labels.copy())
        assert np.array_equal(got, want), f'oracle mismatch: {name}'
        out[name] = got.astype(np.uint8)
        print(f'  variants/{name}: {np.count_nonzero(got != labels)} labels changed')
        return got

    zero = np.zeros(VAR_N, dtype=np.uint16)
    bsurf = both('ahn_building', lambda l: AHNFuser(Labels.BUILDING, ahn_reader, target='building', epsilon=0.1,
                                                    refine_ground=False).get_labels(points, l, l == 0, tc),
                 lambda l: R.ahn_fuser_labels(points, l, l == 0, ahn_tile, Labels.BUILDING, 'building', 0.1), zero)
    ground = both('ahn_ground', lambda l: AHNFuser(Labels.GROUND, ahn_reader, target='ground', epsilon=0.3,
                                                   refine_ground=False).get_labels(points, l, l == 0, tc),
                  lambda l: R.ahn_fuser_labels(points, l, l == 0, ahn_tile, Labels.GROUND, 'ground', 0.3), zero)
    bld_polys = bld_reader.filter_tile(tc, bgt_types=['pand'], merge=True)
    nocap = both('building_no_ahn', lambda l: BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld_reader, offset=0)
                 .get_labels(points, l, l == 0, tc),
                 lambda l: R.building_fuser_labels(points, l, l == 0, bld_polys, Labels.BUILDING, None), ground)
    bike = road_reader.filter_tile(tc, bgt_types=['fietspad'], merge=True)
    both('road_fietspad', lambda l: BGTRoadFuser(Labels.ROAD, bgt_reader=road_reader, bgt_types=['fietspad'])
         .get_labels(points, l, l == 0, tc),
         lambda l: R.road_fuser_labels(points, l, bike, Labels.ROAD), ground)
    # LabelConnectedComp with an exclusion list: 2 % of the car points as seeds
    seeded = ground.copy()
    cars = np.where(tile['kind'] == 2)[0]
    seeded[cars[::50]] = Labels.CAR
    out['car_seed_ids'] = cars[::50].astype(np.int32)

    def ref_lcc(l):
        return LabelConnectedComp(Labels.CAR, **VAR_LCC).get_labels(points, l, l == 0, tc)

    def oracle_lcc(l):
        lm = R.lcc_get_label_mask(points, l, l == 0, Labels.CAR, **VAR_LCC)
        l[lm] = Labels.CAR
        return l
    both('lcc_exclude', ref_lcc, oracle_lcc, seeded)
    # LayerLCC with reset_noise: some building points and some unlabelled ones arrive as NOISE
    noisy = nocap.copy()
    rng = np.random.default_rng(4)
    flip = rng.random(VAR_N) < 0.1
    noisy[flip & (noisy != Labels.GROUND)] = Labels.NOISE
    out['noise_ids'] = np.where(noisy == Labels.NOISE)[0].astype(np.int32)
    both('layer_lcc_reset_noise',
         lambda l: LayerLCC(Labels.BUILDING, ahn_reader, reset_noise=True, params=[dict(p) for p in VAR_LAYER])
         .get_labels(points, l, l == 0, tc),
         lambda l: R.layer_lcc_labels(points, l, l == 0, ahn_tile, Labels.BUILDING, [dict(p) for p in VAR_LAYER],
                                      reset_noise=True), noisy)
    for k, kw in enumerate(VAR_RG):
        both(f'region_growing_{k}',
             lambda l: RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,), **kw).get_labels(points, l, None, tc),
             lambda l: R.region_growing_labels(points, l, Labels.BUILDING, exclude_labels=(Labels.GROUND,), **kw), nocap)
    np.savez_compressed(os.path.join(GOLDEN, 'variants_golden.npz'), **out)
    shutil.rmtree(tmp, ignore_errors=True)
    assert np.count_nonzero(bsurf) > 1000


CONFIG1_N, CONFIG1_SEED = 4_000_000, 23869702
CONFIG1_LAYERS = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                  {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]


def config1_cloud(data_dir):
    """The config-1 cloud (regenerated from its seed wherever it is needed: the fixture stores labels only)."""
    from urban_pointcloud_processing_b200.utils.ahn_utils import load_ahn_tile
    ahn = load_ahn_tile(os.path.join(data_dir, 'ahn_2386_9702.npz'))
    df = pd.read_csv(os.path.join(data_dir, 'bgt_buildings_demo.csv'))
    rings = [np.array(ast.literal_eval(p), dtype=np.float64) for p in df.polygon.values]
    return synthetic.make_demo_cloud(CONFIG1_N, ahn, rings, seed=CONFIG1_SEED)


def golden_config1(upcp):
    """BASELINE.json config 1 as SURVEY 8(d) specifies it: 4 M synthetic points over the REAL demo AHN
    raster (NaN holes) and BGT footprints, the notebook-0 processor list restricted to the hot path
    (AHN ground -> BGT road -> BGT building -> LayerLCC) through the reference's own Pipeline, then the
    reference's RegionGrowing on the result (Region growing.ipynb).  shapely is not installed here, so
    offset = 0 and refine_ground = False (flagged in DESIGN.md)."""
    import time
    from upcp.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from upcp.region_growing import LayerLCC
    from upcp.region_growing.region_growing import RegionGrowing
    from upcp.utils.ahn_utils import NPZReader
    from upcp.utils.bgt_utils import BGTPolyReader
    from upcp.labels import Labels
    from upcp.pipeline import Pipeline
    points, kind = config1_cloud(os.path.join(GOLDEN, 'demo'))
    ahn_reader = NPZReader(os.path.join(REF_DATA, 'ahn'))
    bld = BGTPolyReader(bgt_file=os.path.join(REF_DATA, 'bgt', 'bgt_buildings_demo.csv'))
    road = BGTPolyReader(bgt_file=os.path.join(REF_DATA, 'bgt', 'bgt_roads_demo.csv'))
    procs = (AHNFuser(Labels.GROUND, ahn_reader, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn_reader),
             LayerLCC(Labels.BUILDING, ahn_reader, params=[dict(p) for p in CONFIG1_LAYERS]))
    labels = np.zeros(len(points), dtype=np.uint16)
    t0 = time.time()
    labels = Pipeline(processors=procs, ahn_reader=ahn_reader, caching=True).process_cloud(TILECODE, points, labels)
    t1 = time.time()
    print(f'  config1 pipeline: {t1 - t0:.1f} s, hist {dict(zip(*np.unique(labels, return_counts=True)))}')
    rg = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,))
    grown = rg.get_labels(points, labels.copy(), labels == 0, TILECODE)          # REAL reference class
    t2 = time.time()
    print(f'  config1 region growing: {t2 - t1:.1f} s, hist {dict(zip(*np.unique(grown, return_counts=True)))}')
    # the restatement gives the same labels (this is what pins the oracle's control flow at this size)
    ahn_tile = ahn_reader.filter_tile(TILECODE)
    o = np.zeros(len(points), dtype=np.uint16)
    o = R.ahn_fuser_labels(points, o, o == 0, ahn_tile, Labels.GROUND, 'ground', 0.2)
    o = R.road_fuser_labels(points, o, road.filter_tile(TILECODE, bgt_types=BGTRoadFuser.ALL_TYPES, merge=True), Labels.ROAD)
    o = R.building_fuser_labels(points, o, o == 0, bld.filter_tile(TILECODE, bgt_types=['pand'], merge=True),
                                Labels.BUILDING, ahn_tile, 0.2)
    o = R.layer_lcc_labels(points, o, o == 0, ahn_tile, Labels.BUILDING, [dict(p) for p in CONFIG1_LAYERS])
    assert np.array_equal(o, labels), 'oracle mismatch on config 1 (pipeline)'
    np.savez_compressed(os.path.join(GOLDEN, 'config1_golden.npz'), n=np.array(CONFIG1_N), seed=np.array(CONFIG1_SEED),
                        labels_pipeline=labels.astype(np.uint8), labels_region_growing=grown.astype(np.uint8),
                        checksum_points=np.array(int(mm(points).astype(np.int64).sum())))


def golden_alpha(upcp):
    """Outer edges of the alpha complex from the REAL get_alpha_shape_edges (scipy only) on seeded clusters --
    pins utils/alpha_shape_utils.alpha_shape_edges and the oracle's literal loop."""
    from upcp.utils import alpha_shape_utils as RA
    rng = np.random.default_rng(20240099)
    out = {}
    for k in range(8):
        n = int(rng.integers(50, 2500))
        pts = np.column_stack((rng.normal(0, 0.6 + 0.5 * k, n), rng.normal(0, 0.3 + 0.15 * k, n)))
        if k % 3 == 1:                                     # an L-shaped, hollow cluster
            pts = pts[(np.abs(pts[:, 0]) > 0.4) | (pts[:, 1] > 0.0)]
        if k == 7:
            pts = np.vstack([pts, pts[:40]])                 # duplicate xy (points above each other)
        pts = from_mm(mm(pts + np.array([119320.0, 485120.0])))
        for alpha in (2.0, 0.7):
            edges = RA.get_alpha_shape_edges(pts, alpha, only_outer=True)          # REAL reference
            assert edges == R.alpha_shape_outer_edges(pts, alpha), 'oracle alpha-shape mismatch'
            und = np.array(sorted(tuple(sorted(e)) for e in edges), dtype=np.int32).reshape(-1, 2)
            out[f'edges_{k}_{alpha}'] = und
        out[f'points_mm_{k}'] = mm(pts).astype(np.int64)
    out['n_cases'] = np.array(8)
    np.savez_compressed(os.path.join(GOLDEN, 'alpha_golden.npz'), **out)
    print(f"  alpha: {sum(len(v) for k_, v in out.items() if k_.startswith('edges'))} outer edges over 16 cases")


def ahn_cloud(seed=424242, n=60000):
    """A synthetic AHN-like aerial cloud over tile 2386_9702 (+ 1 m buffer): ~ 20 points / m2, classes ground (2),
    building (6), other (1), water (9), with gaps (no ground under buildings, an empty strip)."""
    rng = np.random.default_rng(seed)
    x = 119299.0 + rng.random(n) * 52.0
    y = 485099.0 + rng.random(n) * 52.0
    cls = np.full(n, 2, dtype=np.uint8)
    z = 0.4 + 0.3 * np.sin((x - 119300) / 7.0) * np.cos((y - 485100) / 9.0) + rng.normal(0, 0.03, n)
    for (x0, y0, x1, y1, h) in ((119305, 485104, 119318, 485117, 9.0), (119327, 485125, 119341, 485139, 14.5)):
        inside = (x > x0) & (x < x1) & (y > y0) & (y < y1)
        cls[inside] = 6
        z[inside] = h + 0.02 * (x[inside] - x0) + rng.normal(0, 0.05, inside.sum())
    other = rng.random(n) < 0.08
    cls[other & (cls == 2)] = 1
    z[other & (cls == 1)] += rng.random((other & (cls == 1)).sum()) * 6.0
    strip = (y > 485142) & (y < 485144.5)
    keep = ~strip
    cls[(x > 119344) & (cls == 2)] = 9
    # a few grid-exact points: positions that coincide with a raster cell centre (confusion distance)
    x[:50] = 119300.05 + 0.1 * rng.integers(0, 500, 50); y[:50] = 485149.95 - 0.1 * rng.integers(0, 500, 50)
    X = mm(x[keep]); Y = mm(y[keep]); Z = mm(z[keep])
    return X.astype(np.int32), Y.astype(np.int32), Z.astype(np.int32), cls[keep]


def golden_ahn_preprocessing(upcp):
    """The REAL SpatialInterpolator (scipy cKDTree + the reference's loop) through the reference's
    _get_ahn_surface arithmetic on a synthetic AHN cloud: ground (idw, 8 nearest within 1 m, power 2) and
    building (max within 0.5 m) rasters, rounded to 2 decimals / float16 as process_ahn_las_tile stores them."""
    from upcp.utils.interpolation import SpatialInterpolator
    X, Y, Z, cls = ahn_cloud()
    x, y, z = from_mm(X), from_mm(Y), from_mm(Z)
    (x_min, y_max), (x_max, y_min) = (119300, 485150), (119350, 485100)
    res = 0.1
    grid_y, grid_x = np.mgrid[y_max - res / 2:y_min:-res, x_min + res / 2:x_max:res]
    positions = np.vstack((grid_x.reshape(-1), grid_y.reshape(-1))).T
    out = {'X': X, 'Y': Y, 'Z': Z, 'classification': cls}
    for name, label, method, max_dist in (('ground', 2, 'idw', 1), ('building', 6, 'max', 0.5)):
        m = cls == label
        pts = np.vstack((x, y, z)).T[m]
        idw = SpatialInterpolator(pts[:, 0:2], pts[:, 2], method=method)                 # REAL reference class
        grid = idw(positions, n_neighbors=8, max_dist=max_dist, power=2., fill_value=np.nan)
        out[name + '_f64'] = grid.reshape(grid_x.shape)
        out[name] = np.around(grid.reshape(grid_x.shape), decimals=2).astype('float16')
    np.savez_compressed(os.path.join(GOLDEN, 'ahn_preprocessing_golden.npz'), **out)
    print(f"  ahn preprocessing: ground NaN {np.isnan(out['ground']).mean():.3f}, building NaN {np.isnan(out['building']).mean():.3f}")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    upcp = ref_harness.load_reference()
    print('reference loaded from', upcp.__file__)
    steps = {'demo_data': lambda u: copy_demo_data(), 'pip': golden_pip, 'raster': golden_raster,
             'scalars': golden_scalars, 'demo': golden_demo, 'pipeline': golden_pipeline, 'objects': golden_objects,
             'variants': golden_variants, 'config1': golden_config1, 'alpha': golden_alpha,
             'ahn_preprocessing': golden_ahn_preprocessing}
    wanted = sys.argv[1:] or list(steps)
    for name in wanted:
        print(name)
        steps[name](upcp)
    print('golden vectors written to', GOLDEN)


if __name__ == '__main__':
    main()
