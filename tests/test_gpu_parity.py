"""GPU: parity of the CUDA path (through the C ABI / the drop-in processors) with the oracle
and with the golden vectors generated from the real reference code.

Bar: bit-exact for masks, labels, indices, component membership (after canonical
relabelling) and region membership; floating point outputs (raster values, squared
distances, covariances) bit-exact, normals within 1e-9 (device acos/cos vs glibc), with
points whose angle test lies within 1e-6 degrees of the threshold counted and excluded.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, from_mm, unpack
from oracle import natives, reference_py as R
from test_oracle_golden import PIP_CASES, polygons_from_golden
from urban_pointcloud_processing_b200 import _lib, synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser, NoiseFilter
from urban_pointcloud_processing_b200.labels import Labels
from urban_pointcloud_processing_b200.pipeline import Pipeline
from urban_pointcloud_processing_b200.region_growing import LabelConnectedComp, LayerLCC
from urban_pointcloud_processing_b200.region_growing import region_growing as rgmod
from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing
from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils, clip_utils
from urban_pointcloud_processing_b200.utils.bgt_utils import RingPolygon
from urban_pointcloud_processing_b200.utils.interpolation import FastGridInterpolator

pytestmark = pytest.mark.gpu


def readers_for(tile, drop_holes=False):
    blds = tile['building_polygons']
    if drop_holes:
        blds = [RingPolygon(np.array(p.exterior.coords)) for p in blds]
    tc = tile['tilecode']
    ahn = ahn_utils.ArrayAHNReader({tc: tile['ahn_tile']})
    bld = bgt_utils.ArrayBGTReader({tc: blds}, {tc: ['pand'] * len(blds)})
    road = bgt_utils.ArrayBGTReader({tc: tile['road_polygons']}, {tc: ['rijbaan_lokale_weg', 'fietspad']})
    return ahn, bld, road


# ------------------------------------------------------------------ tile / masks ---
def test_tile_upload_layouts_and_bbox():
    rng = np.random.default_rng(0)
    pts_c = np.ascontiguousarray(rng.random((10007, 3)) * 50 + np.array([119300, 485100, -2]))
    for pts in (pts_c, np.asfortranarray(pts_c), pts_c[::2], pts_c[:, :2].copy()):
        t = dev.DeviceTile(pts)
        assert np.array_equal(t.x.cpu().numpy(), pts[:, 0]) and np.array_equal(t.y.cpu().numpy(), pts[:, 1])
        if pts.shape[1] == 3:
            assert np.array_equal(t.z.cpu().numpy(), pts[:, 2])
        else:
            assert not t.z.cpu().numpy().any()
    t = dev.DeviceTile(pts_c)
    sel = rng.random(len(pts_c)) < 0.3
    ball, bsel, m = t.bbox(dev.upload_mask(sel, t.n, t.device))
    assert np.array_equal(ball, np.concatenate([pts_c.min(0), pts_c.max(0)]))
    assert np.array_equal(bsel, np.concatenate([pts_c[sel].min(0), pts_c[sel].max(0)])) and m == sel.sum()
    # tile cache: only READ-ONLY arrays keep their HBM copy (same live array -> same copy); a writable array is
    # uploaded on every call, so an in-place edit between two calls is always seen (ADVICE r1 / VERDICT r1 #5)
    assert dev.get_tile(pts_c) is not dev.get_tile(pts_c)
    frozen = pts_c.copy(); frozen.flags.writeable = False
    a = dev.get_tile(frozen)
    assert dev.get_tile(frozen) is a and dev.get_tile(frozen.copy()) is not a and dev.get_tile(a) is a
    staging = pts_c.copy()
    ground = ahn_utils.ArrayAHNReader({'x': {'x': np.array([0.0, 1e6]), 'y': np.array([1e6, 0.0]),
                                             'ground_surface': np.zeros((2, 2)), 'building_surface': np.zeros((2, 2))}})
    fuser = AHNFuser(Labels.GROUND, ground, target='ground', epsilon=0.5, refine_ground=False)
    staging[:, 2] = 0.0
    first = fuser.get_label_mask(staging, np.zeros(len(staging), dtype=np.uint16), None, 'x')
    staging[:, 2] = 10.0                       # refilled in place: same array object, same sampled rows elsewhere
    second = fuser.get_label_mask(staging, np.zeros(len(staging), dtype=np.uint16), None, 'x')
    assert first.all() and not second.any()


def test_mask_and_label_kernels():
    rng = np.random.default_rng(1)
    n = 100003
    device = dev.current_device()
    for dtype in (np.uint16, np.uint8, np.int64):
        labels = rng.choice(np.array([0, 1, 9, 10, 40, 99]), size=n).astype(dtype)
        mask = rng.random(n) < 0.5
        d_labels = dev.upload_labels(labels, device)
        d_mask = dev.upload_mask(mask, n, device)
        assert np.array_equal(d_labels.cpu().numpy(), labels.astype(np.int32))
        got = dev.download_mask(dev.mask_build(d_labels, d_mask, eq=(10, 99), ne=(9,)))
        assert np.array_equal(got, (mask | (labels == 10) | (labels == 99)) & (labels != 9))
        got = dev.download_mask(dev.mask_build(d_labels, None, base=True, ne=(9, 1)))
        assert np.array_equal(got, (labels != 9) & (labels != 1))
        got = dev.download_mask(dev.mask_build(d_labels, None, base=False, eq=(0,)))
        assert np.array_equal(got, labels == 0)
        other = dev.upload_mask(rng.random(n) < 0.5, n, device)
        for op, f in (('and', np.logical_and), ('or', np.logical_or), ('andnot', lambda a, b: a & ~b)):
            assert np.array_equal(dev.download_mask(dev.mask_combine(d_mask, other, op)),
                                  f(mask, dev.download_mask(other)))
        assert dev.mask_count(d_mask) == mask.sum()
        dev.labels_assign(d_labels, d_mask, 62)
        want = labels.copy(); want[mask] = 62
        host = labels.copy()
        assert dev.download_labels(d_labels, host) is host and np.array_equal(host, want)
        assert np.array_equal(dev.label_hist(d_labels).cpu().numpy(), R.label_histogram(want))
    # the stream kernels take 4 / 16 elements per thread when the buffers are aligned: ragged sizes and views that
    # start at an odd element (unaligned pointers) must take the element-wise path to the same result
    import torch
    for n in (0, 1, 3, 17, 4099):
        for off in (0, 1, 3):
            labels = rng.choice(np.array([0, 1, 9, 10]), size=n + off).astype(np.int32)
            mask = rng.random(n + off) < 0.5
            other = rng.random(n + off) < 0.5
            d_l = torch.from_numpy(labels).to(device)[off:]
            d_m = dev.upload_mask(mask, n + off, device)[off:]
            d_o = dev.upload_mask(other, n + off, device)[off:]
            l, m, o = labels[off:], mask[off:], other[off:]
            assert np.array_equal(dev.download_mask(dev.mask_build(d_l, d_m, eq=(10,), ne=(9,))), (m | (l == 10)) & (l != 9))
            assert np.array_equal(dev.download_mask(dev.mask_combine(d_m, d_o, 'andnot')), m & ~o)
            assert dev.mask_count(d_m) == m.sum()
            dev.labels_assign(d_l, d_m, 7)
            want = l.copy(); want[m] = 7
            assert np.array_equal(d_l.cpu().numpy(), want)


def test_las_ingest_scaled_on_device_and_process_file(tmp_path):
    """LAS path: int32 X, Y, Z up, x = X * scale + offset on the device == the host-side scaled view;
    Pipeline.process_file == Pipeline.process_cloud on the same cloud."""
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    tile = synthetic.make_tile(200000, seed=77)
    pts = tile['points']
    scales, offsets = np.array([0.001, 0.001, 0.001]), np.array([100000.0, 400000.0, 0.0])
    P = np.rint((pts - offsets) / scales).astype(np.int32)
    las = las_io.create(P, scales, offsets)
    t = dev.DeviceTile.from_las(las)
    assert np.array_equal(t.x.cpu().numpy(), las.x) and np.array_equal(t.y.cpu().numpy(), las.y)
    assert np.array_equal(t.z.cpu().numpy(), las.z) and t.h2d_bytes == 12 * len(P)
    # awkward scales / offsets: the two roundings must match numpy's
    las2 = las_io.create(P, [0.00025, 0.01, 1.0 / 3.0], [0.1, -7.3, 1e6])
    t2 = dev.DeviceTile.from_las(las2)
    assert np.array_equal(t2.x.cpu().numpy(), las2.x) and np.array_equal(t2.z.cpu().numpy(), las2.z)
    f_in, f_out = str(tmp_path / f"filtered_{tile['tilecode']}.las"), str(tmp_path / f"labelled_{tile['tilecode']}.las")
    las.write(f_in)
    ahn, bld, road = readers_for(tile)

    def procs():
        return (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
                BGTRoadFuser(Labels.ROAD, bgt_reader=road),
                BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
                RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)))
    Pipeline(processors=procs(), ahn_reader=ahn, caching=True).process_file(f_in, f_out)
    out = las_utils.read_las(f_out)
    points = np.vstack((las.x, las.y, las.z)).T
    want = Pipeline(processors=procs(), ahn_reader=ahn, caching=True).process_cloud(
        tile['tilecode'], points, np.zeros(len(points), dtype=np.uint16))
    assert np.array_equal(out.label, want.astype(np.uint8)) and (want != 0).sum() > 100000
    assert np.array_equal(out.X, P[:, 0])


def test_process_folder_pipelined_equals_per_file(tmp_path):
    """Pipeline.process_folder (file i+1 read and file i-1 written on host threads while the device labels
    file i) writes exactly what one process_file call per file writes; prefixes / suffix as the reference."""
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    scales, offsets = np.array([0.001, 0.001, 0.001]), np.array([100000.0, 400000.0, 0.0])
    tiles, tcs = [], ['2386_9702', '2386_9703', '2387_9702']
    in_dir, out_a, out_b = tmp_path / 'in', tmp_path / 'a', tmp_path / 'b'
    in_dir.mkdir(); out_b.mkdir()
    for k, tc in enumerate(tcs):
        tile = synthetic.make_tile(60000 + 7000 * k, seed=300 + k)
        P = np.rint((tile['points'] - offsets) / scales).astype(np.int32)
        las_io.create(P, scales, offsets).write(str(in_dir / f'filtered_{tc}.las'))
        tiles.append(tile)
    (in_dir / 'notes.txt').write_text('ignored')
    ahn = ahn_utils.ArrayAHNReader({tc: t['ahn_tile'] for tc, t in zip(tcs, tiles)})
    bld = bgt_utils.ArrayBGTReader({tc: t['building_polygons'] for tc, t in zip(tcs, tiles)})
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
             RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,)))
    pipe = Pipeline(processors=procs, ahn_reader=ahn, caching=True)
    pipe.process_folder(str(in_dir), str(out_a), in_prefix='filtered_', out_prefix='labelled_')
    for tc in tcs:
        pipe.process_file(str(in_dir / f'filtered_{tc}.las'), str(out_b / f'labelled_{tc}.las'))
        a, b = las_utils.read_las(str(out_a / f'labelled_{tc}.las')), las_utils.read_las(str(out_b / f'labelled_{tc}.las'))
        assert np.array_equal(a.label, b.label) and np.array_equal(a.X, b.X) and (np.asarray(a.label) != 0).sum() > 10000
    assert sorted(f.name for f in out_a.iterdir()) == sorted(f'labelled_{tc}.las' for tc in tcs)
    out_c = tmp_path / 'c'
    pipe.process_folder(str(in_dir), str(out_c), in_prefix='filtered_', out_prefix='labelled_', workers=2)   # files in flight
    for tc in tcs:
        c, b = las_utils.read_las(str(out_c / f'labelled_{tc}.las')), las_utils.read_las(str(out_b / f'labelled_{tc}.las'))
        assert np.array_equal(c.label, b.label) and np.array_equal(c.X, b.X)
    assert pipe.process_folder(str(tmp_path / 'missing')) is None
    # the reference's own file type: `.laz` in, `.laz` out (native LASzip decoder / encoder, csrc/laz.cu)
    laz_in, laz_out = tmp_path / 'laz_in', tmp_path / 'laz_out'
    laz_in.mkdir()
    for tc in tcs:
        las_utils.read_las(str(in_dir / f'filtered_{tc}.las')).write(str(laz_in / f'filtered_{tc}.laz'))
        assert open(laz_in / f'filtered_{tc}.laz', 'rb').read()[104] & 0x80
    pipe.process_folder(str(laz_in), str(laz_out), in_prefix='filtered_', out_prefix='labelled_')
    for tc in tcs:
        z, b = las_utils.read_las(str(laz_out / f'labelled_{tc}.laz')), las_utils.read_las(str(out_b / f'labelled_{tc}.las'))
        assert open(laz_out / f'labelled_{tc}.laz', 'rb').read()[104] & 0x80            # written compressed
        assert np.array_equal(z.label, b.label) and np.array_equal(z.records, b.records)


# ------------------------------------------------------------------------ raster ---
def test_raster_lookup_matches_reference_golden(golden):
    g = golden('raster_golden.npz')
    ahn = np.load(os.path.join(GOLDEN, 'demo', 'ahn_2386_9702.npz'))
    for surface, key in (('ground_surface', 'ground'), ('building_surface', 'building')):
        got = FastGridInterpolator(ahn['x'], ahn['y'], ahn[key].astype(float))(g['points'])
        assert np.array_equal(got, g[surface], equal_nan=True)


def test_raster_classify_modes_match_oracle():
    tile = synthetic.make_tile(300000, seed=5)
    pts, ahn = tile['points'], tile['ahn_tile']
    t = dev.DeviceTile(pts)
    rng = np.random.default_rng(2)
    sel = rng.random(len(pts)) < 0.7
    d_sel = dev.upload_mask(sel, t.n, t.device)
    for surface in ('ground_surface', 'building_surface'):
        raster = dev.DeviceRaster(ahn['x'], ahn['y'], ahn[surface])
        v = R.fast_grid_lookup(ahn['x'], ahn['y'], ahn[surface], pts)
        assert np.array_equal(raster.lookup(t).cpu().numpy(), v, equal_nan=True)
        z = pts[:, 2]
        with np.errstate(invalid='ignore'):
            want = {(_lib.CLS_ABS_LT, 0.2, 0): np.abs(z - v) < 0.2, (_lib.CLS_LT, 0.2, 0): z < v + 0.2,
                    (_lib.CLS_CAP_LE, 0.2, 0): ~np.isfinite(v) | (z <= v + 0.2),
                    (_lib.CLS_BAND_OC, 0.5, 12.0): (z > v + 0.5) & (z <= v + 12.0),
                    (_lib.CLS_BAND_OC, -np.inf, 1.75): (z > v - np.inf) & (z <= v + 1.75),
                    (_lib.CLS_BAND_CC, 0.02, 0.5): (z <= v + 0.5) & (z >= v - 0.02),
                    (_lib.CLS_BELOW_NEG, 0.2, 0): (z - v) < -0.2}
        for (mode, a, b), w in want.items():
            got = dev.download_mask(raster.classify(t, mode, a=a, b=b, d_sel=d_sel))
            assert np.array_equal(got, w & sel), (surface, mode)
    # fused label write
    d_labels = torch.zeros(t.n, dtype=torch.int32, device=t.device)
    raster.classify(t, _lib.CLS_LT, a=0.2, d_sel=d_sel, d_labels=d_labels, assign=7, want_mask=False)
    with np.errstate(invalid='ignore'):
        assert np.array_equal(d_labels.cpu().numpy() == 7, (pts[:, 2] < v + 0.2) & sel)


# --------------------------------------------------------------------------- PIP ---
@pytest.mark.parametrize('name', PIP_CASES)
def test_pip_matches_reference_golden(golden, name):
    g = golden('pip_golden.npz')
    pts = np.ascontiguousarray(g[f'{name}__points'])
    expected = unpack(g[f'{name}__expected'], int(g[f'{name}__n']))
    polys = polygons_from_golden(g, name)
    t = dev.DeviceTile(pts)
    for grid_dim in (0, 7, 300):            # the acceleration grid must not change the result
        got = dev.download_mask(dev.PipPlan(polys, grid_dim=grid_dim).query(t))
        assert np.array_equal(got, expected), (name, grid_dim)
    one = np.zeros(len(pts), dtype=bool)
    for p in polys[:5]:
        one |= clip_utils.poly_clip(pts, p)          # the poly_clip drop-in, polygon by polygon
    ref = np.zeros(len(pts), dtype=bool)
    for p in polys[:5]:
        ref |= R.poly_clip(pts, p)
    assert np.array_equal(one, ref)


def test_buffered_polygons_match_distance_oracle():
    """polygon.buffer(offset) evaluated as a distance (upcp_pip_plan_create_buffered): star-shaped rings with
    holes, positive and negative offsets, probes on and around the offset curve -- bit for bit the numpy
    restatement of the same expressions.  (GEOS itself is in neither container: unpinned, exposure counted.)"""
    rng = np.random.default_rng(77)
    polys = synthetic.make_polygon_lattice(n_side=6, seed=3)
    demo = os.path.join(GOLDEN, 'demo')
    real = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_buildings_demo.csv'))
    for offset in (0.25, 0.05, -0.03):
        for name, plist in (('lattice', [p.buffer(offset) for p in polys]),
                            ('bgt', real.filter_tile('2386_9702', bgt_types=['pand'], offset=offset, merge=True))):
            assert all(abs(p.offset - offset) < 1e-15 for p in plist)
            rings = [np.array(p.exterior.coords) for p in plist]
            allv = np.vstack(rings)
            lo, hi = allv.min(0) - 1.0, allv.max(0) + 1.0
            pts = lo + rng.random((400_000, 2)) * (hi - lo)
            # probes hugging the offset curve: ring points pushed out / in by about |offset|
            k = rng.integers(0, len(allv), 100_000)
            ang = rng.uniform(0, 2 * np.pi, len(k))
            rad = abs(offset) * (1.0 + rng.normal(0, 0.02, len(k)))
            pts = np.vstack([pts, allv[k] + np.column_stack((rad * np.cos(ang), rad * np.sin(ang)))])
            pts = np.ascontiguousarray(np.round(pts, 3))
            want = np.zeros(len(pts), dtype=bool)
            exposed = 0
            for p in plist:
                m, e = R.poly_clip_buffered(pts, p, offset, return_exposure=True)
                want |= m
                exposed += e
            t = dev.DeviceTile(pts)
            got = dev.download_mask(dev.PipPlan(plist).query(t))
            assert np.array_equal(got, want), (name, offset, int((got != want).sum()))
            plain = dev.download_mask(dev.PipPlan([RingPolygon(np.array(p.exterior.coords),
                                                               [np.array(i.coords) for i in p.interiors]) for p in plist]).query(t))
            assert (got | plain).sum() == (got.sum() if offset > 0 else plain.sum())       # buffer grows / shrinks the set
            print(f'{name}, offset {offset}: {got.sum()} inside (unbuffered {plain.sum()}), {exposed} probes in the '
                  f'round-join sliver where GEOS (quad_segs 8) would differ')
    # the fusers accept the offset (notebook 0: BGTBuildingFuser(offset=0.25)) and agree with the oracle
    tile = synthetic.make_tile(200_000, seed=91)
    ahn, bld, road = readers_for(tile)
    labels = np.zeros(len(tile['points']), dtype=np.uint16)
    got = BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0.25, ahn_reader=ahn).get_labels(
        tile['points'], labels.copy(), labels == 0, tile['tilecode'])
    want = R.building_fuser_labels(tile['points'], labels.copy(), labels == 0,
                                   [p.buffer(0.25) for p in tile['building_polygons']], Labels.BUILDING, tile['ahn_tile'], 0.2)
    plain = R.building_fuser_labels(tile['points'], labels.copy(), labels == 0, tile['building_polygons'], Labels.BUILDING,
                                    tile['ahn_tile'], 0.2)
    assert np.array_equal(got, want) and (got == Labels.BUILDING).sum() > (plain == Labels.BUILDING).sum()


def test_ahn_fuser_refine_ground_matches_oracle():
    """AHNFuser(refine_ground=True) -- the reference's default (ahn_fuser.py:40-41,76-126): tentative ground under
    unlabelled clusters is removed again.  Device band test + clustering + clipping against all cluster hulls,
    host Delaunay per cluster; against the oracle's restatement of the same control flow (hull = outer edges
    of the alpha complex, pinned by alpha_golden.npz; buffer as a distance, unpinned against GEOS)."""
    tile = synthetic.make_tile(400_000, seed=63, facade_fraction=0.25)
    pts, tc = tile['points'], tile['tilecode']
    ahn, _, _ = readers_for(tile)
    for params in ({}, {'use_concave': False}, {'concave_alpha': 0.7, 'buffer': 0.2, 'min_comp_size': 20}):
        labels = np.zeros(len(pts), dtype=np.uint16)
        labels[::17] = 3                                       # some points arrive labelled (not UNKNOWN)
        mask = labels == 0
        want = R.ahn_fuser_labels(pts, labels.copy(), mask, tile['ahn_tile'], Labels.GROUND, 'ground', 0.2,
                                  refine_ground=True, params=dict(params))
        plain = R.ahn_fuser_labels(pts, labels.copy(), mask, tile['ahn_tile'], Labels.GROUND, 'ground', 0.2)
        got = AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=True,
                       params=dict(params)).get_labels(pts, labels.copy(), mask, tc)
        removed = int(((plain == Labels.GROUND) & (want != Labels.GROUND)).sum())
        print(f'refine_ground {params}: {removed} of {(plain == Labels.GROUND).sum()} tentative ground points removed, '
              f'{int((got != want).sum())} labels differ from the oracle')
        assert np.array_equal(got, want) and removed > 100


def test_pip_lattice_5000_polygons_matches_oracle():
    polys = synthetic.make_polygon_lattice()
    tile = synthetic.make_tile(400000, seed=9)
    pts = tile['points']
    rng = np.random.default_rng(3)
    sel = rng.random(len(pts)) < 0.9
    t = dev.DeviceTile(pts)
    got = dev.download_mask(dev.PipPlan(polys).query(t, dev.upload_mask(sel, t.n, t.device)))
    want = np.zeros(len(pts), dtype=bool)
    sub = pts[sel]
    acc = np.zeros(len(sub), dtype=bool)
    for p in polys:
        acc |= R.poly_clip(sub, p)
    want[sel] = acc
    assert np.array_equal(got, want) and want.sum() > 1000


# -------------------------------------------------------------------------- sort ---
@pytest.mark.parametrize('n', [1, 31, 4096, 4097, 300001])
@pytest.mark.parametrize('bits,dtype', [(30, np.uint32), (32, np.uint32), (7, np.uint32), (63, np.uint64), (33, np.uint64)])
def test_radix_sort_pairs(n, bits, dtype):
    lib = _lib.load()
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 2 ** bits, size=n, dtype=np.uint64).astype(dtype)
    keys[::3] = keys[0]                                # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    device = dev.current_device()
    tdt = torch.int32 if dtype == np.uint32 else torch.int64
    k_in = torch.from_numpy(keys.view(np.int32 if dtype == np.uint32 else np.int64)).to(device)
    v_in = torch.from_numpy(vals.view(np.int32)).to(device)
    k_out = torch.empty(n, dtype=tdt, device=device); v_out = torch.empty(n, dtype=torch.int32, device=device)
    ws = torch.empty(lib.upcp_sort_ws_bytes(n), dtype=torch.uint8, device=device)
    fn = lib.upcp_sort_pairs_u32 if dtype == np.uint32 else lib.upcp_sort_pairs_u64
    _lib.check(fn(dev.ptr(k_in), dev.ptr(v_in), dev.ptr(k_out), dev.ptr(v_out), n, bits, dev.ptr(ws), ws.numel(),
                  dev.stream_ptr()))
    order = np.argsort(keys, kind='stable')
    assert np.array_equal(k_out.cpu().numpy().view(dtype), keys[order])
    assert np.array_equal(v_out.cpu().numpy().view(np.uint32), vals[order])


# ---------------------------------------------------------------------------- CC ---
@pytest.mark.parametrize('grid_size,min_size', [(0.4, 20), (0.1, 100), (0.05, 10), (0.02, 5)])
def test_lcc_components_match_oracle(grid_size, min_size):
    tile = synthetic.make_tile(200000, seed=21)
    pts = tile['points'][tile['kind'] != 0]            # non-ground cloud (config 2 shape)
    lcc = LabelConnectedComp(grid_size=grid_size, min_component_size=min_size)
    got = lcc.get_components(pts)
    want = R.lcc_get_components(pts, grid_size, min_size)
    assert lcc.octree_level == R.get_octree_level(pts, grid_size)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.array_equal(R.canonical_components(got), R.canonical_components(want))
    assert (want >= 0).any() and (want < 0).any()


def test_lcc_label_mask_matches_oracle_and_mutates_labels():
    tile = synthetic.make_tile(250000, seed=22)
    pts = tile['points']
    rng = np.random.default_rng(4)
    labels = np.zeros(len(pts), dtype=np.uint16)
    labels[tile['kind'] == 0] = Labels.GROUND
    cars = np.where(tile['kind'] == 2)[0]
    labels[cars[rng.random(len(cars)) < 0.05]] = Labels.CAR
    for kwargs in (dict(grid_size=0.1, min_component_size=50, threshold=0.01),
                   dict(grid_size=0.05, min_component_size=100, threshold=0.1),     # nothing qualifies
                   dict(grid_size=0.2, min_component_size=20, threshold=0.02),
                   dict(exclude_labels=[Labels.GROUND], grid_size=0.2, min_component_size=10, threshold=0.03)):
        mask = (labels == 0) & (rng.random(len(pts)) < 0.9)
        l1, l2 = labels.copy(), labels.copy()
        got = LabelConnectedComp(Labels.CAR, **kwargs).get_label_mask(pts, l1, mask)
        want = R.lcc_get_label_mask(pts, l2, mask, Labels.CAR, **kwargs)
        assert np.array_equal(got, want)
        assert want.any() or kwargs['grid_size'] == 0.05
        assert np.array_equal(l1 == Labels.CAR, (labels == Labels.CAR) | want)     # :128-129 side effect


def test_lcc_edge_cases():
    lcc = LabelConnectedComp(7, grid_size=0.1, min_component_size=1, threshold=0.0)
    # 2-D input (z assumed 0), a single point, an empty selection, identical points
    pts2 = np.random.default_rng(5).random((5000, 2)) * 3 + np.array([119300.0, 485100.0])
    got = LabelConnectedComp(grid_size=0.1, min_component_size=3).get_components(pts2)
    assert np.array_equal(R.canonical_components(got), R.canonical_components(R.lcc_get_components(pts2, 0.1, 3)))
    one = np.array([[119300.5, 485100.5, 1.0]])
    assert R.canonical_components(lcc.get_components(one)).tolist() == [0]
    same = np.repeat(one, 50, axis=0)
    assert (R.canonical_components(LabelConnectedComp(grid_size=0.1, min_component_size=10).get_components(same)) == 0).all()
    labels = np.zeros(5, dtype=np.uint16)
    m = lcc.get_label_mask(np.repeat(one, 5, axis=0) + np.arange(5)[:, None], labels, np.zeros(5, dtype=bool))
    assert not m.any()


def test_lcc_permutation_invariance_and_level_boundaries():
    tile = synthetic.make_tile(120000, seed=23)
    pts = tile['points'][tile['kind'] >= 2]
    perm = np.random.default_rng(6).permutation(len(pts))
    lcc = LabelConnectedComp(grid_size=0.2, min_component_size=15)
    a = lcc.get_components(pts)
    b = lcc.get_components(pts[perm])
    # same partition of the same points, whatever the input order (b[j] belongs to point perm[j])
    b_orig = np.empty_like(b); b_orig[perm] = b
    assert np.array_equal(R.canonical_components(a), R.canonical_components(b_orig))
    assert (a >= 0).any() and (a < 0).any()


# ------------------------------------------------------------- pipeline (golden) ---
def test_pipeline_stages_match_reference_golden(golden):
    g = golden('pipeline_golden.npz')
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']))
    points = tile['points']
    assert np.array_equal(points, from_mm(g['points_mm']))
    tc = tile['tilecode']
    ahn, bld, road = readers_for(tile, drop_holes=True)
    stages = [
        ('ahn_ground', AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False)),
        ('bgt_road', BGTRoadFuser(Labels.ROAD, bgt_reader=road)),
        ('bgt_building', BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn)),
        ('layer_lcc', LayerLCC(Labels.BUILDING, ahn, params=[{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                                                             {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}])),
        ('lcc_car', LabelConnectedComp(Labels.CAR, grid_size=0.1, min_component_size=30, threshold=0.02)),
        ('noise', NoiseFilter(Labels.NOISE, ahn, epsilon=0.2, min_component_size=10)),
        ('region_growing', RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD))),
    ]
    labels = np.zeros(len(points), dtype=np.uint16)
    for name, proc in stages:
        if name == 'lcc_car':
            labels[g['car_seed_ids']] = Labels.CAR
        out = proc.get_labels(points, labels, labels == 0, tc)
        assert out is labels                                        # in-place contract
        assert np.array_equal(labels, g[f'labels_after_{name}']), name
    comps = LabelConnectedComp(grid_size=0.2, min_component_size=20).get_components(points[tile['kind'] >= 2])
    assert np.array_equal(R.canonical_components(comps), g['components_kind_ge2'])


def test_pipeline_process_cloud_matches_stagewise(golden):
    g = golden('pipeline_golden.npz')
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']))
    ahn, bld, road = readers_for(tile, drop_holes=True)
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
             LayerLCC(Labels.BUILDING, ahn, params=[{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                                                    {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]))
    labels = np.zeros(int(g['n']), dtype=np.uint16)
    out = Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud(tile['tilecode'], tile['points'], labels)
    assert out is labels and np.array_equal(labels, g['labels_after_layer_lcc'])
    with pytest.raises(ValueError):
        Pipeline(processors=procs, ahn_reader=None, caching=True)


def test_process_clouds_batch_equals_per_tile_calls():
    """Pipeline.process_clouds (upload of tile i+1 overlapped with the kernels of tile i) labels every
    tile exactly as one process_cloud call per tile does: different tiles and sizes, pinned / pageable,
    C- / F-ordered host arrays, in-place contract per tile."""
    tiles = [synthetic.make_tile(n, seed=20240000 + k) for k, n in enumerate((90000, 150000, 40000, 150001))]
    tcs = [f'{1000 + k}_{2000 + k}' for k in range(len(tiles))]             # distinct tilecodes per tile
    ahn = ahn_utils.ArrayAHNReader({tc: t['ahn_tile'] for tc, t in zip(tcs, tiles)})
    bld = bgt_utils.ArrayBGTReader({tc: t['building_polygons'] for tc, t in zip(tcs, tiles)})
    road = bgt_utils.ArrayBGTReader({tc: t['road_polygons'] for tc, t in zip(tcs, tiles)})
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
             LayerLCC(Labels.BUILDING, ahn, params=[{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                                                    {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]),
             RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)))
    pipe = Pipeline(processors=procs, ahn_reader=ahn, caching=False)
    clouds = []
    for k, t in enumerate(tiles):
        p = t['points']
        if k % 2 == 0:                                                        # pinned, F-ordered (pipeline.py:124)
            buf = torch.empty((3, len(p)), dtype=torch.float64).pin_memory()
            buf.numpy()[...] = p.T
            clouds.append(buf.numpy().T)
        else:                                                                 # pageable, C-ordered
            clouds.append(np.ascontiguousarray(p))
    want = [pipe.process_cloud(tc, c, np.zeros(len(c), dtype=np.uint16)) for tc, c in zip(tcs, clouds)]
    dev.clear_tile_cache()
    labels = [np.zeros(len(c), dtype=np.uint16 if k != 1 else np.uint8) for k, c in enumerate(clouds)]
    for workers in (2, 1, 3):                   # tiles in flight: worker threads + streams / one stream + copy stream
        labels = [np.zeros(len(c), dtype=np.uint16 if k != 1 else np.uint8) for k, c in enumerate(clouds)]
        out = pipe.process_clouds(tcs, clouds, labels, workers=workers)
        assert all(o is l for o, l in zip(out, labels))
        for k in range(len(tiles)):
            assert np.array_equal(labels[k], want[k]), (workers, k)
            assert (labels[k] != 0).sum() > 1000
    # device-resident batch, the same tile several times over (shared read-only inputs)
    d_tiles = [dev.DeviceTile(c) for c in clouds]
    reps = [0, 1, 2, 3, 0, 0, 2]
    d_labels = [torch.zeros(d_tiles[k].n, dtype=torch.int32, device=d_tiles[k].device) for k in reps]
    pipe.process_tiles_device([tcs[k] for k in reps], [d_tiles[k] for k in reps], d_labels, workers=3)
    for k, d in zip(reps, d_labels):
        assert np.array_equal(d.cpu().numpy(), want[k].astype(np.int32))
    # a failing tile surfaces in the caller
    with pytest.raises(Exception):
        pipe.process_clouds(tcs[:2] + ['0000_0000'], clouds[:3], [np.zeros(len(c), np.uint16) for c in clouds[:3]], workers=2)
    assert pipe.process_clouds([], [], []) == []
    with pytest.raises(ValueError):
        pipe.process_clouds(tcs[:2], clouds[:1], labels[:2])


# ------------------------------------------------ object fusers ("next" row 2) ---
def test_component_stats_match_numpy():
    """upcp_component_stats: grouping, per-cluster reductions and the mean of the finite ground values
    against plain numpy on the oracle-independent device clustering."""
    from urban_pointcloud_processing_b200.region_growing.label_connected_comp import lcc_device
    from urban_pointcloud_processing_b200.utils import object_utils
    tile = synthetic.make_tile(400000, seed=91)
    pts = tile['points']
    t = dev.DeviceTile(pts)
    sel = tile['kind'] >= 1
    d_sel = dev.upload_mask(sel, t.n, t.device)
    gz = R.interpolate(tile['ahn_tile'], pts, 'ground_surface')
    gz[~sel] = np.nan
    d_gz = ahn_utils.device_raster_for(ahn_utils.ArrayAHNReader({'t': tile['ahn_tile']}), 't', 'ground_surface').lookup(t, d_sel=d_sel)
    assert np.array_equal(d_gz.cpu().numpy(), gz, equal_nan=True)
    d_comp, _, _ = lcc_device(t, d_sel, 0.2, 50, level_from='sel')
    comp = d_comp.cpu().numpy()
    for cap in (1 << 16, 3):                                                  # 3: forces the capacity retry
        st = object_utils.ComponentStats(t, d_comp, d_gz, cap=cap)
        ids = np.unique(comp[comp >= 0])
        assert st.n_components == len(ids) > 20 and st.n_points == (comp >= 0).sum() and st.n_gz_out_of_range == 0
        dense = st.d_dense.cpu().numpy()
        assert np.array_equal(dense >= 0, comp >= 0) and np.array_equal(ids[dense[comp >= 0]], comp[comp >= 0])
        order = st.d_order.cpu().numpy().view(np.uint32)
        for k in list(range(0, len(ids), 7)) + [len(ids) - 1]:
            m = comp == ids[k]
            assert np.array_equal(order[st.seg_first[k]:st.seg_first[k + 1]], np.where(m)[0])
            assert st.size[k] == m.sum() and st.max_z[k] == pts[m, 2].max()
            assert np.array_equal(st.bbox[k], [pts[m, 0].min(), pts[m, 1].min(), pts[m, 0].max(), pts[m, 1].max()])
            fin = gz[m][np.isfinite(gz[m])]
            assert st.count_finite[k] == len(fin)
            if len(fin):
                assert st.mean_gz[k] == np.mean(fin)                          # bit-identical (float16-derived values)
            else:
                assert np.isnan(st.mean_gz[k])
            assert np.array_equal(st.xy_of(k), pts[m][:, :2])
        acc = np.zeros(len(ids), dtype=bool); acc[::3] = True
        assert np.array_equal(dev.download_mask(st.mark(acc)), np.isin(comp, ids[acc]))
    # no clusters at all / empty tile
    none = torch.full((t.n,), -1, dtype=torch.int32, device=t.device)
    st = object_utils.ComponentStats(t, none, d_gz)
    assert st.n_components == 0 and not st.mark(np.zeros(0, bool)).any()


def test_prism_clip_matches_oracle():
    from urban_pointcloud_processing_b200.utils import object_utils
    tile = synthetic.make_tile(300000, seed=92)
    pts = tile['points']
    t = dev.DeviceTile(pts)
    rng = np.random.default_rng(9)
    sel = rng.random(len(pts)) < 0.8
    rings, zr = [], []
    for k in range(700):                                                      # > one shared-memory chunk of prisms
        c, a = rng.uniform(2, 48, 2) + np.array(tile['origin']), rng.uniform(0, np.pi)
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        box = (np.array([[-1.1, -0.5], [1.1, -0.5], [1.1, 0.5], [-1.1, 0.5], [-1.1, -0.5]]) @ rot.T) + c
        if k % 50 == 0:                                                       # vertices / edges through cloud points
            box[0] = box[4] = pts[rng.integers(len(pts)), :2]
        rings.append(box); zr.append((rng.uniform(-1, 2), rng.uniform(2, 9)))
    zr[3] = (pts[5, 2], pts[5, 2])                                            # closed z interval
    got = dev.download_mask(object_utils.prism_clip_device(t, dev.upload_mask(sel, t.n, t.device), rings, zr))
    want = np.zeros(len(pts), dtype=bool)
    for ring, (b, tp) in zip(rings, zr):
        want |= R.poly_clip(pts, RingPolygon(ring)) & ((pts[:, 2] <= tp) & (pts[:, 2] >= b))
    want &= sel
    assert np.array_equal(got, want) and want.sum() > 1000
    assert not object_utils.prism_clip_device(t, None, [], []).any()


def test_object_fusers_match_reference_golden(golden):
    """CarFuser / BGTStreetFurnitureFuser drop-ins vs the REAL reference classes (objects_golden.npz)."""
    from test_oracle_golden import OBJ_CAR_PARAMS, OBJ_POLE_PARAMS, objects_tile
    from urban_pointcloud_processing_b200.fusion import BGTStreetFurnitureFuser, CarFuser
    g = golden('objects_golden.npz')
    tile, poles = objects_tile(g)
    points, tc = tile['points'], tile['tilecode']
    ahn, _, road = readers_for(tile)
    pole_reader = bgt_utils.ArrayBGTPointReader({tc: poles})
    labels = np.zeros(len(points), dtype=np.uint16)
    labels = AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False).get_labels(
        points, labels, labels == 0, tc)
    assert np.array_equal(labels, g['labels_after_ground'])
    car = CarFuser(Labels.CAR, ahn, road, grid_size=0.1, min_component_size=300, params=dict(OBJ_CAR_PARAMS),
                   exclude_types=['fietspad'])
    out = car.get_labels(points, labels, labels == 0, tc)
    assert out is labels and np.array_equal(labels, g['labels_after_car']) and len(car.last_objects) >= 3
    sf = BGTStreetFurnitureFuser(Labels.STREET_LIGHT, 'lichtmast', pole_reader, ahn, grid_size=0.1,
                                 min_component_size=100, params=dict(OBJ_POLE_PARAMS))
    labels = sf.get_labels(points, labels, labels == 0, tc)
    assert np.array_equal(labels, g['labels_after_street_light']) and sf.last_accepted >= 3
    # nothing to do: no road parts / no point objects -> labels unchanged
    before = labels.copy()
    CarFuser(Labels.CAR, ahn, bgt_utils.ArrayBGTReader({}), params=dict(OBJ_CAR_PARAMS)).get_labels(
        points, labels, labels == 0, tc)
    BGTStreetFurnitureFuser(Labels.TREE, 'geen', pole_reader, ahn, params=dict(OBJ_POLE_PARAMS)).get_labels(
        points, labels, labels == 0, tc)
    assert np.array_equal(labels, before)
    # chained on the device by the Pipeline like every other drop-in
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False), car, sf)
    piped = Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud(tc, points, np.zeros(len(points), np.uint16))
    assert np.array_equal(piped, g['labels_after_street_light'])


def test_parameter_variants_match_reference_golden(golden):
    """variants_golden.npz: non-default constructor arguments through the REAL reference classes -- AHNFuser on the
    building surface, BGTBuildingFuser without the AHN cap, BGTRoadFuser with a type list, LabelConnectedComp with
    an exclusion list, LayerLCC(reset_noise=True), RegionGrowing with other angles / neighbourhoods / a curvature
    threshold inside the range of the ratios (host LAPACK path)."""
    from test_oracle_golden import VAR_LAYER, VAR_LCC, VAR_RG
    g = golden('variants_golden.npz')
    tile = synthetic.make_tile(int(g['n']), seed=int(g['seed']), facade_fraction=0.5)
    points, tc = tile['points'], tile['tilecode']
    assert np.array_equal(points, from_mm(g['points_mm']))
    ahn, bld, road = readers_for(tile, drop_holes=True)

    def run(proc, labels, mask='unlabelled'):
        labels = labels.copy()
        out = proc.get_labels(points, labels, (labels == 0) if mask == 'unlabelled' else mask, tc)
        assert out is labels
        return labels
    zero = np.zeros(len(points), dtype=np.uint16)
    got = run(AHNFuser(Labels.BUILDING, ahn, target='building', epsilon=0.1, refine_ground=False), zero)
    assert np.array_equal(got, g['ahn_building'])
    ground = run(AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.3, refine_ground=False), zero)
    assert np.array_equal(ground, g['ahn_ground'])
    nocap = run(BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0), ground)
    assert np.array_equal(nocap, g['building_no_ahn'])
    got = run(BGTRoadFuser(Labels.ROAD, bgt_reader=road, bgt_types=['fietspad']), ground)
    assert np.array_equal(got, g['road_fietspad'])
    seeded = ground.copy()
    seeded[g['car_seed_ids']] = Labels.CAR
    got = run(LabelConnectedComp(Labels.CAR, **VAR_LCC), seeded)
    assert np.array_equal(got, g['lcc_exclude'])
    noisy = nocap.copy()
    noisy[g['noise_ids']] = Labels.NOISE
    got = run(LayerLCC(Labels.BUILDING, ahn, reset_noise=True, params=[dict(p) for p in VAR_LAYER]), noisy)
    assert np.array_equal(got, g['layer_lcc_reset_noise'])
    for k, kw in enumerate(VAR_RG):
        got = run(RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,), **kw), nocap, mask=None)
        assert np.array_equal(got, g[f'region_growing_{k}']), kw


def test_demo_tile_matches_reference_golden(golden):
    """configs[0] substitute: REAL demo AHN + BGT data of tile 2386_9702, reference pipeline."""
    g = golden('demo_golden.npz')
    points = np.asfortranarray(from_mm(g['points_mm']))
    demo = os.path.join(GOLDEN, 'demo')
    ahn = ahn_utils.NPZReader(demo)
    bld = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_buildings_demo.csv'))
    road = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_roads_demo.csv'))
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn))
    labels = np.zeros(len(points), dtype=np.uint16)
    Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud('2386_9702', points, labels)
    assert np.array_equal(labels, g['labels'])
    # the reader's own interpolate / cache path
    gz = ahn.interpolate('2386_9702', points, None, 'ground_surface')
    t = ahn.filter_tile('2386_9702')
    assert np.array_equal(gz, R.fast_grid_lookup(t['x'], t['y'], t['ground_surface'], points), equal_nan=True)


def test_config1_demo_tile_4m_matches_reference_golden(golden):
    """BASELINE config 1 as SURVEY 8(d) specifies it: 4 M synthetic points over the REAL demo AHN raster (NaN
    holes) and BGT footprints; AHN ground -> BGT road -> BGT building -> LayerLCC through Pipeline, then
    RegionGrowing -- bit for bit the labels of the reference's own classes (tests/golden/config1_golden.npz,
    oracle/make_golden.py config1).  offset = 0 / refine_ground = False: shapely is in neither container."""
    from conftest import CONFIG1_LAYERS, config1_cloud
    g = golden('config1_golden.npz')
    points, _ = config1_cloud(g)
    demo = os.path.join(GOLDEN, 'demo')
    ahn = ahn_utils.NPZReader(demo)
    bld = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_buildings_demo.csv'))
    road = bgt_utils.BGTPolyReader(bgt_file=os.path.join(demo, 'bgt_roads_demo.csv'))
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
             LayerLCC(Labels.BUILDING, ahn, params=[dict(p) for p in CONFIG1_LAYERS]))
    labels = np.zeros(len(points), dtype=np.uint16)
    out = Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud('2386_9702', points, labels)
    assert out is labels and np.array_equal(labels, g['labels_pipeline'])
    rg = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,))
    grown = rg.get_labels(points, labels.copy(), labels == 0, '2386_9702')
    n_diff = int(np.count_nonzero(grown != g['labels_region_growing']))
    print(f'config 1: {len(points)} points, region growing added {np.count_nonzero(grown == 10) - np.count_nonzero(labels == 10)} '
          f'points, {n_diff} labels differ from the reference, {rg.last_stats["n_curvature_on_host"]} curvature flags on the host')
    assert n_diff == 0


def test_shared_npz_reader_with_tiles_in_flight(tmp_path):
    """One real NPZReader shared by the worker threads of process_clouds, every worker on ITS OWN tilecode
    (the single-slot tile cache used to hand tile A's raster to tile B)."""
    codes = ['2386_9702', '2387_9702', '2388_9702', '2389_9702']
    tiles = []
    for k, code in enumerate(codes):
        ox = 119300.0 + 50.0 * k
        t = synthetic.make_tile(150_000, seed=777 + k, origin=(ox, 485100.0))
        assert t['tilecode'] == code
        a = t['ahn_tile']
        np.savez(tmp_path / f'ahn_{code}.npz', x=a['x'], y=a['y'], ground=(a['ground_surface'] + 0.7 * k).astype(np.float16),
                 building=a['building_surface'].astype(np.float16))
        t['points'][:, 2] += 0.7 * k                   # every tile has its own ground level: a foreign raster labels nothing
        t['points'][:, 2] = np.round(t['points'][:, 2] * 1000.0) * 0.001
        tiles.append(t)
    reader = ahn_utils.NPZReader(tmp_path, caching=True)
    bld = bgt_utils.ArrayBGTReader({t['tilecode']: t['building_polygons'] for t in tiles})
    procs = lambda: (AHNFuser(Labels.GROUND, reader, target='ground', epsilon=0.2, refine_ground=False),    # noqa: E731
                     BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=reader),
                     LayerLCC(Labels.BUILDING, reader, params=[{'bottom': 0., 'top': 12., 'grid_size': 0.1, 'threshold': 0.5}]))
    want = []
    for t in tiles:                                    # one tile at a time
        lab = np.zeros(len(t['points']), dtype=np.uint16)
        want.append(Pipeline(processors=procs(), ahn_reader=reader, caching=True).process_cloud(t['tilecode'], t['points'], lab))
        assert np.count_nonzero(lab == Labels.GROUND) > 30_000
    for rep in range(3):
        batch = tiles * 2
        got = Pipeline(processors=procs(), ahn_reader=reader, caching=True).process_clouds(
            [t['tilecode'] for t in batch], [t['points'] for t in batch],
            [np.zeros(len(t['points']), dtype=np.uint16) for t in batch], workers=4)
        for k, lab in enumerate(got):
            assert np.array_equal(lab, want[k % len(tiles)]), (rep, k)


def test_ahn_preprocessing_matches_reference_golden(golden, tmp_path):
    """AHN cloud -> ground (idw, 8 nearest within 1 m) and building (max within 0.5 m) rasters,
    against the REAL SpatialInterpolator (scipy cKDTree + the reference's Python loop; ahn_preprocessing_golden.npz).
    The stored float16 rasters must be identical; the fp64 values before rounding agree to ~1e-13 (BLAS dot /
    pairwise sum vs sequential accumulation), so a cell could only differ on a rounding knife edge -- counted."""
    from urban_pointcloud_processing_b200.preprocessing import ahn_preprocessing as AP
    from urban_pointcloud_processing_b200.utils import las_io
    from urban_pointcloud_processing_b200.utils.interpolation import SpatialInterpolator
    g = golden('ahn_preprocessing_golden.npz')
    las = las_io.create(np.column_stack((g['X'], g['Y'], g['Z'])), [0.001] * 3, [0.0] * 3, point_format_id=1)
    las.records[:, 15] = g['classification']
    f = str(tmp_path / 'ahn_2386_9702.las')
    las.write(f)
    out = np.load(AP.process_ahn_las_tile(f, out_folder=str(tmp_path / 'out')))
    assert out['x'].shape == (500,) and out['y'][0] > out['y'][-1] and out['ground'].dtype == np.float16
    for name in ('ground', 'building'):
        same = (out[name] == g[name]) | (np.isnan(out[name]) & np.isnan(g[name]))
        scaled = g[name + '_f64'] * 100.0
        knife = np.abs(np.abs(scaled - np.floor(scaled)) - 0.5) < 1e-6
        print(f'{name}: {int((~same).sum())} of {same.size} raster cells differ from the reference ({int(knife.sum())} cells sit '
              f'within 1e-8 m of a rounding edge), NaN share {np.isnan(out[name]).mean():.3f}')
        assert (same | knife).all() and (~same).sum() <= 3
    # the interpolator itself, before rounding, incl. other neighbour counts / powers
    x, y, z = g['X'] * 0.001, g['Y'] * 0.001, g['Z'] * 0.001
    m = g['classification'] == 2
    pos = np.column_stack((119300.05 + 0.1 * np.arange(500), np.full(500, 485120.05)))
    got = SpatialInterpolator(np.column_stack((x, y))[m], z[m], method='idw')(pos, n_neighbors=8, max_dist=1, power=2.)
    want = g['ground_f64'][299, :]                    # row of y = 485120.05
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok) and np.abs(got[ok] - want[ok]).max() < 1e-11
    from scipy.spatial import cKDTree
    tree = cKDTree(np.column_stack((x, y))[m])
    for k, power in ((3, 1.0), (12, 2.0)):
        got = SpatialInterpolator(np.column_stack((x, y))[m], z[m], method='idw')(pos, n_neighbors=k, max_dist=0.7, power=power)
        d, i = tree.query(pos, k=k, distance_upper_bound=0.7)
        want = np.full(len(pos), np.nan)
        for r in range(len(pos)):
            v = np.isfinite(d[r])
            if v.any():
                w = 1.0 / (d[r][v] ** power)
                want[r] = z[m][i[r][v][0]] if (d[r][v] <= 1e-12).any() else np.dot(w, z[m][i[r][v]]) / w.sum()
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok) and np.abs(got[ok] - want[ok]).max() < 1e-11


def test_ahn_preprocessing_of_the_demo_laz_equals_the_shipped_npz(tmp_path):
    """The reference ships datasets/ahn/ahn_2386_9702.laz AND the ahn_2386_9702.npz it made from it
    (notebooks/1.AHN_preprocessing).  Native LAZ decode (csrc/laz.cu) + process_ahn_las_tile on the GPU must
    reproduce that file: x, y, ground, building, all 250 000 float16 cells of both surfaces."""
    from urban_pointcloud_processing_b200.preprocessing import ahn_preprocessing as AP
    demo = os.path.join(GOLDEN, 'demo')
    out = np.load(AP.process_ahn_las_tile(os.path.join(demo, 'ahn_2386_9702.laz'), out_folder=str(tmp_path)))
    ref = np.load(os.path.join(demo, 'ahn_2386_9702.npz'))
    assert np.array_equal(out['x'], ref['x']) and np.array_equal(out['y'], ref['y'])
    for name in ('ground', 'building'):
        assert out[name].dtype == ref[name].dtype == np.float16
        same = (out[name] == ref[name]) | (np.isnan(out[name]) & np.isnan(ref[name]))
        print(f'{name}: {int((~same).sum())} of {same.size} cells differ from the raster the reference shipped')
        assert same.all()


def test_layer_lcc_dense_matches_oracle():
    """A cloud dense enough for the layers to actually grow (the 60 k golden tile is sparse)."""
    tile = synthetic.make_tile(1500000, seed=31, facade_fraction=0.6)
    pts = tile['points']
    ahn, bld, road = readers_for(tile)
    labels = np.zeros(len(pts), dtype=np.uint16)
    labels = R.ahn_fuser_labels(pts, labels, labels == 0, tile['ahn_tile'], 9, 'ground', 0.2)
    labels = R.building_fuser_labels(pts, labels, labels == 0, tile['building_polygons'], 10, tile['ahn_tile'], 0.2)
    params = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
              {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
    want = R.layer_lcc_labels(pts, labels.copy(), labels == 0, tile['ahn_tile'], 10, [dict(p) for p in params])
    got = LayerLCC(10, ahn, params=[dict(p) for p in params]).get_labels(pts, labels.copy(), labels == 0, tile['tilecode'])
    assert np.array_equal(got, want)
    assert np.count_nonzero(want == 10) > np.count_nonzero(labels == 10)


# --------------------------------------------------------------- kNN / normals / RG ---
def _rg_cloud(n, seed):
    tile = synthetic.make_tile(n, seed=seed, facade_fraction=0.7)
    pts = tile['points']
    labels = np.zeros(len(pts), dtype=np.uint16)
    labels = R.ahn_fuser_labels(pts, labels, labels == 0, tile['ahn_tile'], 9, 'ground', 0.2)
    labels = R.building_fuser_labels(pts, labels, labels == 0, tile['building_polygons'], 10, tile['ahn_tile'], 0.2)
    return tile, pts, labels


@pytest.mark.parametrize('cell', [None, 0.05, 0.37])
def test_knn_normals_cov_match_oracle(cell):
    tile, pts, labels = _rg_cloud(150000, 41)
    sel = labels != 9
    coords = np.ascontiguousarray(pts[sel])
    t = dev.DeviceTile(pts)
    d_sel = dev.upload_mask(sel, t.n, t.device)
    _, bbox_sel, m = t.bbox(d_sel)
    d_knn, d_d2, d_nrm, d_cov, d_curv = rgmod.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, cell_size=cell,
                                                                 want_d2=True)
    want_idx, want_d2 = natives.knn(coords, 15, want_d2=True)
    assert np.array_equal(d_knn.cpu().numpy(), want_idx)
    assert np.array_equal(d_d2.cpu().numpy(), want_d2)
    assert np.array_equal(d_cov.cpu().numpy(), natives.cov_knn(coords, want_idx))
    want_n = natives.normals_hybrid(coords, 0.2, 30)
    got_n = d_nrm.cpu().numpy()
    bad = np.abs(got_n - want_n).max(axis=1) > 1e-9
    assert bad.sum() == 0, f'{bad.sum()} normals differ by more than 1e-9'
    print(f'normals bit-identical: {np.mean((got_n == want_n).all(axis=1)):.4f}')
    report_tie_exposure(coords)
    check_normals_independently(t, d_sel, m, bbox_sel, coords, got_n, 0.2, 30, cell)


def report_tie_exposure(coords, k_row=15, k_nn=30):
    """How exposed the fixture is to the one place where the restated neighbour order could differ from
    nanoflann's: exact squared-distance ties.  A tie INSIDE a list only permutes two entries; a tie
    ACROSS the cut (rank k | k + 1) changes which point is in the list."""
    _, d2 = natives.knn(coords, k_nn + 1, want_d2=True)
    cut_row = int((d2[:, k_row - 1] == d2[:, k_row]).sum())
    cut_nn = int((d2[:, k_nn - 1] == d2[:, k_nn]).sum())
    in_row = int((d2[:, 1:k_row] == d2[:, :k_row - 1]).any(axis=1).sum())
    in_nn = int((d2[:, 1:k_nn] == d2[:, :k_nn - 1]).any(axis=1).sum())
    print(f'exact d2 ties over {len(coords)} queries: across the cut at rank {k_row}: {cut_row}, at rank {k_nn}: {cut_nn}; '
          f'inside the {k_row}-row: {in_row}, inside the {k_nn} nearest: {in_nn}')
    return cut_row, cut_nn, in_row, in_nn


def check_normals_independently(t, d_sel, m, bbox_sel, coords, normals, radius, max_nn, cell=None):
    """The device normals against an eigen-solver that shares NOTHING with upcp_math.cuh / upcp_oracle.c (both
    transcribe Open3D's analytic FastEigen3x3): LAPACK eigh on the SAME covariance.  The covariance is Open3D's
    single-pass cumulant form over the hybrid-search neighbours in ascending order -- re-derived here with
    numpy (cumsum adds sequentially, in neighbour order) from the device's own 30-neighbour rows, so the rows,
    the hybrid cut and the summation are checked on the way.  (A two-pass covariance is NOT comparable: with
    coordinates ~ 5e5 the single-pass form carries ~ eps |x|^2 ~ 5e-5 m^2 of rounding noise, part of what
    the reference computes.)  The smallest eigenvector is defined up to sign and as well as the gap allows."""
    d_knn, d_d2, d_nrm, _, _ = rgmod.knn_normals_device(t, d_sel, m, bbox_sel, max_nn, max_nn, radius, cell_size=cell,
                                                       want_d2=True, want_cov=False, want_curv=False)
    assert np.array_equal(d_nrm.cpu().numpy(), normals)                       # normals do not depend on knn
    idx, d2 = d_knn.cpu().numpy().astype(np.int64), d_d2.cpu().numpy()
    inside = (d2 < radius * radius) & (idx >= 0)
    assert (np.diff(inside.astype(np.int8), axis=1) <= 0).all()               # ascending: the cut is a prefix
    n = inside.sum(axis=1)
    P = np.where(inside[:, :, None], coords[np.maximum(idx, 0)], 0.0)         # (m, max_nn, 3); zeros add nothing
    s1 = np.cumsum(P, axis=1)[:, -1, :]
    prod = np.stack([P[:, :, 0] * P[:, :, 0], P[:, :, 0] * P[:, :, 1], P[:, :, 0] * P[:, :, 2],
                     P[:, :, 1] * P[:, :, 1], P[:, :, 1] * P[:, :, 2], P[:, :, 2] * P[:, :, 2]], axis=2)
    s2 = np.cumsum(prod, axis=1)[:, -1, :]
    ok = n >= 3
    assert np.array_equal(normals[~ok], np.tile([0.0, 0.0, 1.0], ((~ok).sum(), 1)))
    nn = n[ok].astype(np.float64)[:, None]
    mean, sq = s1[ok] / nn, s2[ok] / nn
    C = np.empty((ok.sum(), 3, 3))
    C[:, 0, 0] = sq[:, 0] - mean[:, 0] * mean[:, 0]; C[:, 0, 1] = C[:, 1, 0] = sq[:, 1] - mean[:, 0] * mean[:, 1]
    C[:, 0, 2] = C[:, 2, 0] = sq[:, 2] - mean[:, 0] * mean[:, 2]; C[:, 1, 1] = sq[:, 3] - mean[:, 1] * mean[:, 1]
    C[:, 1, 2] = C[:, 2, 1] = sq[:, 4] - mean[:, 1] * mean[:, 2]; C[:, 2, 2] = sq[:, 5] - mean[:, 2] * mean[:, 2]
    w, v = np.linalg.eigh(C)
    scale = np.abs(w).max(axis=1)
    well = (w[:, 1] - w[:, 0]) > 1e-6 * scale                                # direction determined to ~ 1e-9
    cosv = np.abs(np.einsum('ij,ij->i', v[:, :, 0], normals[ok]))
    worst = float((1.0 - cosv[well]).max())
    print(f'normals vs LAPACK eigh on the re-derived single-pass covariance: {well.sum()} of {ok.sum()} points with a '
          f'separated smallest eigenvalue, worst 1 - |cos| = {worst:.2e}; {(~well).sum()} near-degenerate skipped')
    assert well.sum() > 0.9 * ok.sum() and worst < 1e-9


def test_knn_sorted_space_equals_compact_space():
    """upcp_knn_normals_sorted: the same rows, numbered by the library's spatial order."""
    tile, pts, labels = _rg_cloud(120000, 43)
    sel = labels != 9
    t = dev.DeviceTile(pts)
    d_sel = dev.upload_mask(sel, t.n, t.device)
    _, bbox_sel, m = t.bbox(d_sel)
    d_knn, _, d_nrm, d_cov, _, d_flag = rgmod.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, threshold_curve=1.0)
    s_knn, s_nrm, s_cov, s_flag, s_order = rgmod.knn_normals_sorted_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, 1.0)
    order = s_order.cpu().numpy().astype(np.int64)               # original index of sorted position s
    ids = np.where(sel)[0]
    assert np.array_equal(np.sort(order), ids)                   # a permutation of the selected points
    compact_of_sorted = np.searchsorted(ids, order)              # compact id of sorted position s
    assert np.array_equal(compact_of_sorted[s_knn.cpu().numpy()], d_knn.cpu().numpy()[compact_of_sorted])
    assert np.array_equal(s_nrm.cpu().numpy(), d_nrm.cpu().numpy()[compact_of_sorted])
    assert np.array_equal(s_flag.cpu().numpy(), d_flag.cpu().numpy()[compact_of_sorted])
    # sorted space writes the covariance only where the seed flag is undecided (2): a threshold inside
    # the range of the eigenvalue ratios leaves many points undecided
    d_flag3 = rgmod.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, threshold_curve=0.3)[5]
    s_cov3, s_flag3, s_order3 = rgmod.knn_normals_sorted_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, 0.3)[2:5]
    # (the order inside a grid cell is the arrival order of a counting sort: every call has its own)
    cos3 = np.searchsorted(ids, s_order3.cpu().numpy().astype(np.int64))
    und = s_flag3.cpu().numpy() == 2
    assert np.array_equal(s_flag3.cpu().numpy(), d_flag3.cpu().numpy()[cos3]) and und.sum() > 1000
    assert np.array_equal(s_cov3.cpu().numpy()[und], d_cov.cpu().numpy()[cos3][und])
    # mask_take / mask_put round trip
    full = np.zeros(t.n, dtype=bool); full[ids[::3]] = True
    d_full = dev.upload_mask(full, t.n, t.device)
    taken = rgmod.mask_take(d_full, s_order)
    assert np.array_equal(taken.cpu().numpy().astype(bool), full[order])
    assert np.array_equal(dev.download_mask(rgmod.mask_put(taken, s_order, t.n)), full)


def test_knn_small_and_degenerate_clouds():
    rng = np.random.default_rng(7)
    for pts in (rng.random((7, 3)), np.round(rng.random((3000, 3)) * 2, 1), np.repeat(rng.random((1, 3)), 40, axis=0),
                np.column_stack((np.arange(200) * 0.01, np.zeros(200), np.zeros(200)))):
        pts = np.ascontiguousarray(pts + np.array([119300.0, 485100.0, 0.0]))
        t = dev.DeviceTile(pts)
        _, bbox, m = t.bbox(None)
        d_knn, d_d2, d_nrm, d_cov, _ = rgmod.knn_normals_device(t, None, m, bbox, 15, 30, 0.2, want_d2=True)
        want_idx, want_d2 = natives.knn(pts, 15, want_d2=True)
        assert np.array_equal(d_knn.cpu().numpy(), want_idx)
        assert np.array_equal(d_d2.cpu().numpy(), want_d2)
        np.testing.assert_allclose(d_nrm.cpu().numpy(), natives.normals_hybrid(pts, 0.2, 30), rtol=0, atol=1e-9)


def test_knn_extreme_extents_stay_exact():
    """Clouds that force the dense grid to coarsen (few points over kilometres), degenerate axes
    (a line along y: one cell per row) and extents where the fp32 filter must hand over to the
    fp64 tier: every tier is exact, so the rows still equal the oracle's."""
    rng = np.random.default_rng(8)
    clusters = rng.random((40, 3)) * np.array([5000.0, 5000.0, 300.0])
    wide = (clusters[rng.integers(0, 40, 20000)] + rng.normal(0, 0.05, (20000, 3)))
    line_y = np.column_stack((np.zeros(5000), np.sort(rng.random(5000)) * 80.0, np.zeros(5000)))
    plane = np.column_stack((rng.random(30000) * 3.0, rng.random(30000) * 3.0, np.zeros(30000)))
    far = rng.random((20000, 3)) * np.array([2.0, 2.0, 0.5]) + np.array([9.0e6, 4.0e6, 0.0])      # 1e-9-m ulp region
    mixed = np.vstack((rng.random((20000, 3)) * 0.5, rng.random((300, 3)) * 400.0))               # dense blob + far sparse
    for name, pts in (('wide', wide), ('line_y', line_y), ('plane', plane), ('far', far), ('mixed', mixed)):
        pts = np.ascontiguousarray(np.round(pts + np.array([119300.0, 485100.0, 0.0]) * (name != 'far'), 3))
        t = dev.DeviceTile(pts)
        _, bbox, m = t.bbox(None)
        for cell in (None, 0.02):
            d_knn, d_d2, d_nrm, d_cov, _ = rgmod.knn_normals_device(t, None, m, bbox, 15, 30, 0.2, cell_size=cell,
                                                                    want_d2=True)
            want_idx, want_d2 = natives.knn(pts, 15, want_d2=True)
            assert np.array_equal(d_knn.cpu().numpy(), want_idx), (name, cell)
            assert np.array_equal(d_d2.cpu().numpy(), want_d2), (name, cell)
            assert np.array_equal(d_cov.cpu().numpy(), natives.cov_knn(pts, want_idx)), (name, cell)
            np.testing.assert_allclose(d_nrm.cpu().numpy(), natives.normals_hybrid(pts, 0.2, 30), rtol=0, atol=1e-9)


@pytest.mark.parametrize('n,seed', [(60000, 51), (400000, 52)])
def test_region_growing_matches_oracle(n, seed):
    tile, pts, labels = _rg_cloud(n, seed)
    want, near = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), angle_tol=1e-6, return_near=True)
    rg = RegionGrowing(10, exclude_labels=(9,))
    got = rg.get_labels(pts, labels.copy(), None, tile['tilecode'])
    assert near.sum() == 0, f'{near.sum()} angle tests within 1e-6 deg of the threshold (excluded points)'
    assert np.array_equal(got, want)
    assert np.count_nonzero(want == 10) > np.count_nonzero(labels == 10)
    assert rg.last_stats['n_curvature_on_host'] < 0.01 * rg.last_stats['n_selected']
    # idempotence: growing again from the grown region changes nothing
    again = RegionGrowing(10, exclude_labels=(9,)).get_labels(pts, got.copy(), None, tile['tilecode'])
    assert np.array_equal(again, got)


@pytest.mark.parametrize('radius', [0.2, 0.1, 0.5])
def test_region_growing_radius_method_matches_oracle(radius):
    """RegionGrowing._region_growing('radius') (region_growing.py:104-106: growth and curvature test over everything
    within grow_region_radius) against the oracle's literal loop -- which equals the REAL reference class run over
    the restated natives (checked when the fixture logic was written, oracle/ref_harness.search_radius_vector_3d)."""
    tile, pts, labels = _rg_cloud(250_000 if radius < 0.5 else 120_000, 57)
    labels[(labels == 10) & (np.random.default_rng(5).random(len(labels)) < 0.9)] = 0      # few seeds: the region has to GROW
    want = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), grow_region_radius=radius, method='radius')
    rg = RegionGrowing(10, exclude_labels=(9,), grow_region_radius=radius)
    rg._set_mask(labels)
    rg._convert_input_cloud(pts)
    mask = rg._region_growing('radius')
    got = labels.copy()
    got[mask] = 10
    print(f'radius {radius}: region {int((want == 10).sum())} (seeds {int((labels == 10).sum())}), {int((got != want).sum())} labels differ, '
          f'{rg.last_stats["n_curvature_on_host"]} curvature flags resolved on the host in {rg.last_stats["rounds"]} rounds')
    assert np.array_equal(got, want) and (want == 10).sum() > (labels == 10).sum() + 1000
    # a curvature threshold inside the range of the ratios: many more points go through the exact host path
    if radius == 0.2:
        want = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), threshold_curve=0.45, method='radius')
        rg = RegionGrowing(10, exclude_labels=(9,), threshold_curve=0.45)
        rg._set_mask(labels); rg._convert_input_cloud(pts)
        got = labels.copy()
        got[rg._region_growing('radius')] = 10
        print(f'threshold_curve 0.45: {int((got != want).sum())} labels differ, {rg.last_stats["n_curvature_on_host"]} flags on the host')
        assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        RegionGrowing(10, max_nn=50)


def test_region_growing_parameters_and_edge_cases():
    tile, pts, labels = _rg_cloud(50000, 53)
    for kw in (dict(threshold_angle=5), dict(threshold_angle=60, grow_region_knn=8, max_nn=12, grow_region_radius=0.5),
               dict(threshold_curve=0.2)):
        want = R.region_growing_labels(pts, labels.copy(), 10, exclude_labels=(9,), **kw)
        got = RegionGrowing(10, exclude_labels=(9,), **kw).get_labels(pts, labels.copy(), None, tile['tilecode'])
        assert np.array_equal(got, want), kw
    # no seeds -> unchanged; antiparallel normals are NOT merged (vector_angle has no abs)
    l0 = np.where(labels == 10, 0, labels)
    assert np.array_equal(RegionGrowing(10, exclude_labels=(9,)).get_labels(pts, l0.copy(), None, ''), l0)


@pytest.mark.gpu
def test_device_generated_tiles_config5():
    """the bench tiles are generated ON THE DEVICE (upcp_synth_tile) from seed 20240000 + i.
    The generator is deterministic, tiles differ, the points lie on the LAS millimetre grid inside their tile --
    and for sampled tiles the labels of the device pipeline equal the oracle's on the downloaded copy."""
    import torch
    from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from urban_pointcloud_processing_b200.labels import Labels
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    from urban_pointcloud_processing_b200.region_growing import LabelConnectedComp, LayerLCC
    from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
    n = 150000
    layers = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5}, {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
    seen = []
    for i in (0, 1, 7):
        seed, origin = 20240000 + i, synthetic.tile_origin(i)
        tile, d_kind, layout = synthetic.make_device_tile(n, seed=seed, origin=origin)
        again, d_kind2, _ = synthetic.make_device_tile(n, seed=seed, origin=origin)
        pts = torch.stack((tile.x, tile.y, tile.z), dim=1).cpu().numpy()
        assert np.array_equal(pts, torch.stack((again.x, again.y, again.z), dim=1).cpu().numpy())
        assert torch.equal(d_kind, d_kind2)
        assert all(not np.array_equal(pts, p) for p in seen)
        seen.append(pts)
        kind = d_kind.cpu().numpy()
        frac = np.bincount(kind, minlength=6) / n
        cum = np.asarray(layout['cum_kind'], dtype=float)
        assert np.allclose(frac, np.diff(np.concatenate(([0.0], cum))), atol=0.01)
        assert np.all(pts[:, 0] >= origin[0] - 1e-9) and np.all(pts[:, 0] <= origin[0] + synthetic.TILE_SIZE + 1e-9)
        assert np.all(pts[:, 1] >= origin[1] - 1e-9) and np.all(pts[:, 1] <= origin[1] + synthetic.TILE_SIZE + 1e-9)
        assert np.allclose(pts * 1000.0, np.round(pts * 1000.0), atol=1e-4)          # millimetre grid
        # the device pipeline on the device-resident tile against the oracle on the downloaded copy
        tc, a = layout['tilecode'], layout['ahn_tile']
        ahn = ahn_utils.ArrayAHNReader({tc: a})
        bld = bgt_utils.ArrayBGTReader({tc: layout['building_polygons']})
        road = bgt_utils.ArrayBGTReader({tc: layout['road_polygons']})
        procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
                 BGTRoadFuser(Labels.ROAD, bgt_reader=road),
                 BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
                 LayerLCC(Labels.BUILDING, ahn, params=[dict(p) for p in layers]),
                 LabelConnectedComp(Labels.CAR, grid_size=0.1, min_component_size=30, threshold=0.02),
                 RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)))
        labels0 = np.zeros(n, dtype=np.uint16)
        labels0[np.where(kind == 2)[0][::40]] = Labels.CAR
        d_labels = dev.upload_labels(labels0, tile.device)
        Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_tile_device(tc, tile, d_labels)
        got = d_labels.cpu().numpy()
        want = R.process_cloud([
            lambda p, l, m: R.ahn_fuser_labels(p, l, m, a, Labels.GROUND, 'ground', 0.2),
            lambda p, l, m: R.road_fuser_labels(p, l, layout['road_polygons'], Labels.ROAD),
            lambda p, l, m: R.building_fuser_labels(p, l, m, layout['building_polygons'], Labels.BUILDING, a, 0.2),
            lambda p, l, m: R.layer_lcc_labels(p, l, m, a, Labels.BUILDING, [dict(q) for q in layers]),
            lambda p, l, m: (l.__setitem__(R.lcc_get_label_mask(p, l, m, Labels.CAR, grid_size=0.1,
                                                                min_component_size=30, threshold=0.02), Labels.CAR), l)[1],
            lambda p, l, m: R.region_growing_labels(p, l, Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)),
        ], np.asfortranarray(pts), labels0.copy())
        n_diff = int(np.count_nonzero(got != want.astype(np.int32)))
        print(f'device tile {i}: {n} points, kinds {np.round(frac, 3).tolist()}, {n_diff} labels differ from the oracle')
        assert n_diff == 0


@pytest.mark.gpu
def test_notebook0_processor_sequence_with_its_defaults():
    """The in-scope part of the reference's "0. Complete solution" notebook, composed as there and with its arguments:
    AHNFuser(refine_ground=True) -> BGTRoadFuser -> NoiseFilter -> BGTBuildingFuser(offset=0.25) -> CarFuser ->
    LayerLCC(building top / bottom) through Pipeline.process_cloud, against the oracle's composition of the same steps
    (buffers as distances; the cluster size limits are scaled to the test cloud's density)."""
    from urban_pointcloud_processing_b200.fusion import CarFuser, NoiseFilter
    tile = synthetic.make_tile(500_000, seed=2024, n_cars=12)
    pts, tc, a = tile['points'], tile['tilecode'], tile['ahn_tile']
    ahn, bld, road = readers_for(tile)
    car_params = {'min_height': 1.2, 'max_height': 2.4, 'min_width': 1.4, 'max_width': 2.4, 'min_length': 3.0, 'max_length': 6.0}
    layers = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5}, {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=True),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             NoiseFilter(Labels.NOISE, ahn_reader=ahn, epsilon=0.2, min_component_size=20),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0.25, ahn_reader=ahn),
             CarFuser(Labels.CAR, ahn, bgt_reader=road, exclude_types=['fietspad'], grid_size=0.05, min_component_size=150,
                      params=car_params),
             LayerLCC(Labels.BUILDING, ahn, params=[dict(p) for p in layers]))
    labels = np.zeros(len(pts), dtype=np.uint16)
    got = Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud(tc, pts, labels.copy())
    roads = road.filter_tile(tc, exclude_types=['fietspad'], merge=False)
    want = R.process_cloud([
        lambda p, l, m: R.ahn_fuser_labels(p, l, m, a, Labels.GROUND, 'ground', 0.2, refine_ground=True, params={}),
        lambda p, l, m: R.road_fuser_labels(p, l, tile['road_polygons'], Labels.ROAD),
        lambda p, l, m: R.noise_filter_labels(p, l, m, a, Labels.NOISE, epsilon=0.2, grid_size=0.1, min_component_size=20),
        lambda p, l, m: R.building_fuser_labels(p, l, m, [q.buffer(0.25) for q in tile['building_polygons']], Labels.BUILDING, a, 0.2),
        lambda p, l, m: R.car_fuser_labels(p, l, m, a, roads, Labels.CAR, grid_size=0.05, min_component_size=150,
                                           overlap_perc=20, params=car_params),
        lambda p, l, m: R.layer_lcc_labels(p, l, m, a, Labels.BUILDING, [dict(q) for q in layers]),
    ], pts, labels.copy())
    hist = dict(zip(*[v.tolist() for v in np.unique(got, return_counts=True)]))
    print(f'notebook-0 sequence: {len(pts)} points, labels {hist}, {int((got != want).sum())} differ from the oracle')
    assert np.array_equal(got, want)
    assert all(hist.get(int(k), 0) > 0 for k in (Labels.GROUND, Labels.ROAD, Labels.BUILDING, Labels.NOISE))
