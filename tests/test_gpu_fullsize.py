"""GPU: the BASELINE.json configurations at their FULL sizes (20 M / 50 M points).

Where the oracle still finishes in seconds (raster lookup, octree-cell connected components)
the full-size result is compared with it directly, bit for bit.  Where it does not
(point-in-polygon against 5 000 rings, region growing on 50 M points) the full-size result is
checked on a random sample against the oracle and through size-independent properties of the
domain: independence of the acceleration structure, idempotence of the fixpoint, closure and
support of the grown region re-derived from the device's own kNN graph with plain torch.
"""
import numpy as np
import pytest
import torch

from oracle import reference_py as R
from test_gpu_parity import readers_for
from urban_pointcloud_processing_b200 import _lib, synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
from urban_pointcloud_processing_b200.labels import Labels
from urban_pointcloud_processing_b200.pipeline import Pipeline
from urban_pointcloud_processing_b200.region_growing import LabelConnectedComp, LayerLCC
from urban_pointcloud_processing_b200.region_growing import region_growing as rgmod
from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing

pytestmark = pytest.mark.gpu

N20 = 20_000_000
N50 = 50_000_000


@pytest.fixture(scope='module')
def tile20():
    return synthetic.make_tile(N20, seed=20240000)


# ------------------------------------------------ config 2: LabelConnectedComp, 20 M ---
def test_config2_lcc_20m_matches_oracle(tile20):
    pts = tile20['points']
    mask = tile20['kind'] != 0                                   # non-ground, ~11 M points
    sub = pts[mask]
    # grid 0.1 (level 9): direct comparison with the oracle at full size
    lcc = LabelConnectedComp(Labels.CAR, grid_size=0.1, min_component_size=1500, threshold=0.1)
    got = lcc.get_components(sub)
    want = R.lcc_get_components(sub, 0.1, 1500)
    assert lcc.octree_level == R.get_octree_level(sub, 0.1) == 9
    assert np.array_equal(R.canonical_components(got), R.canonical_components(want))
    assert (want >= 0).sum() > 5_000_000 and (want < 0).sum() > 100_000
    # min_component_size only removes components: the partitions of the kept points nest
    got_100 = lcc_components(sub, 0.1, 100)
    kept = got >= 0
    assert (got_100[kept] >= 0).all()
    assert np.array_equal(R.canonical_components(got_100[kept]), R.canonical_components(got[kept]))
    # get_label_mask with 1 % of the car points as seeds: membership follows from the components
    labels = np.zeros(N20, dtype=np.uint16)
    labels[tile20['kind'] == 0] = Labels.GROUND
    cars = np.where(tile20['kind'] == 2)[0]
    labels[cars[::100]] = Labels.CAR
    sel = labels != Labels.GROUND
    lm = LabelConnectedComp(Labels.CAR, grid_size=0.1, min_component_size=100, threshold=0.005).get_label_mask(
        pts, labels.copy(), labels == 0)
    comp_sel = got_100                                           # same selection (non-ground), same grid
    assert sel.sum() == len(comp_sel)
    seeds = (labels[sel] == Labels.CAR)
    ids = comp_sel.astype(np.int64)
    ok = ids >= 0
    size = np.bincount(ids[ok]); seed_cnt = np.bincount(ids[ok], weights=seeds[ok].astype(np.float64), minlength=len(size))
    fill = np.zeros(len(size), dtype=bool)
    nz = size > 0
    fill[nz] = seed_cnt[nz] / size[nz] > 0.005                  # label_connected_comp.py:117-121
    want_mask = np.zeros(N20, dtype=bool)
    want_mask[np.where(sel)[0][ok]] = fill[ids[ok]]
    assert np.array_equal(lm, want_mask) and lm.sum() > 10_000


def lcc_components(points, grid_size, min_size):
    return LabelConnectedComp(grid_size=grid_size, min_component_size=min_size).get_components(points)


def test_config2_lcc_20m_level10_matches_oracle(tile20):
    """grid 0.05 (level 10) and min_component_size in {100, 1500, 5000} (BASELINE config 2) against the
    oracle DIRECTLY at full size, plus grid 0.1 with 5000."""
    sub = tile20['points'][tile20['kind'] != 0]
    assert R.get_octree_level(sub, 0.05) == 10
    want_all = R.lcc_get_components(sub, 0.05, 1)               # the full partition, once (20 s of CPU)
    sizes = np.bincount(want_all.astype(np.int64))
    for min_size in (100, 1500, 5000):
        got = lcc_components(sub, 0.05, min_size)
        want = np.where(sizes[want_all.astype(np.int64)] >= min_size, want_all, -1)      # label_connected_comp.py:89-97
        assert np.array_equal(R.canonical_components(got), R.canonical_components(want)), min_size
        print(f'level 10, min {min_size}: {len(np.unique(got[got >= 0]))} components, {(got < 0).sum()} points dropped')
    # the oracle's own filter for the largest size, and the coarser grid with it
    assert np.array_equal(R.canonical_components(lcc_components(sub, 0.05, 5000)),
                          R.canonical_components(R.lcc_get_components(sub, 0.05, 5000)))
    assert np.array_equal(R.canonical_components(lcc_components(sub, 0.1, 5000)),
                          R.canonical_components(R.lcc_get_components(sub, 0.1, 5000)))
    # how exposed the float32 cell arithmetic is on this cloud: points whose level-10 cell coordinate
    # lies within one float32 ulp of a cell boundary (a differently rounded division would move them)
    from oracle import natives
    xs, ys, zs = (np.ascontiguousarray(sub[:, i].astype(np.float32)) for i in range(3))
    p = natives.octree_params(xs, ys, zs)
    dim_min, cs = np.asarray(p[:3], dtype=np.float32), np.float32(p[3])
    width = np.float32(cs * np.float32(1 << (21 - 10)))
    near = np.zeros(len(sub), dtype=bool)
    for i, c in enumerate((xs, ys, zs)):
        t = (c - dim_min[i]) / width                             # float32, like the cell position itself
        near |= (np.nextafter(t, np.float32(np.inf)).astype(np.int64) != t.astype(np.int64)) | \
                (np.nextafter(t, np.float32(-np.inf)).astype(np.int64) != t.astype(np.int64))
    print(f'level 10: {near.sum()} of {len(sub)} points within 1 float32 ulp of a cell boundary')


def test_config2_lcc_20m_level10_invariants(tile20):
    """grid 0.05 (level 10): cell-graph invariants re-derived in numpy from the oracle's own cell
    assignment + permutation invariance."""
    sub = tile20['points'][tile20['kind'] != 0]
    got = lcc_components(sub, 0.05, 100)
    level = R.get_octree_level(sub, 0.05)
    assert level == 10
    from oracle import natives
    xs, ys, zs = (np.ascontiguousarray(sub[:, i].astype(np.float32)) for i in range(3))
    p = natives.octree_params(xs, ys, zs)
    dim_min, cs = np.asarray(p[:3], dtype=np.float32), np.float32(p[3])
    shift = 21 - level
    cell = [np.clip(((c - dim_min[i]) / cs).astype(np.int64), 0, (1 << 21) - 1) >> shift
            for i, c in enumerate((xs, ys, zs))]
    key = (cell[2] << 42) | (cell[1] << 21) | cell[0]
    # (1) all points of one cell share one unfiltered component: sizes per cell sum up consistently
    full = lcc_components(sub, 0.05, 1)
    order = np.argsort(key, kind='stable')
    ks, cf = key[order], full[order]
    first = np.r_[True, ks[1:] != ks[:-1]]
    cell_comp = cf[first]
    assert np.array_equal(cf, np.repeat(cell_comp, np.diff(np.r_[np.where(first)[0], len(ks)])))
    # (2) 26-adjacent occupied cells share a component (sampled cells, all 13 forward offsets)
    ukeys = ks[first]
    rng = np.random.default_rng(11)
    samp = rng.choice(len(ukeys), size=300_000, replace=False)
    ux, uy, uz = ukeys & ((1 << 21) - 1), (ukeys >> 21) & ((1 << 21) - 1), ukeys >> 42
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if (dz, dy, dx) <= (0, 0, 0):
                    continue
                nk = ((uz[samp] + dz) << 42) | ((uy[samp] + dy) << 21) | (ux[samp] + dx)
                pos = np.searchsorted(ukeys, nk)
                hit = (pos < len(ukeys)) & (ukeys[np.minimum(pos, len(ukeys) - 1)] == nk) & (ux[samp] + dx >= 0)
                assert np.array_equal(cell_comp[samp][hit], cell_comp[pos[hit]])
    # (3) the size filter is exactly "component size < min"
    sizes = np.bincount(full.astype(np.int64))
    assert np.array_equal(got >= 0, sizes[full.astype(np.int64)] >= 100)
    # (4) permutation invariance
    perm = rng.permutation(len(sub))
    b = lcc_components(sub[perm], 0.05, 100)
    b_orig = np.empty_like(b); b_orig[perm] = b
    assert np.array_equal(R.canonical_components(got), R.canonical_components(b_orig))


# ------------------------------------- config 4: PIP vs 5 000 rings + AHN raster, 20 M ---
def test_config4_pip_and_raster_20m(tile20):
    pts, ahn = tile20['points'], tile20['ahn_tile']
    t = dev.DeviceTile(pts)
    # raster lookup + threshold: full-size comparison with numpy (FastGridInterpolator semantics)
    for surface, mode, op in (('ground_surface', _lib.CLS_ABS_LT, lambda z, v: np.abs(z - v) < 0.2),
                              ('building_surface', _lib.CLS_CAP_LE, lambda z, v: ~np.isfinite(v) | (z <= v + 0.2))):
        raster = dev.DeviceRaster(ahn['x'], ahn['y'], ahn[surface])
        v = R.fast_grid_lookup(ahn['x'], ahn['y'], ahn[surface], pts)
        assert np.array_equal(raster.lookup(t).cpu().numpy(), v, equal_nan=True)
        with np.errstate(invalid='ignore'):
            want = op(pts[:, 2], v)
        assert np.array_equal(dev.download_mask(raster.classify(t, mode, a=0.2)), want)
    # point-in-polygon against the 5 000-ring lattice
    polys = synthetic.make_polygon_lattice()
    assert len(polys) == 5041
    got = dev.download_mask(dev.PipPlan(polys).query(t))
    # (1) the acceleration grid does not change a single bit
    assert np.array_equal(got, dev.download_mask(dev.PipPlan(polys, grid_dim=97).query(t)))
    # (2) a random sample of the full-size run against the oracle
    rng = np.random.default_rng(12)
    samp = np.sort(rng.choice(N20, size=150_000, replace=False))
    sub = np.ascontiguousarray(pts[samp])
    acc = np.zeros(len(sub), dtype=bool)
    for p in polys:
        acc |= R.poly_clip(sub, p)
    assert np.array_equal(got[samp], acc)
    # (3) coverage is what the lattice geometry implies (star-shaped rings of radius 0.25-0.45 m on a 0.7 m pitch)
    assert 0.15 < got.mean() < 0.6
    # (4) selection restricts, never adds
    sel = rng.random(N20) < 0.5
    got_sel = dev.download_mask(dev.PipPlan(polys).query(t, dev.upload_mask(sel, t.n, t.device)))
    assert np.array_equal(got_sel, got & sel)


# --------------------------------------------- config 3: RegionGrowing, 50 M, facades ---
def test_config3_region_growing_50m_fixpoint():
    tile = synthetic.make_tile(N50, seed=20240003, facade_fraction=0.7)
    pts = tile['points']
    ahn, bld, road = readers_for(tile)
    tc = tile['tilecode']
    labels = np.zeros(N50, dtype=np.uint16)
    AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False).get_labels(pts, labels, labels == 0, tc)
    BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn).get_labels(pts, labels, labels == 0, tc)
    before = labels.copy()
    rg = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,))
    out = rg.get_labels(pts, labels, None, tc)
    assert out is labels
    grown = (labels == Labels.BUILDING)
    seeds0 = (before == Labels.BUILDING)
    # only unlabelled... no: RegionGrowing relabels every non-excluded point it reaches; seeds stay, ground never changes
    assert (grown >= seeds0).all() and np.array_equal(labels == Labels.GROUND, before == Labels.GROUND)
    assert grown.sum() > seeds0.sum() + 100_000
    assert np.array_equal(labels[~grown], before[~grown])
    # growing again from the grown region: monotone, and it can only add what the region's NON-seed
    # points (curvature test failed, region_growing.py:132) reach once they are initial seeds
    again = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,)).get_labels(pts, labels.copy(), None, tc)
    regrown = (again == Labels.BUILDING)
    assert (regrown >= grown).all() and np.array_equal(again[~regrown], labels[~regrown])
    n_regrown = int(regrown.sum() - grown.sum())
    # closure and support, re-derived with torch from the device's kNN graph (compact index space)
    t = dev.get_tile(pts)
    sel = before != Labels.GROUND
    d_sel = dev.upload_mask(sel, t.n, t.device)
    _, bbox_sel, m = t.bbox(d_sel)
    d_knn, _, d_nrm, d_cov, _, d_flag = rgmod.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, want_curv=False,
                                                                 threshold_curve=1.0)
    d_can, _ = rgmod.resolve_seed_flags(d_flag, d_cov, 1.0)
    region = torch.from_numpy(grown[sel]).to(t.device)
    s0 = torch.from_numpy(seeds0[sel]).to(t.device)
    is_seed = s0 | (region & (d_can == 1))
    supported = s0.clone()
    n_near = 0
    chunk = 4_000_000
    for lo in range(0, m, chunk):
        hi = min(lo + chunk, m)
        nb = d_knn[lo:hi, 1:].long()                                   # neighbour_idx[1:k] (region_growing.py:112)
        ns = d_nrm[lo:hi].unsqueeze(1)
        nn = d_nrm[nb]
        cosang = (ns * nn).sum(-1) / (ns.norm(dim=-1) * nn.norm(dim=-1))
        ang = torch.rad2deg(torch.acos(cosang.clamp(-1, 1)))
        near = (ang - 20.0).abs() < 1e-6
        n_near += int(near.sum())
        accept = (ang < 20.0) & ~near & is_seed[lo:hi].unsqueeze(1)
        # closure: everything a seed accepts is in the region
        assert bool(region[nb[accept]].all())
        supported[nb[accept]] = True
    # support: every region point is an initial seed or accepted by some seed of the region
    unsupported = region & ~supported
    assert int(unsupported.sum()) <= n_near, (int(unsupported.sum()), n_near)
    n_nonseed = int((region & ~is_seed).sum())
    assert n_regrown == 0 or n_nonseed > 0
    print(f'config 3: m = {m}, region = {int(region.sum())}, non-seed region points = {n_nonseed}, '
          f'added by a second pass = {n_regrown}, angle tests within 1e-6 deg of the threshold: {n_near}')


def test_config3_knn_rows_and_normals_50m_match_oracle_on_a_crop():
    """Config 3 at full size, checked against the ORACLE rather than against the device's own graph: the kNN rows,
    squared distances and hybrid normals the device computed for the 50 M-point tile (facade-heavy, generated on
    the device) are compared with natives.knn / natives.normals_hybrid on a ~2.5 M-point spatial crop, for the
    interior points of the crop (whose neighbourhoods lie inside it), for grow_region_radius 0.1 / 0.2 / 0.5; then
    RegionGrowing runs for the two other radii of the configuration.  The exposure to the one ambiguity of
    the restated neighbour order -- exact distance ties -- is counted on the same crop."""
    from oracle import natives
    from test_gpu_parity import report_tie_exposure
    tile, kind, layout = synthetic.make_device_tile(N50, seed=20240003, origin=synthetic.tile_origin(3), facade_fraction=0.7)
    del kind
    tc = layout['tilecode']
    ahn = __import__('urban_pointcloud_processing_b200.utils.ahn_utils', fromlist=['x']).ArrayAHNReader({tc: layout['ahn_tile']})
    bld = __import__('urban_pointcloud_processing_b200.utils.bgt_utils', fromlist=['x']).ArrayBGTReader({tc: layout['building_polygons']})
    d_labels = torch.zeros(N50, dtype=torch.int32, device=tile.device)
    Pipeline(processors=(AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
                         BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn)),
             ahn_reader=ahn, caching=True).process_tile_device(tc, tile, d_labels)
    d_sel = (d_labels != Labels.GROUND).to(torch.uint8)
    _, bbox_sel, m = tile.bbox(d_sel)
    # crop: 11 m square around a building corner (facade, balcony clutter and street level all occur)
    cx, cy = layout['boxes'][0][2], layout['boxes'][0][3]
    side, margin = 11.0, 0.75
    ox, oy = layout['origin']
    cx = min(max(cx, ox + side / 2), ox + 50.0 - side / 2); cy = min(max(cy, oy + side / 2), oy + 50.0 - side / 2)
    d_in = (d_sel != 0) & ((tile.x - cx).abs() <= side / 2) & ((tile.y - cy).abs() <= side / 2)
    d_ids = torch.nonzero(d_in).flatten()
    compact = torch.cumsum(d_sel.to(torch.int32), 0, dtype=torch.int32) - 1           # compact id of every selected point
    d_cid = compact[d_ids].long()
    crop_of_compact = torch.full((m,), -1, dtype=torch.int32, device=tile.device)
    crop_of_compact[d_cid] = torch.arange(len(d_cid), dtype=torch.int32, device=tile.device)
    coords = torch.stack((tile.x[d_ids], tile.y[d_ids], tile.z[d_ids]), dim=1).cpu().numpy()
    assert len(coords) > 1_500_000
    want_idx, want_d2 = natives.knn(coords, 15, want_d2=True)
    interior = (np.abs(coords[:, 0] - cx) <= side / 2 - margin) & (np.abs(coords[:, 1] - cy) <= side / 2 - margin) & \
        (want_d2[:, 14] < 0.7 ** 2)
    d_int = torch.from_numpy(np.where(interior)[0]).to(tile.device)
    report_tie_exposure(coords)
    for radius in (0.2, 0.1, 0.5):
        d_knn, d_d2, d_nrm, _, _ = rgmod.knn_normals_device(tile, d_sel, m, bbox_sel, 15, 30, radius, want_d2=True,
                                                          want_cov=False, want_curv=False)
        rows = crop_of_compact[d_knn[d_cid[d_int]].long()].cpu().numpy()                  # neighbours as crop indices
        assert np.array_equal(rows, want_idx[interior]), radius
        assert np.array_equal(d_d2[d_cid[d_int]].cpu().numpy(), want_d2[interior]), radius
        want_n = natives.normals_hybrid(coords, radius, 30)[interior]
        got_n = d_nrm[d_cid[d_int]].cpu().numpy()
        bad = np.abs(got_n - want_n).max(axis=1) > 1e-9
        print(f'config 3, radius {radius}: {int(interior.sum())} interior points of a {len(coords)}-point crop of the {m} selected: '
              f'kNN rows / d2 identical, normals bit-identical {np.mean((got_n == want_n).all(axis=1)):.4f}, beyond 1e-9: {int(bad.sum())}')
        assert bad.sum() == 0
        del d_knn, d_d2, d_nrm
    for radius in (0.1, 0.5):
        d_out = d_labels.clone()
        rg = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,), grow_region_radius=radius)
        rg._apply_device(tile, d_out, None, tc)
        grown, seeds0 = d_out == Labels.BUILDING, d_labels == Labels.BUILDING
        assert bool((grown | ~seeds0).all()) and bool(torch.equal(d_out == Labels.GROUND, d_labels == Labels.GROUND))
        assert int(grown.sum()) > int(seeds0.sum()) + 50_000 and bool(torch.equal(d_out[~grown], d_labels[~grown]))
        print(f'config 3, radius {radius}: region {int(grown.sum())} from {int(seeds0.sum())} seeds, '
              f'{rg.last_stats["n_curvature_on_host"]} curvature flags on the host')


# --------------------------------------------------- config 5: full pipeline, one tile ---
def test_config5_pipeline_20m_host_api_equals_device_arm(tile20):
    pts = np.asfortranarray(tile20['points'])
    ahn, bld, road = readers_for(tile20)
    tc = tile20['tilecode']

    def procs():
        return (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
                BGTRoadFuser(Labels.ROAD, bgt_reader=road),
                BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
                LayerLCC(Labels.BUILDING, ahn, params=[{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                                                       {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]),
                RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)))
    labels = np.zeros(N20, dtype=np.uint16)
    Pipeline(processors=procs(), ahn_reader=ahn, caching=True).process_cloud(tc, pts, labels)
    # stage by stage through get_labels gives the same labels (pipeline.py:88-93)
    stagewise = np.zeros(N20, dtype=np.uint16)
    for p in procs():
        stagewise = p.get_labels(pts, stagewise, stagewise == 0, tc)
    assert np.array_equal(labels, stagewise)
    hist = R.label_histogram(labels)
    assert hist.sum() == N20 and hist[Labels.GROUND] > 5_000_000 and hist[Labels.BUILDING] > 5_000_000
    d_hist = dev.label_hist(dev.upload_labels(labels, dev.current_device())).cpu().numpy()
    assert np.array_equal(d_hist, hist)


# ------------------------------- "next" row 2 at full size: object fusers on the 20 M tile ---
def test_object_fusers_20m_reductions_and_clip():
    """CarFuser / BGTStreetFurnitureFuser on the 20 M-point tile.  The oracle's per-cluster Python loops do not
    finish in seconds at this size, so: the per-cluster reductions are re-derived with numpy bincounts from
    the device's own cluster ids (clusters themselves are checked against the oracle in config 2), every
    accepted car is re-judged on the host, and the final mask must equal the OR of the oracle's
    poly_box_clip over the accepted rectangles."""
    from urban_pointcloud_processing_b200.fusion import BGTStreetFurnitureFuser, CarFuser
    from urban_pointcloud_processing_b200.region_growing.label_connected_comp import lcc_device
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils, math_utils, object_utils
    tile20 = synthetic.make_tile(N20, seed=20240001, n_cars=12)          # few cars: most stand alone
    pts, tc = tile20['points'], tile20['tilecode']
    ahn, _, road = readers_for(tile20)
    labels = np.zeros(N20, dtype=np.uint16)
    labels[tile20['kind'] == 0] = Labels.GROUND
    mask = labels == 0
    t = dev.get_tile(pts)
    d_mask = dev.upload_mask(mask, t.n, t.device)
    d_gz = ahn_utils.device_raster_for(ahn, tc, 'ground_surface').lookup(t, d_sel=d_mask)
    d_comp, _, _ = lcc_device(t, d_mask, 0.1, 1500, level_from='sel')
    st = object_utils.ComponentStats(t, d_comp, d_gz)
    comp, gz = d_comp.cpu().numpy(), d_gz.cpu().numpy()
    inc = comp >= 0
    ids, dense = np.unique(comp[inc], return_inverse=True)
    assert st.n_components == len(ids) > 10 and st.n_points == inc.sum() > 5_000_000
    assert np.array_equal(st.d_dense.cpu().numpy()[inc], dense)
    assert np.array_equal(st.size, np.bincount(dense))
    fin = np.isfinite(gz[inc])
    assert np.array_equal(st.count_finite, np.bincount(dense[fin], minlength=len(ids)))
    mean = np.bincount(dense[fin], weights=gz[inc][fin], minlength=len(ids)) / np.maximum(st.count_finite, 1)
    has = st.count_finite > 0
    assert np.abs(st.mean_gz[has] - mean[has]).max() < 1e-9 and np.isnan(st.mean_gz[~has]).all()
    order = np.argsort(dense, kind='stable')
    assert np.array_equal(st.d_order.cpu().numpy().view(np.uint32)[:inc.sum()], np.where(inc)[0][order])
    zmax = np.full(len(ids), -np.inf); np.maximum.at(zmax, dense, pts[inc, 2])
    assert np.array_equal(st.max_z, zmax)

    params = {'min_height': 1.2, 'max_height': 2.4, 'min_width': 1.4, 'max_width': 2.4, 'min_length': 3.0, 'max_length': 6.0}
    car = CarFuser(Labels.CAR, ahn, road, grid_size=0.1, min_component_size=1500, params=params, exclude_types=[])
    got = car.get_label_mask(pts, labels, mask, tc)
    assert len(car.last_objects) >= 1
    sub = pts[mask]
    want = np.zeros(len(sub), dtype=bool)
    for ring, bottom, top in car.last_objects:
        _, _, w, l, _ = math_utils.minimum_bounding_rectangle(ring[:4])
        assert 1.4 < w < 2.4 and 3.0 < l < 6.0
        assert max(R.rect_overlap_percentage(ring, p) for p in tile20['road_polygons']) > 20
        want |= R.poly_clip(sub, bgt_utils.RingPolygon(ring)) & ((sub[:, 2] <= top) & (sub[:, 2] >= bottom))
    full = np.zeros(N20, dtype=bool); full[mask] = want
    assert np.array_equal(got, full) and got.sum() > 10_000

    poles = bgt_utils.ArrayBGTPointReader({tc: [('lichtmast', x, y) for x, y in tile20['pole_xy']]})
    pole_params = {'min_height': 1.0, 'max_height': 12.0, 'min_width': 0.1, 'max_width': 6.0, 'min_length': 0.1, 'max_length': 6.0}
    sf = BGTStreetFurnitureFuser(Labels.STREET_LIGHT, 'lichtmast', poles, ahn, grid_size=0.1, min_component_size=1500,
                                 params=pole_params)
    lm = sf.get_label_mask(pts, labels, mask, tc)
    assert sf.last_accepted >= 1 and lm.sum() > 5_000 and not (lm & ~mask).any()
    # whole clusters are labelled: the mask is a union of clusters of the same clustering
    hit = np.bincount(dense, weights=lm[inc].astype(np.float64))
    assert np.array_equal((hit > 0) & (hit < st.size), np.zeros(len(ids), dtype=bool)) and lm[~inc].sum() == 0
