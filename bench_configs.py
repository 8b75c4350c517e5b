"""Per-configuration benchmark lines of BASELINE.md section 4 (`python bench.py --config 1|2|3|4`).

One JSON line per configuration: the workload of BASELINE.json `configs[i]` at its FULL size on one B200,
device-resident (`value`) and through the public host API with pinned numpy arrays (`e2e`), the algorithmic
HBM bytes of its stage against the measured copy bandwidth (`roofline`), and the oracle port
timed on the host cores on a bounded sample (`cpu_baseline`).  The default `bench.py` line (config 5: batches
of 20 M-point tiles through the full processor list) is the headline; these lines are its parts.
"""
import json
import os
import time

import numpy as np

import bench as B

GOLDEN_DEMO = os.path.join(B.ROOT, 'tests', 'golden', 'demo')


_SAMPLER = None          # nvidia-smi clocks / throttle reasons; the window is the first (device-resident) timed region


def _timed(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    wall = (t1 - t0) * 1e3
    if _SAMPLER is not None and getattr(_SAMPLER, 't0', None) is None:
        _SAMPLER.window(t0, t1)
    return max(e0.elapsed_time(e1), 0.0) / reps, wall / reps, out


def _pinned_points(torch, tile):
    pinned = torch.empty((3, tile.n), dtype=torch.float64).pin_memory()
    pinned[0].copy_(tile.x); pinned[1].copy_(tile.y); pinned[2].copy_(tile.z)
    torch.cuda.synchronize()
    return pinned, pinned.numpy().T


def _line(args, metric, workload, value, ms, extra, roofline, cpu, e2e, launches):
    peak, peak_src = B.peaks()
    if roofline is not None:
        roofline = dict(roofline, bound='hbm', peak=peak, unit='GB/s', frac=roofline['achieved'] / peak, peak_source=peak_src,
                        traffic=None, csrc_sha=B.sources_sha())
    line = {'metric': metric, 'value': value, 'unit': 'Mpts/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': dict({'workload': workload}, **extra), 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': roofline, 'cpu_baseline': cpu, 'clocks': _SAMPLER.stop() if _SAMPLER is not None else None}
    print(json.dumps(line))


def _cpu(value_pts, dt, sample, what, per=None):
    cores = os.cpu_count() or 1
    return {'value': value_pts / dt / 1e6, 'unit': 'Mpts/s', 'cores': cores, 'kind': 'port', 'sample': sample,
            'what': what, 'per_processor_s': per}


def config1(args, torch, dev, _lib):
    """4 M synthetic points over the REAL demo tile: AHN ground -> BGT road -> BGT building -> LayerLCC, then RegionGrowing."""
    import ast
    import pandas as pd
    from oracle import reference_py as R
    from urban_pointcloud_processing_b200 import synthetic
    from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from urban_pointcloud_processing_b200.labels import Labels
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    from urban_pointcloud_processing_b200.region_growing import LayerLCC
    from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
    n = 4_000_000
    ahn_tile = ahn_utils.load_ahn_tile(os.path.join(GOLDEN_DEMO, 'ahn_2386_9702.npz'))
    df = pd.read_csv(os.path.join(GOLDEN_DEMO, 'bgt_buildings_demo.csv'))
    rings = [np.array(ast.literal_eval(p), dtype=np.float64) for p in df.polygon.values]
    points, _ = synthetic.make_demo_cloud(n, ahn_tile, rings, seed=23869702)
    pinned = torch.from_numpy(np.ascontiguousarray(points.T)).pin_memory()
    pts_host = pinned.numpy().T
    ahn = ahn_utils.NPZReader(GOLDEN_DEMO)
    bld = bgt_utils.BGTPolyReader(bgt_file=os.path.join(GOLDEN_DEMO, 'bgt_buildings_demo.csv'))
    road = bgt_utils.BGTPolyReader(bgt_file=os.path.join(GOLDEN_DEMO, 'bgt_roads_demo.csv'))
    layers = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5}, {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
    procs = (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
             BGTRoadFuser(Labels.ROAD, bgt_reader=road),
             BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
             LayerLCC(Labels.BUILDING, ahn, params=[dict(p) for p in layers]))
    pipe = Pipeline(processors=procs, ahn_reader=ahn, caching=True)
    rg = RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND,))
    tile = dev.DeviceTile(pts_host)
    d_lab = torch.zeros(n, dtype=torch.int32, device=tile.device)

    def device_step():
        d_lab.zero_()
        pipe.process_tile_device('2386_9702', tile, d_lab)
        rg._apply_device(tile, d_lab, None, '2386_9702')
    l0 = _lib.launch_count()
    ms, wall, _ = _timed(torch, device_step, args.steps, warm=max(args.warmup, 3))
    launches = (_lib.launch_count() - l0) // (args.steps + max(args.warmup, 1))
    lab_pinned = torch.zeros(n, dtype=torch.int16).pin_memory()
    labels_host = lab_pinned.numpy().view(np.uint16)

    def host_step():
        dev.clear_tile_cache()
        labels_host[:] = 0
        pipe.process_cloud('2386_9702', pts_host, labels_host)
        rg.get_labels(pts_host, labels_host, labels_host == 0, '2386_9702')
    ems, ewall, _ = _timed(torch, host_step, args.steps)
    golden = np.load(os.path.join(B.ROOT, 'tests', 'golden', 'config1_golden.npz'))
    same = bool(np.array_equal(labels_host, golden['labels_region_growing']))
    e2e = {'value': n / (max(ems, ewall) * 1e-3) / 1e6, 'unit': 'Mpts/s', 'h2d_bytes_per_step': int(2 * (n * 24 + n * 2)),
           'd2h_bytes_per_step': int(2 * n * 2), 'labels_equal_reference_golden': same,
           'api': 'Pipeline.process_cloud + RegionGrowing.get_labels, numpy arrays in pinned host memory (the cloud is uploaded by both calls)'}
    cpu = None
    if not args.no_cpu_baseline:
        B.host_threads()
        bp = bld.filter_tile('2386_9702', bgt_types=['pand'], merge=True)
        rp = road.filter_tile('2386_9702', bgt_types=['rijbaan_lokale_weg', 'parkeervlak', 'fietspad'], merge=True)
        sec = []

        def stage(name, fn, lab):
            t0 = time.perf_counter()
            out = fn(lab)
            sec.append((name, round(time.perf_counter() - t0, 3)))
            return out
        o = np.zeros(n, dtype=np.uint16)
        t0 = time.perf_counter()
        o = stage('AHNFuser', lambda l: R.ahn_fuser_labels(points, l, l == 0, ahn_tile, 9, 'ground', 0.2), o)
        o = stage('BGTRoadFuser', lambda l: R.road_fuser_labels(points, l, rp, 1), o)
        o = stage('BGTBuildingFuser', lambda l: R.building_fuser_labels(points, l, l == 0, bp, 10, ahn_tile, 0.2), o)
        o = stage('LayerLCC', lambda l: R.layer_lcc_labels(points, l, l == 0, ahn_tile, 10, [dict(p) for p in layers]), o)
        o = stage('RegionGrowing', lambda l: R.region_growing_labels(points, l, 10, exclude_labels=(9,)), o)
        dt = time.perf_counter() - t0
        cpu = _cpu(n, dt, f'the full 4 M-point cloud, one pass ({dt:.1f} s)', 'oracle port, all processors', sec)
    _line(args, 'labelled Mpts/s, config 1 (notebook-0 processor list + RegionGrowing on the real demo tile)',
          '4 M synthetic points over the REAL demo AHN raster / BGT footprints of tile 2386_9702: AHNFuser(ground) + BGTRoadFuser + '
          'BGTBuildingFuser(+AHN cap) + LayerLCC(building, 2 layers), then RegionGrowing(building); offset = 0, refine_ground = False '
          '(no shapely)', n / (ms * 1e-3) / 1e6, ms, {'points': n}, None, cpu, e2e, launches)


def config2(args, torch, dev, _lib):
    """LabelConnectedComp on the ~11 M non-ground points of a 20 M-point tile, grid 0.05 / 0.1, min size 100 / 1500 / 5000."""
    from oracle import reference_py as R
    from urban_pointcloud_processing_b200 import synthetic
    from urban_pointcloud_processing_b200.region_growing import LabelConnectedComp
    n = args.points
    tile, kind, layout = synthetic.make_device_tile(n, seed=20240000, origin=synthetic.tile_origin(0))
    d_labels0 = torch.zeros(n, dtype=torch.int32, device=tile.device)
    d_labels0[kind == 0] = 9
    cars = torch.nonzero(kind == 2).flatten()
    d_labels0[cars[::100]] = 40
    d_mask = (d_labels0 == 0).to(torch.uint8)
    m = int((d_labels0 != 9).sum().item())
    variants = [(g, s) for g in (0.05, 0.1) for s in (100, 1500, 5000)]
    per, tot_ms = [], 0.0
    d_lab = torch.empty_like(d_labels0)
    _lib.prof_enable(True)
    l0 = _lib.launch_count()
    for g, s in variants:
        lcc = LabelConnectedComp(40, exclude_labels=(9,), grid_size=g, min_component_size=s, threshold=0.005)

        def step():
            d_lab.copy_(d_labels0)
            return lcc._label_mask_device(tile, d_lab, d_mask)
        ms, wall, out = _timed(torch, step, args.steps)
        per.append({'grid_size': g, 'min_component_size': s, 'level': lcc.octree_level, 'ms': round(ms, 3),
                    'Mpts_per_s': round(m / (ms * 1e-3) / 1e6, 1), 'points_labelled': int(out.sum().item())})
        tot_ms += ms
    launches = (_lib.launch_count() - l0) // (len(variants) * (args.steps + 1))
    prof = _lib.prof_read(); _lib.prof_enable(False)
    ms_avg = tot_ms / len(variants)
    # S3 (with the 4-byte keys of level <= 10 charged at 20 B per element and pass): 24 + 1 + 2 + 12 + 12 + 4 x 20 + 4 + 2
    b_alg = 137.0
    roof = {'kernel': 'LabelConnectedComp stage (keys, radix sort, cell list, union-find, output)', 'achieved': b_alg * m / (ms_avg * 1e-3) / 1e9,
            'algorithmic_bytes_per_unit': b_alg, 'units_per_launch': m, 'launch_ms': ms_avg,
            'kernels_ms_per_call': {k: round(v[0] / max(len(variants) * (args.steps + 1), 1), 3) for k, v in prof.items() if v[0] > 0}}
    pinned, pts_host = _pinned_points(torch, tile)
    labels_host = d_labels0.cpu().numpy().astype(np.uint16)
    mask_host = labels_host == 0
    lcc = LabelConnectedComp(40, exclude_labels=(9,), grid_size=0.05, min_component_size=1500, threshold=0.005)

    def host_step():
        dev.clear_tile_cache()
        return lcc.get_label_mask(pts_host, labels_host.copy(), mask_host)
    ems, ewall, got = _timed(torch, host_step, max(2, args.steps // 4))
    e2e = {'value': m / (max(ems, ewall) * 1e-3) / 1e6, 'unit': 'Mpts/s', 'h2d_bytes_per_step': int(n * 24 + n * 2 + n),
           'd2h_bytes_per_step': int(n), 'api': 'LabelConnectedComp.get_label_mask(points, labels, mask), grid 0.05 / min 1500, pinned host arrays'}
    cpu = None
    if not args.no_cpu_baseline:
        B.host_threads()
        t0 = time.perf_counter()
        want = R.lcc_get_label_mask(pts_host, labels_host.copy(), mask_host, 40, exclude_labels=(9,), grid_size=0.05, min_component_size=1500, threshold=0.005)
        dt = time.perf_counter() - t0
        cpu = _cpu(m, dt, f'the same {m} selected points, grid 0.05 / min 1500, one pass ({dt:.1f} s)', 'oracle port of LabelConnectedComp')
        e2e['labels_equal_oracle'] = bool(np.array_equal(got, want))
    _line(args, 'selected Mpts/s, config 2 (LabelConnectedComp)', f'LabelConnectedComp(car) on the {m} non-ground points of a synthetic '
          f'{n}-point tile, 1 % of the car points pre-labelled; grid 0.05 / 0.1 x min_component_size 100 / 1500 / 5000 (mean of the six)',
          m / (ms_avg * 1e-3) / 1e6, ms_avg, {'points': n, 'selected': m, 'variants': per}, roof, cpu, e2e, launches)


def config3(args, torch, dev, _lib):
    """Facade RegionGrowing from BGT building seeds on a 50 M-point facade-heavy tile, radius 0.1 / 0.2 / 0.5."""
    from oracle import reference_py as R
    from urban_pointcloud_processing_b200 import synthetic
    from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
    n = 50_000_000 if args.points == 20_000_000 else args.points
    tile, kind, layout = synthetic.make_device_tile(n, seed=20240000, origin=synthetic.tile_origin(0), facade_fraction=0.7)
    del kind
    tc = layout['tilecode']
    ahn = ahn_utils.ArrayAHNReader({tc: layout['ahn_tile']})
    bld = bgt_utils.ArrayBGTReader({tc: layout['building_polygons']})
    seeds = Pipeline(processors=(AHNFuser(9, ahn, target='ground', epsilon=0.2, refine_ground=False),
                                 BGTBuildingFuser(10, bgt_reader=bld, offset=0, ahn_reader=ahn)), ahn_reader=ahn, caching=True)
    d_labels0 = torch.zeros(n, dtype=torch.int32, device=tile.device)
    seeds.process_tile_device(tc, tile, d_labels0)
    m = int((d_labels0 != 9).sum().item())
    d_lab = torch.empty_like(d_labels0)
    per, tot_ms = [], 0.0
    _lib.prof_enable(True)
    l0 = _lib.launch_count()
    reps = max(1, args.steps // 4)
    for r in (0.1, 0.2, 0.5):
        rg = RegionGrowing(10, exclude_labels=(9,), grow_region_knn=15, grow_region_radius=r, max_nn=30)

        def step():
            d_lab.copy_(d_labels0)
            rg._apply_device(tile, d_lab, None, tc)
        ms, wall, _ = _timed(torch, step, reps)
        per.append({'grow_region_radius': r, 'ms': round(ms, 2), 'Mpts_per_s': round(m / (ms * 1e-3) / 1e6, 1),
                    'region_size': int((d_lab == 10).sum().item()), 'curvature_flags_on_host': rg.last_stats.get('n_curvature_on_host')})
        tot_ms += ms
    launches = (_lib.launch_count() - l0) // (3 * (reps + 1))
    prof = _lib.prof_read(); _lib.prof_enable(False)
    ms_avg = tot_ms / 3
    calls = 3 * (reps + 1)
    s4 = prof['knn_stage (build+search+epilogue)']
    roof = {'kernel': 'knn_stage (build+search+epilogue)', 'achieved': 140.0 * s4[2] / (s4[0] * 1e-3) / 1e9,
            'algorithmic_bytes_per_unit': 140.0, 'units_per_launch': s4[2] / max(s4[1], 1), 'launch_ms': s4[0] / max(s4[1], 1),
            'queries_per_s': s4[2] / (s4[0] * 1e-3),
            'kernels_ms_per_call': {k: round(v[0] / calls, 3) for k, v in prof.items() if v[0] > 0}}
    pinned, pts_host = _pinned_points(torch, tile)
    labels_host = d_labels0.cpu().numpy().astype(np.uint16)
    rg = RegionGrowing(10, exclude_labels=(9,))

    def host_step():
        dev.clear_tile_cache()
        return rg.get_labels(pts_host, labels_host.copy(), None, tc)
    ems, ewall, _ = _timed(torch, host_step, 1)
    e2e = {'value': m / (max(ems, ewall) * 1e-3) / 1e6, 'unit': 'Mpts/s', 'h2d_bytes_per_step': int(n * 24 + n * 2),
           'd2h_bytes_per_step': int(n), 'api': 'RegionGrowing.get_labels(points, labels, mask, tilecode), radius 0.2, pinned host arrays'}
    cpu = None
    if not args.no_cpu_baseline:
        B.host_threads()
        pts_c, lab_c, side = B.crop_sample(pts_host, labels_host, layout['boxes'][0], layout['origin'], args.cpu_sample)
        mc = int((lab_c != 9).sum())
        t0 = time.perf_counter()
        R.region_growing_labels(pts_c, lab_c.copy(), 10, exclude_labels=(9,))
        dt = time.perf_counter() - t0
        cpu = _cpu(mc, dt, f'{side:.1f} x {side:.1f} m full-density crop ({len(pts_c)} points, {mc} selected), radius 0.2, one pass ({dt:.1f} s)',
                   'oracle port of RegionGrowing (C natives for kNN / normals / growth; the reference loop is Python)')
    _line(args, 'selected Mpts/s, config 3 (RegionGrowing: neighbour search + normals + growth)',
          f'RegionGrowing(building) from BGT building seeds on a synthetic facade-heavy {n}-point tile ({m} non-ground points selected), '
          'knn 15, hybrid normals max_nn 30, radius 0.1 / 0.2 / 0.5 (mean of the three)', m / (ms_avg * 1e-3) / 1e6, ms_avg,
          {'points': n, 'selected': m, 'variants': per}, roof, cpu, e2e, launches)


def config4(args, torch, dev, _lib):
    """Point-in-polygon against 5 041 footprints + AHN cap + AHN ground fusion on 20 M points."""
    from oracle import reference_py as R
    from urban_pointcloud_processing_b200 import synthetic
    from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
    n = args.points
    tile, kind, layout = synthetic.make_device_tile(n, seed=20240000, origin=synthetic.tile_origin(0))
    del kind
    tc = layout['tilecode']
    polys = synthetic.make_polygon_lattice(n_side=71, seed=5000, origin=layout['origin'])
    ahn = ahn_utils.ArrayAHNReader({tc: layout['ahn_tile']})
    bld = bgt_utils.ArrayBGTReader({tc: polys})
    pipe = Pipeline(processors=(AHNFuser(9, ahn, target='ground', epsilon=0.2, refine_ground=False),
                                BGTBuildingFuser(10, bgt_reader=bld, offset=0, ahn_reader=ahn)), ahn_reader=ahn, caching=True)
    d_lab = torch.zeros(n, dtype=torch.int32, device=tile.device)

    def step():
        d_lab.zero_()
        pipe.process_tile_device(tc, tile, d_lab)
    _lib.prof_enable(True)
    l0 = _lib.launch_count()
    ms, wall, _ = _timed(torch, step, args.steps, warm=max(args.warmup, 3))
    launches = (_lib.launch_count() - l0) // (args.steps + max(args.warmup, 1))
    prof = _lib.prof_read(); _lib.prof_enable(False)
    calls = args.steps + max(args.warmup, 1)
    pk, rk = prof['pip_query_kernel'], prof['raster_kernel']
    peak = B.peaks()[0]
    roof = {'kernel': 'pip_query_kernel', 'achieved': 19.0 * pk[2] / (pk[0] * 1e-3) / 1e9, 'algorithmic_bytes_per_unit': 19.0,
            'units_per_launch': pk[2] / max(pk[1], 1), 'launch_ms': pk[0] / max(pk[1], 1),
            'raster_kernel': {'achieved': 27.0 * rk[2] / (rk[0] * 1e-3) / 1e9, 'frac': 27.0 * rk[2] / (rk[0] * 1e-3) / 1e9 / peak,
                              'launch_ms': rk[0] / max(rk[1], 1), 'launches_per_step': rk[1] / calls},
            'host_plan_build_ms_inside_step': round(ms - (pk[0] + rk[0]) / calls, 3),
            'polygons': len(polys), 'vertices': int(sum(len(p.exterior.coords) + sum(len(i.coords) for i in p.interiors) for p in polys))}
    pinned, pts_host = _pinned_points(torch, tile)
    lab_pinned = torch.zeros(n, dtype=torch.int16).pin_memory()
    labels_host = lab_pinned.numpy().view(np.uint16)

    def host_step():
        dev.clear_tile_cache()
        labels_host[:] = 0
        pipe.process_cloud(tc, pts_host, labels_host)
    ems, ewall, _ = _timed(torch, host_step, max(2, args.steps // 4))
    e2e = {'value': n / (max(ems, ewall) * 1e-3) / 1e6, 'unit': 'Mpts/s', 'h2d_bytes_per_step': int(n * 24 + n * 2),
           'd2h_bytes_per_step': int(n * 2), 'api': 'Pipeline.process_cloud (AHNFuser + BGTBuildingFuser), pinned host arrays'}
    cpu = None
    if not args.no_cpu_baseline:
        B.host_threads()
        k = 300_000                                   # the reference's poly_clip is one pass over the cloud PER POLYGON
        sub = np.ascontiguousarray(pts_host[:k])
        lab = np.zeros(k, dtype=np.uint16)
        t0 = time.perf_counter()
        lab = R.ahn_fuser_labels(sub, lab, lab == 0, layout['ahn_tile'], 9, 'ground', 0.2)
        lab = R.building_fuser_labels(sub, lab, lab == 0, polys, 10, layout['ahn_tile'], 0.2)
        dt = time.perf_counter() - t0
        cpu = _cpu(k, dt, f'the first {k} points of the tile against all {len(polys)} polygons, one pass ({dt:.1f} s)',
                   'oracle port of AHNFuser + BGTBuildingFuser (poly_clip loop over the polygons, FastGridInterpolator)')
        e2e['labels_equal_oracle_on_sample'] = bool(np.array_equal(labels_host[:k], lab))
    _line(args, 'labelled Mpts/s, config 4 (BGT point-in-polygon + AHN raster fusion)',
          f'AHNFuser(ground) + BGTBuildingFuser(+AHN cap) on a synthetic {n}-point tile against a jittered 71 x 71 lattice of '
          f'{len(polys)} star-shaped footprints (10 % with a hole)', n / (ms * 1e-3) / 1e6, ms, {'points': n}, roof, cpu, e2e, launches)


def run(args, rank, world):
    if rank != 0:
        return
    import torch
    from urban_pointcloud_processing_b200 import _lib
    from urban_pointcloud_processing_b200 import device as dev
    _lib.require_cuda()
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    global _SAMPLER
    _SAMPLER = B.ClockSampler(int(os.environ.get('LOCAL_RANK', '0')))
    _SAMPLER.start()
    {1: config1, 2: config2, 3: config3, 4: config4}[args.config](args, torch, dev, _lib)
