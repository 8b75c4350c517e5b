#!/usr/bin/env python
"""bench.py -- labelled Mpts/s of the per-tile label pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl reference]

A step = one pass of the full label pipeline (AHN ground fusion, BGT road fusion, BGT
building fusion + AHN cap, LayerLCC building, LabelConnectedComp cars, RegionGrowing
building) over ONE synthetic 50 m urban tile of P (default 20 M) points.  With N > 1
(launched by torchrun, one rank per GPU) every rank labels its own tile per step (weak
scaling, no collective on the hot path); the per-tile label histograms are gathered with
one NCCL all_gather after the timed steps.

  value  device-resident throughput: tile coordinates already in HBM, labels reset on the
         device, pipeline chained on the device, label histogram left on the device.  Two
         regions of K steps each are timed: one tile at a time on one stream (`single_stream`;
         the per-kernel cudaEvent brackets, hence `roofline`, are taken here) and `--workers`
         tiles in flight, one CUDA stream + host thread each (`concurrent`,
         Pipeline.process_tiles_device).  `value` / `ms_per_step` are those of the faster region
         whose labels equal the single-stream ones; `config.tiles_in_flight_per_gpu` says which.
  e2e    the same pipeline through the reference-facing plugin API
         (Pipeline.process_cloud(tilecode, points, labels) with numpy arrays living in
         PINNED host memory): H2D of the points + labels and D2H of the labels inside the
         timed region.  Two public entry points are timed: `Pipeline.process_clouds` over the K
         tiles of the run (the upload of tile i+1 rides a second stream while tile i computes;
         reported as `e2e` when it is the faster one) and one `Pipeline.process_cloud` call per
         tile, strictly serial (`e2e_process_cloud`).
  roofline   for the kernel with the largest share of the step (cudaEvent brackets inside
         the library); algorithmic bytes per point from SURVEY.md 8(d) / DESIGN.md.  The
         `knn_search_kernel` bracket is the neighbour search proper: the two lane-per-query
         block tiers plus the warp-per-query tier (and the coarse grid built for it).
  cpu_baseline   the oracle port of the reference pipeline on the host cores, on a
         full-density crop of the same tile (rank 0, N = 1 only).

--impl reference times the reference's CPU implementation of the same pipeline (the oracle
port: reference control flow restated in numpy + plain-C/OpenMP natives; the reference's own
natives -- cccorelib, open3d, shapely -- are not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LAYERS = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
          {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
CAR = dict(grid_size=0.1, min_component_size=100, threshold=0.005)
METRIC = 'labelled Mpts/s per tile pipeline (CC + region growing + fusion)'
WORKLOAD = ('synthetic 50 m urban tile, {n} points: AHNFuser(ground) + BGTRoadFuser + BGTBuildingFuser(+AHN cap) '
            '+ LayerLCC(building, 2 layers) + LabelConnectedComp(car, grid 0.1) + RegionGrowing(building, knn 15, '
            'hybrid r 0.2 / max_nn 30)')
# algorithmic HBM bytes per point, SURVEY.md section 8(d)
MAX_BATCH_BUFFERS = 64        # label buffers kept alive at once by the batch arms (80 MB on the device / 40 MB pinned each)
B_ALG = {'raster_kernel': 27.0, 'pip_query_kernel': 19.0, 'radix_sort (hist+scan+scatter)': 96.0,
         'cc_link_kernel': 8.0, 'knn_search_kernel': 140.0, 'grow_kernel': 421.0}


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the bracketed kernel(s), from the
    committed ncu --set full capture of this command (profiles/r1_traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    try:
        return float(json.load(open(path))['bracket_traffic_bytes_per_launch'][kernel])
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Only the samples taken between t0 and t1 (perf_counter) count: the sampler is started
        before the warm-up so that nvidia-smi is already streaming when the timed region begins."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        t0, t1 = getattr(self, 't0', None), getattr(self, 't1', None)
        for stamp, line in self.lines:
            if t0 is not None and not (t0 <= stamp <= t1):
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_processors(ahn, bld, road):
    from urban_pointcloud_processing_b200.fusion import AHNFuser, BGTBuildingFuser, BGTRoadFuser
    from urban_pointcloud_processing_b200.labels import Labels
    from urban_pointcloud_processing_b200.region_growing import LabelConnectedComp, LayerLCC
    from urban_pointcloud_processing_b200.region_growing.region_growing import RegionGrowing
    return (AHNFuser(Labels.GROUND, ahn, target='ground', epsilon=0.2, refine_ground=False),
            BGTRoadFuser(Labels.ROAD, bgt_reader=road),
            BGTBuildingFuser(Labels.BUILDING, bgt_reader=bld, offset=0, ahn_reader=ahn),
            LayerLCC(Labels.BUILDING, ahn, params=[dict(p) for p in LAYERS]),
            LabelConnectedComp(Labels.CAR, **CAR),
            RegionGrowing(Labels.BUILDING, exclude_labels=(Labels.GROUND, Labels.ROAD)))


def oracle_pipeline(tile, points, labels):
    """The same processor list through the oracle (reference control flow, restated natives)."""
    from oracle import reference_py as R
    a = tile['ahn_tile']

    def car(p, l, m):
        lm = R.lcc_get_label_mask(p, l, m, 40, **CAR)
        l[lm] = 40
        return l
    return R.process_cloud([
        lambda p, l, m: R.ahn_fuser_labels(p, l, m, a, 9, 'ground', 0.2),
        lambda p, l, m: R.road_fuser_labels(p, l, tile['road_polygons'], 1),
        lambda p, l, m: R.building_fuser_labels(p, l, m, tile['building_polygons'], 10, a, 0.2),
        lambda p, l, m: R.layer_lcc_labels(p, l, m, a, 10, [dict(q) for q in LAYERS]),
        car,
        lambda p, l, m: R.region_growing_labels(p, l, 10, exclude_labels=(9, 1)),
    ], points, labels)


def initial_labels(tile, n):
    labels = np.zeros(n, dtype=np.uint16)
    cars = np.where(tile['kind'] == 2)[0]
    labels[cars[::100]] = 40            # 1 % of the car points pre-labelled as seeds (config 2)
    return labels


def crop_sample(tile, labels0, target):
    """Full-density spatial crop of about `target` points around the tile centre."""
    pts = tile['points']
    n = len(pts)
    side = 50.0 * np.sqrt(min(1.0, target / n))
    # centred on a corner of the first building so that facade, ground and street all occur
    cx, cy = tile['boxes'][0][2], tile['boxes'][0][3]
    cx = min(max(cx, tile['origin'][0] + side / 2), tile['origin'][0] + 50.0 - side / 2)
    cy = min(max(cy, tile['origin'][1] + side / 2), tile['origin'][1] + 50.0 - side / 2)
    keep = (np.abs(pts[:, 0] - cx) <= side / 2) & (np.abs(pts[:, 1] - cy) <= side / 2)
    return np.asfortranarray(pts[keep]), labels0[keep].copy(), side


def run_reference(args, rank, world):
    """--impl reference: the oracle port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import build as oracle_build
    from urban_pointcloud_processing_b200 import synthetic
    oracle_build.build()
    cores = os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    tile = synthetic.make_tile(args.points, seed=20240000)
    labels0 = initial_labels(tile, args.points)
    pts, lab, side = crop_sample(tile, labels0, args.ref_sample)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle_pipeline(tile, pts, lab.copy())
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = len(pts) * len(times) / total / 1e6
    sample = f'{side:.1f} x {side:.1f} m full-density crop ({len(pts)} points) of the {args.points}-point tile per step'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mpts/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.format(n=args.points), 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': 'Mpts/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'Mpts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--points', type=int, default=20_000_000)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--ref-sample', type=int, default=1_000_000, dest='ref_sample')
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, dest='cpu_sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--workers', type=int, default=4,
                    help='tiles in flight per GPU (one CUDA stream + host thread each) in the batch arms')
    ap.add_argument('--min-warmup', type=int, default=3, dest='min_warmup',
                    help='timing rule W >= 3; lower only for profiler runs whose numbers are never reported')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from urban_pointcloud_processing_b200 import _lib, batch, synthetic
    from urban_pointcloud_processing_b200 import device as dev
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils

    _lib.require_cuda()                      # no CPU fallback
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    warmup = max(args.warmup, args.min_warmup)             # timing rule: W >= 3
    n = args.points
    # every rank labels its own tile (seed 20240000 + rank); inputs (480 MB) >> 126 MB L2
    tile = synthetic.make_tile(n, seed=20240000 + rank)
    tc = tile['tilecode']
    labels0 = initial_labels(tile, n)
    ahn = ahn_utils.ArrayAHNReader({tc: tile['ahn_tile']})
    bld = bgt_utils.ArrayBGTReader({tc: tile['building_polygons']})
    road = bgt_utils.ArrayBGTReader({tc: tile['road_polygons']})
    pipeline = Pipeline(processors=make_processors(ahn, bld, road), ahn_reader=ahn, caching=True)

    # pinned host buffers for the e2e arm
    pts_pinned = torch.empty((3, n), dtype=torch.float64).pin_memory()
    pts_pinned.numpy()[...] = tile['points'].T
    points_host = pts_pinned.numpy().T           # (n, 3) F-ordered view, as pipeline.py:124 delivers
    lab_init_pinned = torch.from_numpy(labels0.view(np.int16)).pin_memory()
    lab_work_pinned = torch.empty(n, dtype=torch.int16).pin_memory()
    labels_host = lab_work_pinned.numpy().view(np.uint16)

    # device-resident arm
    d_tile = dev.DeviceTile(points_host, device)
    d_labels0 = dev.upload_labels(labels0, device)
    d_labels = torch.empty_like(d_labels0)
    torch.cuda.synchronize()

    def device_step():
        d_labels.copy_(d_labels0)
        pipeline.process_tile_device(tc, d_tile, d_labels)
        return dev.label_hist(d_labels)

    def e2e_step():
        dev.clear_tile_cache()
        lab_work_pinned.copy_(lab_init_pinned)
        pipeline.process_cloud(tc, points_host, labels_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1]), out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        hist = device_step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident pipeline
    _lib.prof_enable(True)
    launches0 = _lib.launch_count()
    t_begin = time.perf_counter()
    ms_dev, wall_dev, hist = timed(device_step, args.steps)
    launches = _lib.launch_count() - launches0
    prof = _lib.prof_read()
    _lib.prof_enable(False)

    # ---- timed region 2: the same K device-resident tile passes, `workers` tiles in flight (one stream and
    # host thread each, Pipeline.process_tiles_device).  No per-kernel brackets here: kernels of different
    # tiles overlap, the roofline numbers come from the single-stream region above.
    conc = None
    if args.workers > 1:
        bufs = [torch.empty_like(d_labels0) for _ in range(min(args.steps, MAX_BATCH_BUFFERS))]

        def concurrent_steps():
            # K tile passes; label buffers are recycled every MAX_BATCH_BUFFERS tiles (bounded memory for large K)
            h_ = None
            for first in range(0, args.steps, len(bufs)):
                part = bufs[:min(len(bufs), args.steps - first)]
                for b_ in part:
                    b_.copy_(d_labels0)
                pipeline.process_tiles_device([tc] * len(part), [d_tile] * len(part), part, workers=args.workers)
                h_ = [dev.label_hist(b_) for b_ in part][-1]
            return h_
        saved_steps, args.steps = args.steps, min(args.steps, 2 * args.workers)
        concurrent_steps()                                   # warm-up: worker streams, their scratch buffers
        args.steps = saved_steps
        launches1 = _lib.launch_count()
        ms_c, wall_c, hist_c = timed(concurrent_steps, 1)
        conc = {'ms_per_step': ms_c / args.steps, 'value': world * n * args.steps / (ms_c * 1e-3) / 1e6,
                'launches': _lib.launch_count() - launches1, 'wall_ms_per_step': wall_c / args.steps,
                'labels_equal_single_stream': bool(torch.equal(bufs[-1], d_labels) and torch.equal(bufs[0], d_labels))}
        del bufs
    t_end = time.perf_counter()
    if rank == 0:
        sampler.window(t_begin, t_end)
    clocks = sampler.stop() if rank == 0 else None
    stage_s = []
    pipeline.process_tile_device(tc, d_tile, d_labels.copy_(d_labels0), sync_timings=True)
    stage_s = [(name, round(sec * 1e3, 3)) for name, sec in pipeline.last_timings]

    # ---- end-of-batch gather of the label statistics (the only collective)
    all_hist = batch.gather_histograms([rank], hist.view(1, -1), world, device)
    labelled = int(all_hist[:, 1:].sum().item())

    # ---- e2e arm
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            e2e_step()
        ms_e2e, wall_e2e, _ = timed(e2e_step, args.steps)
        e2e_ms = max(ms_e2e, wall_e2e)          # host-side work is part of the call: take the wall clock
        e2e = {'value': world * n * args.steps / (e2e_ms * 1e-3) / 1e6, 'unit': 'Mpts/s',
               'ms_per_step': e2e_ms / args.steps,
               'h2d_bytes_per_step': int(n * 24 + n * 2), 'd2h_bytes_per_step': int(n * 2),
               'api': 'Pipeline.process_cloud(tilecode, points, labels) with numpy arrays in pinned host memory'}
        # the e2e labels equal the device-resident ones
        same = bool(np.array_equal(labels_host.astype(np.int32), d_labels.cpu().numpy()))
        e2e['labels_equal_device_arm'] = same

    # ---- e2e, batch form: Pipeline.process_clouds over `steps` tiles in pinned host memory -- the upload of
    # tile i+1 rides a second stream while the kernels of tile i run; every tile's H2D and D2H is inside
    # the timed region
    e2e_batch = None
    if not args.no_e2e:
        lab_bufs = [torch.empty(n, dtype=torch.int16).pin_memory() for _ in range(min(args.steps, MAX_BATCH_BUFFERS))]

        def reset_labels():
            for b_ in lab_bufs:
                b_.copy_(lab_init_pinned)

        def batch_call():
            # K tiles; beyond MAX_BATCH_BUFFERS the pinned label buffers are reset and recycled inside the timed region
            for first in range(0, args.steps, len(lab_bufs)):
                k_ = min(len(lab_bufs), args.steps - first)
                if first:
                    reset_labels()
                pipeline.process_clouds([tc] * k_, [points_host] * k_, [b_.numpy().view(np.uint16) for b_ in lab_bufs[:k_]],
                                        workers=args.workers)
        reset_labels()
        k_w = min(len(lab_bufs), 2 * max(args.workers, 1))
        pipeline.process_clouds([tc] * k_w, [points_host] * k_w, [b_.numpy().view(np.uint16) for b_ in lab_bufs[:k_w]],
                                workers=args.workers)
        reset_labels()
        ms_b, wall_b, _ = timed(batch_call, 1)
        b_ms = max(ms_b, wall_b)
        e2e_batch = {'value': world * n * args.steps / (b_ms * 1e-3) / 1e6, 'unit': 'Mpts/s',
                     'ms_per_step': b_ms / args.steps,
                     'h2d_bytes_per_step': int(n * 24 + n * 2), 'd2h_bytes_per_step': int(n * 2),
                     'api': f'Pipeline.process_clouds(tilecodes, clouds, labels_list, workers={args.workers}) over {args.steps} '
                            'tiles, numpy arrays in pinned host memory; each worker (host thread + CUDA stream) uploads, '
                            'labels and downloads one tile at a time, so copies and latency-bound phases of one tile '
                            'overlap the kernels of the others',
                     'labels_equal_device_arm': bool(all(
                         np.array_equal(b_.numpy().view(np.uint16).astype(np.int32), d_labels.cpu().numpy())
                         for b_ in (lab_bufs[0], lab_bufs[-1])))}
        del lab_bufs

    # ---- supplementary: the LAS-file route (Pipeline.process_file without the disk): raw int32 X, Y, Z
    # in pinned host memory -> device (12 B/pt), coordinates scaled on the device, labels back
    e2e_las = None
    if not args.no_e2e:
        from urban_pointcloud_processing_b200.utils import las_io
        scales, offsets = np.array([0.001, 0.001, 0.001]), np.array([100000.0, 400000.0, 0.0])
        ints = np.rint((tile['points'] - offsets) / scales).astype(np.int32)
        pinned_ints = [torch.from_numpy(np.ascontiguousarray(ints[:, k])).pin_memory() for k in range(3)]

        class _Las:                                   # the attributes DeviceTile.from_las reads
            X, Y, Z = (p_.numpy() for p_ in pinned_ints)
            header = las_io.LasHeader()
        _Las.header.scales, _Las.header.offsets = scales, offsets
        lab_u8 = torch.zeros(n, dtype=torch.uint8).pin_memory()

        def las_step():
            lab_u8.zero_()
            t_las = dev.DeviceTile.from_las(_Las, device)
            dl = dev.upload_labels(lab_u8.numpy(), device)
            pipeline.process_tile_device(tc, t_las, dl)
            dev.download_labels(dl, lab_u8.numpy())
        for _ in range(2):
            las_step()
        ms_las, wall_las, _ = timed(las_step, args.steps)
        las_ms = max(ms_las, wall_las)
        e2e_las = {'value': world * n * args.steps / (las_ms * 1e-3) / 1e6, 'unit': 'Mpts/s',
                   'ms_per_step': las_ms / args.steps, 'h2d_bytes_per_step': int(n * 12 + n), 'd2h_bytes_per_step': int(n),
                   'api': 'DeviceTile.from_las (int32 X, Y, Z + scale / offset, as Pipeline.process_file uploads them) + '
                          'process_tile_device + label download; no initial car seeds (labels start at 0)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    step_ms = max(ms_dev, 0.0) / args.steps
    value = world * n * args.steps / (ms_dev * 1e-3) / 1e6
    single = {'value': value, 'ms_per_step': step_ms, 'gpu_launches': int(launches),
              'what': 'one tile at a time on one stream (the region the roofline brackets were taken in)'}
    tiles_in_flight = 1
    if conc and conc['value'] > value and conc['labels_equal_single_stream']:
        value, step_ms, launches, tiles_in_flight = conc['value'], conc['ms_per_step'], conc['launches'], args.workers
        wall_dev = conc['wall_ms_per_step'] * args.steps
    peak, peak_src = peaks()
    # dominant kernel = largest bracketed share of the step
    top = max(prof.items(), key=lambda kv: kv[1][0])
    name, (tot_ms, cnt, units) = top
    per_launch_ms = tot_ms / max(cnt, 1)
    units_per_launch = units / max(cnt, 1)
    alg_bytes = B_ALG[name] * units_per_launch
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': measured_traffic(name), 'traffic_unit': 'bytes per launch (ncu dram read + write, profiles/r1_traffic.json)',
                'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src,
                'algorithmic_bytes_per_point': B_ALG[name], 'points_per_launch': units_per_launch,
                'launch_ms': per_launch_ms, 'share_of_step': tot_ms / max(ms_dev, 1e-9),
                'region': 'single-stream timed region (K steps, one tile at a time); `value` may come from the region with several tiles in flight',
                'kernels_ms_per_step': {k: round(v[0] / args.steps, 3) for k, v in prof.items()},
                # algorithmic GB/s of every bracketed kernel (B_ALG x units / bracketed time), fraction of the peak
                'kernels_algorithmic_gbs': {k: [round(B_ALG[k] * v[2] / (v[0] * 1e-3) / 1e9, 1),
                                                round(B_ALG[k] * v[2] / (v[0] * 1e-3) / 1e9 / peak, 4)]
                                            for k, v in prof.items() if v[0] > 0}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pts_c, lab_c, side = crop_sample(tile, labels0, args.cpu_sample)
        t0 = time.perf_counter()
        oracle_pipeline(tile, pts_c, lab_c)
        dt = time.perf_counter() - t0
        cpu_baseline = {'value': len(pts_c) / dt / 1e6, 'unit': 'Mpts/s', 'cores': cores, 'kind': 'port',
                        'sample': f'{side:.1f} x {side:.1f} m full-density crop ({len(pts_c)} points) of the same tile, '
                                  f'one pass of the oracle port ({dt:.1f} s)'}

    # headline e2e: the batch call when it is the faster of the two public entry points
    e2e_head = e2e_batch if (e2e_batch and e2e and e2e_batch['value'] >= e2e['value']) else e2e
    line = {'metric': METRIC, 'value': value, 'unit': 'Mpts/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.format(n=n), 'tiles_per_step': world, 'points_per_tile': n,
                       'l2_policy': 'inputs (480 MB of coordinates per tile) are larger than the 126 MB L2',
                       'tiles_per_s': value * 1e6 / n, 'tiles_in_flight_per_gpu': tiles_in_flight},
            'single_stream': single, 'concurrent': conc,
            'clocks': clocks, 'e2e': e2e_head, 'e2e_process_cloud': e2e, 'e2e_las': e2e_las, 'gpu_launches': int(launches), 'roofline': roofline,
            'cpu_baseline': cpu_baseline, 'stage_ms': stage_s, 'points_labelled_last_step': labelled,
            'wall_ms_per_step': wall_dev / args.steps}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
