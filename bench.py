#!/usr/bin/env python
"""bench.py -- labelled Mpts/s of the per-tile label pipeline on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--tiles T] [--workers W]
                    [--config 1|2|3|4] [--impl reference]

A step = one pass of the full label pipeline (AHN ground fusion, BGT road fusion, BGT building
fusion + AHN cap, LayerLCC building, LabelConnectedComp cars, RegionGrowing building) over ONE
synthetic 50 m urban tile of P (default 20 M) points.  The tiles of a run are DISTINCT (BASELINE.json
config 5): tile i is generated on the device from seed 20240000 + i (upcp_synth_tile) on its own spot
of a tile mosaic, so it has its own tile code, AHN raster and BGT query.  With N > 1 (torchrun, one
rank per GPU) the batch is sharded tile-wise (weak scaling: K tiles per rank, no collective on the
hot path); the per-tile label histograms are gathered with one NCCL all_gather after the timed steps
(batch.run_batch).

  value  device-resident throughput: the K tiles of the rank already in HBM, their labels reset
         on the device.  Two regions of K steps each are timed: one tile at a time on one stream
         (`single_stream`; the per-kernel cudaEvent brackets, hence `roofline`, are taken here) and
         `--workers` tiles in flight, one CUDA stream + host thread each (`concurrent`:
         batch.run_batch -> Pipeline.process_tiles_device, every worker on a DIFFERENT tile, the
         readers' raster caches emptied first so that every tile pays its own raster uploads and
         point-in-polygon plan inside the timed region).  `value` / `ms_per_step` are those of the
         faster region whose labels equal the single-stream ones.
  e2e    the same pipeline through the public host API with numpy arrays in PINNED host memory: H2D
         of the points + labels and D2H of the labels inside the timed region.  `e2e` is
         Pipeline.process_clouds over the K tiles (the in-memory form of the reference's batch entry
         point Pipeline.process_folder), `e2e_process_cloud` is one strictly serial
         Pipeline.process_cloud(tilecode, points, labels) call per tile -- the reference's own
         per-tile signature.
  roofline   for the bracket with the largest share of the step (cudaEvent brackets inside the
         library); algorithmic bytes per point from SURVEY.md 8(d) / DESIGN.md.  The neighbour-search
         bracket spans the WHOLE stage S4 it is charged for (grid build, all search tiers, epilogue).
         `traffic` is only reported when profiles/r2_traffic.json was captured from these sources.
  cpu_baseline   the oracle port of the reference pipeline on the host cores, on a full-density crop of
         tile 0 (rank 0, N = 1 only), with its per-processor seconds.  (/root/reference does not travel
         to the GPU box and its natives -- cccorelib, open3d, shapely -- are not installable: the port is
         the only CPU arm possible there.)

--config 1..4 print the per-configuration lines of BASELINE.md section 4 instead (one JSON line each).
--impl reference times the oracle port on ALL host cores (rank 0 only under torchrun).
"""
import argparse
import hashlib
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
# algorithmic HBM bytes per unit of every bracket.  Points for all but the sort, whose bracket counts
# algorithmic BYTES itself (elements x passes x (3 key bytes + 8): 20 B per element and pass for the 4-byte
# keys of level <= 10).  The neighbour-search stage S4 is ONE bracket of 140 B/pt; its three sub-brackets
# (nested inside it) split that: build 24 + 16 + 16 (coordinates in, 32-byte + 16-byte sorted records out),
# epilogue 60 + 24 (kNN row + normal out), the search itself reads the 16-byte records.
B_ALG = {'raster_kernel': 27.0, 'pip_query_kernel': 19.0, 'radix_sort (hist+scan+scatter)': 1.0,
         'cc_link_kernel': 8.0, 'knn_stage (build+search+epilogue)': 140.0, 'grow_kernel': 421.0,
         'knn_build': 56.0, 'knn_search (block tiers + per-query tier)': 16.0, 'knn_epilogue_kernel': 84.0}
NESTED = ('knn_build', 'knn_search (block tiers + per-query tier)', 'knn_epilogue_kernel')   # parts of the S4 bracket


def sources_sha():
    """Fingerprint of the CUDA sources: a committed ncu capture is only quoted when it was taken from them."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'urban_pointcloud_processing_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.endswith(('.cu', '.cuh')):
            h.update(f.encode()); h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the bracketed kernel(s), from the committed
    ncu --set full capture (profiles/r2_traffic.json, tools/capture_profiles.sh) -- or None when there is
    none for the CUDA sources of this checkout (a stale capture is not quoted)."""
    path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    try:
        rec = json.load(open(path))
        if rec.get('csrc_sha') != sources_sha():
            return None
        return float(rec['bracket_traffic_bytes_per_launch'][kernel])
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


def oracle_pipeline(tile, points, labels, seconds=None):
    """The same processor list through the oracle (reference control flow, restated natives).  With
    ``seconds`` (a list) the wall time of every processor is appended, as the reference logs it
    (pipeline.py:94-95)."""
    from oracle import reference_py as R
    a = tile['ahn_tile']

    def car(p, l, m):
        lm = R.lcc_get_label_mask(p, l, m, 40, **CAR)
        l[lm] = 40
        return l
    stages = [
        ('AHNFuser', lambda p, l, m: R.ahn_fuser_labels(p, l, m, a, 9, 'ground', 0.2)),
        ('BGTRoadFuser', lambda p, l, m: R.road_fuser_labels(p, l, tile['road_polygons'], 1)),
        ('BGTBuildingFuser', lambda p, l, m: R.building_fuser_labels(p, l, m, tile['building_polygons'], 10, a, 0.2)),
        ('LayerLCC', lambda p, l, m: R.layer_lcc_labels(p, l, m, a, 10, [dict(q) for q in LAYERS])),
        ('LabelConnectedComp', car),
        ('RegionGrowing', lambda p, l, m: R.region_growing_labels(p, l, 10, exclude_labels=(9, 1))),
    ]

    def timed_stage(name, fn):
        def run(p, l, m):
            t0 = time.perf_counter()
            out = fn(p, l, m)
            if seconds is not None:
                seconds.append((name, round(time.perf_counter() - t0, 3)))
            return out
        return run
    return R.process_cloud([timed_stage(n_, f_) for n_, f_ in stages], points, labels)


def seed_labels_host(kind, n):
    labels = np.zeros(n, dtype=np.uint16)
    cars = np.where(kind == 2)[0]
    labels[cars[::100]] = 40            # 1 % of the car points pre-labelled as seeds (config 2)
    return labels


def crop_sample(points, labels0, box0, origin, target):
    """Full-density spatial crop of about `target` points, centred on a corner of the first building so that
    facade, ground and street all occur."""
    n = len(points)
    side = 50.0 * np.sqrt(min(1.0, target / n))
    cx, cy = box0[2], box0[3]
    cx = min(max(cx, origin[0] + side / 2), origin[0] + 50.0 - side / 2)
    cy = min(max(cy, origin[1] + side / 2), origin[1] + 50.0 - side / 2)
    keep = (np.abs(points[:, 0] - cx) <= side / 2) & (np.abs(points[:, 1] - cy) <= side / 2)
    return np.asfortranarray(points[keep]), labels0[keep].copy(), side


def host_threads():
    """All the host cores for the CPU arm -- SET, not defaulted: torchrun exports OMP_NUM_THREADS=1."""
    cores = os.cpu_count() or 1
    os.environ['OMP_NUM_THREADS'] = str(cores)
    os.environ['NUMBA_NUM_THREADS'] = str(cores)
    return cores


def run_reference(args, rank, world):
    """--impl reference: the oracle port on ALL host cores (rank 0 only; the other ranks exit without work)."""
    if rank != 0:
        return
    cores = host_threads()                       # before the OpenMP runtime of the oracle library starts
    from oracle import build as oracle_build
    from oracle import natives
    from urban_pointcloud_processing_b200 import synthetic
    oracle_build.build()
    threads = natives.set_num_threads(cores) if hasattr(natives, 'set_num_threads') else cores
    tile = synthetic.make_tile(args.points, seed=20240000)
    labels0 = seed_labels_host(tile['kind'], args.points)
    pts, lab, side = crop_sample(tile['points'], labels0, tile['boxes'][0], tile['origin'], args.ref_sample)
    times, per_proc = [], []
    for it in range(args.warmup + args.steps):
        sec = []
        t0 = time.perf_counter()
        oracle_pipeline(tile, pts, lab.copy(), sec)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
            per_proc = sec
    total = sum(times)
    value = len(pts) * len(times) / total / 1e6
    sample = f'{side:.1f} x {side:.1f} m full-density crop ({len(pts)} points) of the {args.points}-point tile per step'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mpts/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.format(n=args.points), 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': 'Mpts/s', 'cores': cores, 'omp_threads': int(threads), 'kind': 'port',
                             'sample': sample, 'per_processor_s': per_proc,
                             'what': 'reference control flow restated in numpy + plain-C / OpenMP natives '
                                     '(oracle/): /root/reference does not travel to the GPU box and its natives '
                                     '(cccorelib, open3d, shapely) are not installable'},
            'e2e': {'value': value, 'unit': 'Mpts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


class Batch:
    """The K distinct device-resident tiles of this rank (tile index = rank + world * j) with their readers."""

    def __init__(self, n, count, rank, world, device):
        import torch
        from urban_pointcloud_processing_b200 import synthetic
        from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
        self.n, self.count = n, count
        self.index = [rank + world * j for j in range(count)]
        self.tiles, self.codes, self.labels0, self.layouts = [], [], [], []
        for g in self.index:
            tile, kind, layout = synthetic.make_device_tile(n, seed=20240000 + g, origin=synthetic.tile_origin(g), device=device)
            lab = torch.zeros(n, dtype=torch.int32, device=device)
            cars = torch.nonzero(kind == 2).flatten()
            lab[cars[::100]] = 40        # 1 % of the car points pre-labelled as seeds (config 2)
            self.tiles.append(tile); self.codes.append(layout['tilecode']); self.labels0.append(lab)
            self.layouts.append(layout)
            del kind
        self.ahn = ahn_utils.ArrayAHNReader({L['tilecode']: L['ahn_tile'] for L in self.layouts})
        self.bld = bgt_utils.ArrayBGTReader({L['tilecode']: L['building_polygons'] for L in self.layouts})
        self.road = bgt_utils.ArrayBGTReader({L['tilecode']: L['road_polygons'] for L in self.layouts})

    def cold_caches(self):
        """Forget the device rasters: the next pass over the batch uploads them again, tile by tile."""
        with self.ahn._lock:
            self.ahn._rasters.clear()

    def host_points(self, j):
        """(n, 3) float64 F-ordered numpy view of tile j in PINNED host memory."""
        import torch
        t = self.tiles[j]
        pinned = torch.empty((3, self.n), dtype=torch.float64).pin_memory()
        pinned[0].copy_(t.x); pinned[1].copy_(t.y); pinned[2].copy_(t.z)
        torch.cuda.synchronize()
        return pinned, pinned.numpy().T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--points', type=int, default=20_000_000)
    ap.add_argument('--tiles', type=int, default=0,
                    help='distinct device-resident tiles per rank (default: min(steps, 24)); steps cycle through them')
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=0, help='1..4: the per-configuration lines of BASELINE.md section 4')
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
    if args.config:
        import bench_configs
        bench_configs.run(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from urban_pointcloud_processing_b200 import _lib, batch
    from urban_pointcloud_processing_b200 import device as dev
    from urban_pointcloud_processing_b200.pipeline import Pipeline

    _lib.require_cuda()                      # no CPU fallback
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    warmup = max(args.warmup, args.min_warmup)             # timing rule: W >= 3
    n, K = args.points, args.steps
    T = args.tiles if args.tiles > 0 else min(K, 24)
    # the rank's tiles, generated on the device (coordinates: 480 MB per tile >> 126 MB L2)
    B = Batch(n, T, rank, world, device)
    pipeline = Pipeline(processors=make_processors(B.ahn, B.bld, B.road), ahn_reader=B.ahn, caching=True)
    d_labels = torch.empty_like(B.labels0[0])
    torch.cuda.synchronize()

    def device_step(j):
        j %= T
        d_labels.copy_(B.labels0[j])
        pipeline.process_tile_device(B.codes[j], B.tiles[j], d_labels)
        return dev.label_hist(d_labels)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """fn() runs the K steps; device time by CUDA events on the current stream (the workers' streams join
        it through events / synchronisation), wall clock beside it, MAX over ranks."""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1]), out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for s in range(warmup):
        device_step(s)
    torch.cuda.synchronize()

    # ---- timed region 1: one tile at a time on one stream (the per-kernel brackets live here)
    _lib.prof_enable(True)
    launches0 = _lib.launch_count()
    t_begin = time.perf_counter()
    ms_dev, wall_dev, hist = timed(lambda: [device_step(s) for s in range(K)][-1])
    launches = _lib.launch_count() - launches0
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    single_labels = {}
    for j in (0, (K - 1) % T):                # reference labels for the equality checks of the other arms
        device_step(j)
        single_labels[j] = d_labels.clone()

    # ---- timed region 2: the same K tile passes through batch.run_batch -> Pipeline.process_tiles_device,
    # `workers` DIFFERENT tiles in flight (one stream + host thread each), raster caches cold.  No per-kernel
    # brackets here: kernels of different tiles overlap.
    conc = None
    all_hist = None
    if args.workers > 1:
        bufs = [torch.empty_like(B.labels0[0]) for _ in range(min(K, MAX_BATCH_BUFFERS))]

        def label_tiles(global_ids):
            # this rank's share of the batch: global tile id g = rank + world * s  ->  step s
            steps = [(g - rank) // world for g in global_ids]
            hists = []
            for first in range(0, len(steps), len(bufs)):            # label buffers recycled beyond MAX_BATCH_BUFFERS
                part = steps[first:first + len(bufs)]
                for b_, s in zip(bufs, part):
                    b_.copy_(B.labels0[s % T])
                pipeline.process_tiles_device([B.codes[s % T] for s in part], [B.tiles[s % T] for s in part],
                                              bufs[:len(part)], workers=args.workers)
                hists += [dev.label_hist(b_) for b_ in bufs[:len(part)]]
            return torch.stack(hists)

        def batch_pass(k_steps, stop=None):
            return batch.run_batch(world * k_steps, None, device, rank=rank, world_size=world,
                                   process_tiles=label_tiles, on_processed=stop)[0]
        batch_pass(min(K, 2 * args.workers))                         # warm-up: worker streams, their scratch buffers
        B.cold_caches()
        launches1 = _lib.launch_count()
        mark = {}

        def stop_clock():                                            # end of the K steps: before the end-of-batch gather
            mark['e'] = torch.cuda.Event(enable_timing=True); mark['e'].record(); mark['t'] = time.perf_counter()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        all_hist = batch_pass(K, stop_clock)
        torch.cuda.synchronize()
        tt = torch.tensor([max(e0.elapsed_time(mark['e']), 0.0), (mark['t'] - t0) * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        barrier()
        ms_c, wall_c = max(float(tt[0]), float(tt[1])), float(tt[1])     # worker streams: the wall clock bounds the region
        same = bool(torch.equal(bufs[0], single_labels[0]) if K <= len(bufs) else True) and \
            bool(torch.equal(bufs[(K - 1) % len(bufs)], single_labels[(K - 1) % T]))
        conc = {'ms_per_step': ms_c / K, 'value': world * n * K / (ms_c * 1e-3) / 1e6,
                'launches': _lib.launch_count() - launches1, 'wall_ms_per_step': wall_c / K,
                'labels_equal_single_stream': same,
                'what': f'batch.run_batch -> Pipeline.process_tiles_device, {args.workers} different tiles in flight, '
                        'raster caches cold (every tile uploads its rasters and builds its point-in-polygon plans)'}
        del bufs
    t_end = time.perf_counter()
    if rank == 0:
        sampler.window(t_begin, t_end)
    clocks = sampler.stop() if rank == 0 else None
    d_labels.copy_(B.labels0[0])
    pipeline.process_tile_device(B.codes[0], B.tiles[0], d_labels, sync_timings=True)
    stage_s = [(name, round(sec * 1e3, 3)) for name, sec in pipeline.last_timings]

    # ---- end-of-batch gather of the label statistics (the only collective)
    if all_hist is None:
        all_hist = batch.gather_histograms([rank], hist.view(1, -1), world, device)
    labelled = int(all_hist[-world:, 1:].sum().item())

    # ---- e2e arms: host buffers (pinned), H2D + D2H inside the timed region
    e2e = e2e_batch = e2e_las = e2e_laz = None
    if not args.no_e2e:
        Th = min(T, 4)                                    # distinct host tiles (0.5 GB of pinned memory each)
        host = [B.host_points(j) for j in range(Th)]
        lab_init = [torch.from_numpy(B.labels0[j].cpu().numpy().astype(np.int16)).pin_memory() for j in range(Th)]
        lab_work = torch.empty(n, dtype=torch.int16).pin_memory()
        labels_host = lab_work.numpy().view(np.uint16)

        def e2e_step(s):
            j = s % Th
            lab_work.copy_(lab_init[j])
            pipeline.process_cloud(B.codes[j], host[j][1], labels_host)
        for s in range(2):
            e2e_step(s)
        ms_e, wall_e, _ = timed(lambda: [e2e_step(s) for s in range(K)])
        e_ms = max(ms_e, wall_e)                          # host-side work is part of the call: take the wall clock
        e2e = {'value': world * n * K / (e_ms * 1e-3) / 1e6, 'unit': 'Mpts/s', 'ms_per_step': e_ms / K,
               'h2d_bytes_per_step': int(n * 24 + n * 2), 'd2h_bytes_per_step': int(n * 2),
               'api': 'Pipeline.process_cloud(tilecode, points, labels), one call per tile, numpy arrays in pinned host memory',
               'labels_equal_device_arm': bool(np.array_equal(labels_host.astype(np.int32),
                                                              single_labels[(K - 1) % Th].cpu().numpy()))
               if (K - 1) % Th in single_labels else None}

        lab_bufs = [torch.empty(n, dtype=torch.int16).pin_memory() for _ in range(min(K, MAX_BATCH_BUFFERS))]

        def reset_labels(first=0):
            for i_, b_ in enumerate(lab_bufs):
                b_.copy_(lab_init[(first + i_) % Th])

        def batch_call():
            # K tiles; beyond MAX_BATCH_BUFFERS the pinned label buffers are reset and recycled inside the timed region
            for first in range(0, K, len(lab_bufs)):
                k_ = min(len(lab_bufs), K - first)
                if first:
                    reset_labels(first)
                pipeline.process_clouds([B.codes[(first + i_) % Th] for i_ in range(k_)],
                                        [host[(first + i_) % Th][1] for i_ in range(k_)],
                                        [b_.numpy().view(np.uint16) for b_ in lab_bufs[:k_]], workers=args.workers)
        reset_labels()
        k_w = min(len(lab_bufs), 2 * max(args.workers, 1))
        pipeline.process_clouds([B.codes[i_ % Th] for i_ in range(k_w)], [host[i_ % Th][1] for i_ in range(k_w)],
                                [b_.numpy().view(np.uint16) for b_ in lab_bufs[:k_w]], workers=args.workers)
        reset_labels()
        B.cold_caches()
        ms_b, wall_b, _ = timed(batch_call)
        b_ms = max(ms_b, wall_b)
        e2e_batch = {'value': world * n * K / (b_ms * 1e-3) / 1e6, 'unit': 'Mpts/s', 'ms_per_step': b_ms / K,
                     'h2d_bytes_per_step': int(n * 24 + n * 2), 'd2h_bytes_per_step': int(n * 2),
                     'api': f'Pipeline.process_clouds(tilecodes, clouds, labels_list, workers={args.workers}) over {K} tiles '
                            f'({Th} distinct clouds in pinned host memory): the in-memory form of the reference\'s batch entry '
                            'point Pipeline.process_folder; each worker (host thread + CUDA stream) uploads, labels and '
                            'downloads one tile at a time',
                     'labels_equal_device_arm': bool(np.array_equal(lab_bufs[0].numpy().view(np.uint16).astype(np.int32),
                                                                    single_labels[0].cpu().numpy())) if K <= len(lab_bufs) else None}
        del lab_bufs

        # supplementary: the LAS-file route (Pipeline.process_file without the disk): raw int32 X, Y, Z in pinned
        # host memory -> device (12 B/pt), coordinates scaled on the device, labels back
        from urban_pointcloud_processing_b200.utils import las_io
        scales, offsets = np.array([0.001, 0.001, 0.001]), np.array([100000.0, 400000.0, 0.0])
        ints = np.rint((host[0][1] - offsets) / scales).astype(np.int32)
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
            pipeline.process_tile_device(B.codes[0], t_las, dl)
            dev.download_labels(dl, lab_u8.numpy())
        for _ in range(2):
            las_step()
        ms_las, wall_las, _ = timed(lambda: [las_step() for _ in range(K)])
        las_ms = max(ms_las, wall_las)
        e2e_las = {'value': world * n * K / (las_ms * 1e-3) / 1e6, 'unit': 'Mpts/s', 'ms_per_step': las_ms / K,
                   'h2d_bytes_per_step': int(n * 12 + n), 'd2h_bytes_per_step': int(n),
                   'api': 'DeviceTile.from_las (int32 X, Y, Z + scale / offset, as Pipeline.process_file uploads them) + '
                          'process_tile_device + label download; no initial car seeds (labels start at 0)'}

        # supplementary: the reference's own file route, Pipeline.process_file from a `.laz` on disk to a labelled
        # `.laz` on disk (native LASzip decoder / encoder on the host threads; rank 0 only, two files)
        e2e_laz = None
        if rank == 0:
            import shutil
            import tempfile
            td = tempfile.mkdtemp(prefix='upcp_bench_')
            try:
                src = os.path.join(td, f'filtered_{B.codes[0]}.laz')
                las0 = las_io.create(ints, scales, offsets, point_format_id=1)
                las0.write(src)
                del las0
                t0_ = time.perf_counter()
                reps_ = 2
                for r_ in range(reps_):
                    pipeline.process_file(src, os.path.join(td, f'labelled_{r_}_{B.codes[0]}.laz'))
                torch.cuda.synchronize()
                laz_s = (time.perf_counter() - t0_) / reps_
                e2e_laz = {'value': n / laz_s / 1e6, 'unit': 'Mpts/s', 'ms_per_step': laz_s * 1e3, 'files': reps_,
                           'laz_bytes_in': os.path.getsize(src),
                           'laz_bytes_out': os.path.getsize(os.path.join(td, f'labelled_0_{B.codes[0]}.laz')),
                           'api': 'Pipeline.process_file(in.laz, out.laz): LASzip decode (host threads) -> int32 upload -> '
                                  'label pipeline -> label download -> LASzip encode; one file after the other '
                                  '(process_folder overlaps the codec of neighbouring files with the labelling)'}
            finally:
                shutil.rmtree(td, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    step_ms = max(ms_dev, 0.0) / K
    value = world * n * K / (ms_dev * 1e-3) / 1e6
    single = {'value': value, 'ms_per_step': step_ms, 'gpu_launches': int(launches),
              'what': 'one tile at a time on one stream, warm raster caches (the region the roofline brackets were taken in)'}
    tiles_in_flight = 1
    if conc and conc['value'] > value and conc['labels_equal_single_stream']:
        value, step_ms, launches, tiles_in_flight = conc['value'], conc['ms_per_step'], conc['launches'], args.workers
        wall_dev = conc['wall_ms_per_step'] * K
    peak, peak_src = peaks()
    # dominant bracket = largest share of the step (the sub-brackets of S4 are parts of its bracket)
    name, (tot_ms, cnt, units) = max(((k, v) for k, v in prof.items() if k not in NESTED), key=lambda kv: kv[1][0])
    per_launch_ms = tot_ms / max(cnt, 1)
    units_per_launch = units / max(cnt, 1)
    alg_bytes = B_ALG[name] * units_per_launch
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    traffic = measured_traffic(name)
    roofline = {'bound': 'hbm', 'kernel': name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic,
                'traffic_unit': 'bytes per launch (ncu dram read + write, profiles/r2_traffic.json)' if traffic is not None
                else 'null: no ncu capture of these CUDA sources is committed (profiles/r2_traffic.json csrc_sha mismatch)',
                'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src,
                'algorithmic_bytes_per_unit': B_ALG[name], 'units_per_launch': units_per_launch,
                'launch_ms': per_launch_ms, 'share_of_step': tot_ms / max(ms_dev, 1e-9),
                'region': 'single-stream timed region (K steps, one tile at a time); `value` may come from the region with several tiles in flight',
                'kernels_ms_per_step': {k: round(v[0] / K, 3) for k, v in prof.items()},
                # algorithmic GB/s of every bracket (B_ALG x units / bracketed time), fraction of the peak
                'kernels_algorithmic_gbs': {k: [round(B_ALG[k] * v[2] / (v[0] * 1e-3) / 1e9, 1),
                                                round(B_ALG[k] * v[2] / (v[0] * 1e-3) / 1e9 / peak, 4)]
                                            for k, v in prof.items() if v[0] > 0},
                'csrc_sha': sources_sha()}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        from oracle import natives
        threads = natives.set_num_threads(cores) if hasattr(natives, 'set_num_threads') else cores
        t0_ = B.tiles[0]
        pts0 = np.asfortranarray(np.column_stack((t0_.x.cpu().numpy(), t0_.y.cpu().numpy(), t0_.z.cpu().numpy())))
        lab0 = B.labels0[0].cpu().numpy().astype(np.uint16)
        L0 = B.layouts[0]
        pts_c, lab_c, side = crop_sample(pts0, lab0, L0['boxes'][0], L0['origin'], args.cpu_sample)
        sec = []
        t0 = time.perf_counter()
        oracle_pipeline(L0, pts_c, lab_c, sec)
        dt = time.perf_counter() - t0
        cpu_baseline = {'value': len(pts_c) / dt / 1e6, 'unit': 'Mpts/s', 'cores': cores, 'omp_threads': int(threads),
                        'kind': 'port', 'per_processor_s': sec,
                        'sample': f'{side:.1f} x {side:.1f} m full-density crop ({len(pts_c)} points) of tile 0, '
                                  f'one pass of the oracle port ({dt:.1f} s)',
                        'what': 'reference control flow restated in numpy + plain-C / OpenMP natives (oracle/): the only '
                                'CPU arm possible on the GPU box (/root/reference does not travel; cccorelib, open3d, '
                                'shapely are not installable)'}

    line = {'metric': METRIC, 'value': value, 'unit': 'Mpts/s', 'n_gpus': world, 'steps': K,
            'warmup': warmup, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.format(n=n), 'tiles_per_step': world, 'points_per_tile': n,
                       'distinct_tiles_per_gpu': T, 'tile_source': 'generated on the device from seed 20240000 + tile index '
                       '(upcp_synth_tile), each on its own tile code / AHN raster / BGT polygons',
                       'l2_policy': 'inputs (480 MB of coordinates per tile, a different tile every step) are larger than the 126 MB L2',
                       'tiles_per_s': value * 1e6 / n, 'tiles_in_flight_per_gpu': tiles_in_flight},
            'single_stream': single, 'concurrent': conc,
            'clocks': clocks, 'e2e': e2e_batch or e2e, 'e2e_process_cloud': e2e, 'e2e_las': e2e_las, 'e2e_laz_file': e2e_laz,
            'gpu_launches': int(launches), 'roofline': roofline,
            'cpu_baseline': cpu_baseline, 'stage_ms': stage_s, 'points_labelled_last_step': labelled,
            'wall_ms_per_step': wall_dev / K}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
