"""Experiment: two tiles in flight on two CUDA streams driven by two host threads (development tool).
Does the device fill the latency-bound phases of one tile (frontier growth, union-find link, host round
trips) with the kernels of the other?  Prints tiles/s for 1, 2 and 3 workers.
(The experiment that preceded Pipeline.process_tiles_device / process_clouds(workers=W), which is the product
form: per-stream workspaces live in device.workspace now, the thread-local patch below predates that.)"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                             # noqa: E402
from urban_pointcloud_processing_b200 import synthetic                  # noqa: E402
from urban_pointcloud_processing_b200 import device as dev              # noqa: E402
from urban_pointcloud_processing_b200.pipeline import Pipeline          # noqa: E402
from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils  # noqa: E402

n = int(os.environ.get('N', '20000000'))
steps = int(os.environ.get('STEPS', '8'))
tile = synthetic.make_tile(n, seed=20240000)
tc = tile['tilecode']
labels0 = bench.initial_labels(tile, n)
device = torch.device('cuda', 0)
d_tile = dev.DeviceTile(tile['points'], device)
d_labels0 = dev.upload_labels(labels0, device)

_tls = threading.local()
_orig_ws = dev.Workspace


def workspace(device, slot=0):
    table = getattr(_tls, 'ws', None)
    if table is None:
        table = _tls.ws = {}
    key = (str(device), slot)
    if key not in table:
        table[key] = _orig_ws(device)
    return table[key]


dev.workspace = workspace


def make_pipeline():
    ahn = ahn_utils.ArrayAHNReader({tc: tile['ahn_tile']})
    bld = bgt_utils.ArrayBGTReader({tc: tile['building_polygons']})
    road = bgt_utils.ArrayBGTReader({tc: tile['road_polygons']})
    return Pipeline(processors=bench.make_processors(ahn, bld, road), ahn_reader=ahn, caching=True)


def worker(k, count, out):
    stream = torch.cuda.Stream(device)
    pipe = make_pipeline()
    with torch.cuda.stream(stream):
        d_labels = torch.empty_like(d_labels0)
        for _ in range(2):                                  # warm-up: rasters, workspaces
            d_labels.copy_(d_labels0)
            pipe.process_tile_device(tc, d_tile, d_labels)
        stream.synchronize()
        out['ready'].wait()
        for _ in range(count):
            d_labels.copy_(d_labels0)
            pipe.process_tile_device(tc, d_tile, d_labels)
        stream.synchronize()
        out[k] = d_labels.cpu().numpy()


for workers in (1, 2, 3):
    out = {'ready': threading.Barrier(workers + 1)}
    threads = [threading.Thread(target=worker, args=(k, steps, out)) for k in range(workers)]
    for t in threads:
        t.start()
    out['ready'].wait()
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    same = all(np.array_equal(out[0], out[k]) for k in range(workers))
    print(f'{workers} worker(s): {workers * steps / dt:6.2f} tiles/s, {dt / (workers * steps) * 1e3:6.2f} ms per tile, '
          f'{workers * steps * n / dt / 1e6:7.1f} Mpts/s, labels equal across workers: {same}', flush=True)
