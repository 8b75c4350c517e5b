"""Micro-benchmark of the neighbour-search stage (kNN + normals) on the bench tile's
non-ground cloud, for a sweep of fine-cell sizes.  Development tool, not a reported number."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from urban_pointcloud_processing_b200 import _lib, synthetic            # noqa: E402
from urban_pointcloud_processing_b200 import device as dev              # noqa: E402
from urban_pointcloud_processing_b200.region_growing import region_growing as rg   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--points', type=int, default=20_000_000)
ap.add_argument('--cells', type=float, nargs='*', default=[0.1, 0.07, 0.05, 0.035])
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--kinds', type=str, default='')
args = ap.parse_args()

tile = synthetic.make_tile(args.points, seed=20240000)
pts = tile['points']
sel = tile['kind'] != 0
if args.kinds:
    sel = np.isin(tile['kind'], [int(k) for k in args.kinds.split(',')])
t = dev.DeviceTile(pts)
d_sel = dev.upload_mask(sel, t.n, t.device)
_, bbox_sel, m = t.bbox(d_sel)
print(f'points {len(pts)}, selected {m}, default cell {rg.default_cell_size(bbox_sel, m, 0.2):.4f}')
for cell in args.cells:
    for rep in range(args.reps):
        _lib.prof_enable(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rg.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, cell_size=cell)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        prof = _lib.prof_read()
        _lib.prof_enable(False)
    ms = prof['knn_search_kernel'][0]
    print(f'cell {cell:.3f}: stage {dt * 1e3:8.2f} ms, search kernel {ms:8.2f} ms = {m / ms / 1e3:7.1f} Mq/s, '
          f'sort {prof["radix_sort (hist+scan+scatter)"][0]:.2f} ms')
