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
ap.add_argument('--cells', type=float, nargs='*', default=[0.0, 0.105, 0.09, 0.08, 0.07])
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--kinds', type=str, default='')
ap.add_argument('--tune', type=str, nargs='*', default=[],
                help='tuning variants to sweep, e.g. knn_tiers=1 knn_tier2_max_block=2000,knn_coarse=4')
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
variants = [{}] + [dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in v.split(',')) for v in args.tune]
for var in variants:
    _lib.tuning_reset()
    if var:
        print('tuning', _lib.tuning(**var))
    for cell in args.cells:
        for rep in range(args.reps):
            _lib.prof_enable(True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rg.knn_normals_sorted_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, 1.0, cell_size=cell if cell > 0 else None)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            prof = _lib.prof_read()
            _lib.prof_enable(False)
        ms = {k.split(' ')[0]: v[0] for k, v in prof.items()}
        cnt = dev.workspace(t.device).buf[:256].view(torch.int32)[:4].cpu().numpy()
        print(f'   queries left for tier 2: {cnt[1]}, for the per-query tier: {cnt[2]}')
        print(f'cell {cell:.3f}: wall {dt * 1e3:7.2f} ms | S4 {ms["knn_stage"]:6.2f} = build {ms["knn_build"]:5.2f} + search '
              f'{ms["knn_search"]:6.2f} + epilogue {ms["knn_epilogue_kernel"]:5.2f} ms -> {m / ms["knn_stage"] / 1e3:7.1f} Mq/s')
_lib.tuning_reset()
