"""Phase timing of RegionGrowing on the bench tile (development tool)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from urban_pointcloud_processing_b200 import _lib, synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.region_growing import region_growing as rg
from oracle import reference_py as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
tile = synthetic.make_tile(n, seed=20240000)
pts = tile['points']
labels = np.zeros(n, dtype=np.uint16)
labels = R.ahn_fuser_labels(pts, labels, labels == 0, tile['ahn_tile'], 9, 'ground', 0.2)
labels = R.building_fuser_labels(pts, labels, labels == 0, tile['building_polygons'], 10, tile['ahn_tile'], 0.2)
t = dev.DeviceTile(pts)
d_labels = dev.upload_labels(labels, t.device)

def phase(name, fn, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(); acc.append((name, (time.perf_counter() - t0) * 1e3))
    return out

for rep in range(int(os.environ.get('REPS', '3'))):
    acc = []
    d_sel = phase('mask_build sel', lambda: dev.mask_build(d_labels, None, base=True, ne=(9, 1)), acc)
    bb = phase('bbox', lambda: t.bbox(d_sel), acc)
    _, bbox_sel, m = bb
    d_seed_full = phase('mask_build seed', lambda: dev.mask_build(d_labels, None, base=False, eq=(10,), ne=(9, 1)), acc)
    ns = phase('mask_count', lambda: dev.mask_count(d_seed_full), acc)
    out = phase('knn_normals', lambda: rg.knn_normals_sorted_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, 1.0, cell_size=(float(os.environ['CELL']) if 'CELL' in os.environ else None)), acc)
    d_knn, d_normals, d_cov, d_flag, d_order = out
    d_seed = phase('mask_take', lambda: rg.mask_take(d_seed_full, d_order), acc)
    gr = phase('region_grow+can_seed', lambda: rg.region_grow_overlapped(d_knn, d_normals, d_seed, d_flag, d_cov, 1.0, 20), acc)
    cs = (None, gr[2])
    sc = phase('mask_put', lambda: rg.mask_put(gr[0], d_order, t.n), acc)
print(f'm={m} seeds={ns} host-resolved curvature={cs[1]} region={int(gr[1][0])}')
for name, ms in acc:
    print(f'{name:18s} {ms:8.2f} ms')
print('total', sum(ms for _, ms in acc))
