"""Where the host curvature resolution spends its time (development tool)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from urban_pointcloud_processing_b200 import synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.region_growing import region_growing as rg
from oracle import reference_py as R

n = 20_000_000
tile = synthetic.make_tile(n, seed=20240000)
pts = tile['points']
labels = np.zeros(n, dtype=np.uint16)
labels = R.ahn_fuser_labels(pts, labels, labels == 0, tile['ahn_tile'], 9, 'ground', 0.2)
t = dev.DeviceTile(pts)
d_labels = dev.upload_labels(labels, t.device)
d_sel = dev.mask_build(d_labels, None, base=True, ne=(9,))
_, bbox_sel, m = t.bbox(d_sel)
d_knn, d_normals, d_cov, d_flag, d_order = rg.knn_normals_sorted_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, 1.0)
torch.cuda.synchronize()
for rep in range(3):
    T = [time.perf_counter()]
    def lap(): torch.cuda.synchronize(); T.append(time.perf_counter())
    idx = torch.nonzero(d_flag == 2).flatten(); lap()
    c = d_cov[idx].cpu().numpy(); lap()
    n_thr = 16
    parts = list(rg._pool().map(lambda a: rg._host_curvature(a, 1.0), np.array_split(c, n_thr))); res = np.concatenate(parts); lap()
    res1 = rg._host_curvature(c, 1.0); lap()
    up = torch.from_numpy(np.ascontiguousarray(res.astype(np.uint8))).to(t.device); lap()
    names = ['nonzero', 'cov fetch', 'eig x16 threads', 'eig 1 thread', 'upload']
    print(len(idx), {k: round((b - a) * 1e3, 2) for k, a, b in zip(names, T[:-1], T[1:])})

# the overlapped call itself, step by step (syncs inserted => no overlap; shows the GPU parts)
from urban_pointcloud_processing_b200 import _lib
lib = _lib.load()
d_seed_full = dev.mask_build(d_labels, None, base=False, eq=(10,))
labels2 = R.building_fuser_labels(pts, labels.copy(), labels == 0, tile['building_polygons'], 10, tile['ahn_tile'], 0.2)
d_labels2 = dev.upload_labels(labels2, t.device)
d_seed = rg.mask_take(dev.mask_build(d_labels2, None, base=False, eq=(10,)), d_order)
for rep in range(3):
    flag = d_flag.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = rg.region_grow_overlapped(d_knn, d_normals, d_seed, flag, d_cov, 1.0, 20)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    # GPU-only pieces
    flag = d_flag.clone()
    ws = dev.workspace(t.device, slot=2).get(lib.upcp_region_grow_ws_bytes(m))
    torch.cuda.synchronize(); a = time.perf_counter()
    _lib.check(lib.upcp_region_grow_begin(dev.ptr(d_knn), 15, dev.ptr(d_normals), dev.ptr(d_seed), dev.ptr(flag), m, 20.0, dev.ptr(ws), ws.numel(), dev.stream_ptr()))
    torch.cuda.synchronize(); b = time.perf_counter()
    d_region = torch.empty(m, dtype=torch.uint8, device=t.device); d_info = torch.zeros(2, dtype=torch.int64, device=t.device)
    _lib.check(lib.upcp_region_grow_end(m, dev.ptr(d_region), dev.ptr(d_info), dev.ptr(ws), ws.numel(), dev.stream_ptr()))
    torch.cuda.synchronize(); c2 = time.perf_counter()
    print(f'overlapped call {1e3 * (t1 - t0):.2f} ms; begin alone {1e3 * (b - a):.2f} ms; end alone {1e3 * (c2 - b):.2f} ms; region {int(out[1][0])}')
