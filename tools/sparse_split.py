"""How many of the sparse queries (15th neighbour farther than the normal radius) are initial seeds? (development tool)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from urban_pointcloud_processing_b200 import synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.region_growing import region_growing as rg
import bench
from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils
from urban_pointcloud_processing_b200.pipeline import Pipeline

n = 20_000_000
tile = synthetic.make_tile(n, seed=20240000)
tc = tile['tilecode']
labels = bench.initial_labels(tile, n)
ahn = ahn_utils.ArrayAHNReader({tc: tile['ahn_tile']})
bld = bgt_utils.ArrayBGTReader({tc: tile['building_polygons']})
road = bgt_utils.ArrayBGTReader({tc: tile['road_polygons']})
procs = bench.make_processors(ahn, bld, road)[:-1]
pts = np.asfortranarray(tile['points'])
Pipeline(processors=procs, ahn_reader=ahn, caching=True).process_cloud(tc, pts, labels)
sel = (labels != 9) & (labels != 1)
seed = labels[sel] == 10
t = dev.get_tile(pts)
d_sel = dev.upload_mask(sel, t.n, t.device)
_, bbox_sel, m = t.bbox(d_sel)
d_knn, d_d2, *_ = rg.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, want_d2=True, want_cov=False, want_curv=False)
d15 = d_d2[:, 14].cpu().numpy()
kind = tile['kind'][sel]
for thr in (0.1, 0.2, 0.3):
    sp = d15 > thr * thr
    print(f'15th neighbour farther than {thr} m: {sp.sum()} of {m}; of those initial seeds: {(sp & seed).sum()}; by kind {np.bincount(kind[sp], minlength=6)}; seeds by kind {np.bincount(kind[sp & seed], minlength=6)}')
print('selected', m, 'seeds', seed.sum())
