"""Which points does the device leave to the host curvature test, and why? (development tool)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from urban_pointcloud_processing_b200 import synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.region_growing import region_growing as rg

tile = synthetic.make_tile(int(os.environ.get("N", "20000000")), seed=20240000)
pts = tile['points']
sel = tile['kind'] != 0
t = dev.DeviceTile(pts)
d_sel = dev.upload_mask(sel, t.n, t.device)
_, bbox_sel, m = t.bbox(d_sel)
d_knn, d_d2, d_nrm, d_cov, d_curv, d_flag = rg.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2, want_d2=True, threshold_curve=1.0)
flag = d_flag.cpu().numpy()
idx = np.where(flag == 2)[0]
print('m', m, 'undecided', len(idx), 'flag0', (flag == 0).sum())
curv = d_curv.cpu().numpy()[idx]
cov = d_cov.cpu().numpy()[idx]
d2 = d_d2.cpu().numpy()[idx]
kinds = tile['kind'][sel][idx]
print('kinds of undecided', np.bincount(kinds, minlength=6))
print('nan rows', np.isnan(curv).any(axis=1).sum(), 'ratio>=1-1e-9', (curv >= 1 - 1e-9).any(axis=1).sum())
for i in range(min(8, len(idx))):
    print('curv', curv[i], 'cov', cov[i], 'd2[-1]', d2[i][-1])
c = cov
M = np.empty((len(c), 3, 3)); M[:, 0, 0] = c[:, 0]; M[:, 0, 1] = M[:, 1, 0] = c[:, 1]; M[:, 0, 2] = M[:, 2, 0] = c[:, 2]
M[:, 1, 1] = c[:, 3]; M[:, 1, 2] = M[:, 2, 1] = c[:, 4]; M[:, 2, 2] = c[:, 5]
ev = np.linalg.eig(M)[0]
res = ev[:, 0] / ev.sum(axis=1) < 1.0
print('host result True:', res.sum(), 'False:', (~res).sum())
