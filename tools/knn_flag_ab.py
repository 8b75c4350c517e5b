"""A/B of an environment-controlled knob of the neighbour search inside ONE process (the library reads
its UPCP_* knobs on every call).  Development tool.
    python tools/knn_flag_ab.py UPCP_KNN_COARSE_BLOCK 0 3 4 5"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from urban_pointcloud_processing_b200 import _lib, synthetic            # noqa: E402
from urban_pointcloud_processing_b200 import device as dev              # noqa: E402
from urban_pointcloud_processing_b200.region_growing import region_growing as rg   # noqa: E402

name, values = sys.argv[1], sys.argv[2:]
tile = synthetic.make_tile(int(os.environ.get('N', '20000000')), seed=20240000)
sel = tile['kind'] != 0
t = dev.DeviceTile(tile['points'])
d_sel = dev.upload_mask(sel, t.n, t.device)
_, bbox_sel, m = t.bbox(d_sel)
ref = None
for v in values:
    os.environ[name] = v
    for rep in range(3):
        _lib.prof_enable(True)
        out = rg.knn_normals_device(t, d_sel, m, bbox_sel, 15, 30, 0.2)
        torch.cuda.synchronize()
        ms = _lib.prof_read()['knn_search_kernel'][0]
        _lib.prof_enable(False)
    knn, nrm = out[0].cpu().numpy(), out[2].cpu().numpy()
    if ref is None:
        ref = (knn, nrm)
    same = bool(np.array_equal(knn, ref[0]) and np.array_equal(nrm, ref[1]))
    print(f'{name}={v}: search {ms:7.2f} ms for {m} queries, kNN rows and normals equal to the first setting: {same}', flush=True)
