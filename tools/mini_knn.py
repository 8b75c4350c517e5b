import sys, numpy as np
sys.path.insert(0, '.')
from urban_pointcloud_processing_b200 import synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.region_growing import region_growing as rg
import torch
tile = synthetic.make_tile(30000, seed=3, facade_fraction=0.7)
t = dev.DeviceTile(tile['points'])
_, bbox, m = t.bbox(None)
out = rg.knn_normals_device(t, None, m, bbox, 15, 30, 0.2)
torch.cuda.synchronize()
print('ok', out[0][:2])
