"""Host-side (Python) cost of one tile through Pipeline.process_tile_device: cProfile over a few tiles with the
device kept busy, so what shows is interpreter / ctypes / torch dispatch time, not kernel time.  Development tool."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
batch = B.Batch(n, 3, 0, 1, torch.device("cuda:0"))
from urban_pointcloud_processing_b200.pipeline import Pipeline      # noqa: E402
pipe = Pipeline(processors=B.make_processors(batch.ahn, batch.bld, batch.road), ahn_reader=batch.ahn, caching=True)
labs = [batch.labels0[j].clone() for j in range(3)]
pipe.process_tile_device(batch.codes[0], batch.tiles[0], labs[0])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for j in (1, 2):
    batch.cold_caches()
    pipe.process_tile_device(batch.codes[j], batch.tiles[j], labs[j])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
